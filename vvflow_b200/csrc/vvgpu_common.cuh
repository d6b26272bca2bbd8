// Shared device helpers for the vvgpu kernels (sm_100a).
//
// Arithmetic discipline: the reference (libvvhd, built for baseline x86-64, CMakeLists.txt:29) is
// strict IEEE double WITHOUT fused multiply-add. Every value that feeds a comparison or an index
// decision (tree boxes, split tests, far criterion, neighbour distances, merge criteria, cut-offs,
// point-in-polygon) is therefore computed with the explicit round-to-nearest intrinsics below,
// which nvcc never contracts. Accumulations whose result only has to agree to 1e-10 use fma().
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

#define VV_MUL(a, b) __dmul_rn((a), (b))
#define VV_ADD(a, b) __dadd_rn((a), (b))
#define VV_SUB(a, b) __dsub_rn((a), (b))

namespace vv {

constexpr double kPi = 3.14159265358979323846;  // elementary.h:3
constexpr double k2Pi = 2. * kPi;
constexpr double k1_2Pi = 1. / (2. * kPi);
constexpr double k1_Pi = 1. / kPi;

typedef unsigned long long u64;
typedef unsigned int u32;

// order-preserving map double -> u64, so that tight bounding boxes are exact atomicMin/Max
__device__ __forceinline__ u64 enc_ordered(double d) {
    u64 b = (u64)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ordered(u64 e) {
    u64 b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
__host__ __device__ __forceinline__ int sgn(double v) { return (v > 0) ? 1 : ((v < 0) ? -1 : 0); }  // TObj.hpp:8
// the sink strength as MConvectiveFast.cpp:165 sees it: its bare `abs(src.g)` resolves to ::abs(int) in the reference
// build, i.e. the circulation is truncated to an integer first (pinned against the compiled reference in the tests)
__device__ __forceinline__ double sink_abs(double g) { const int t = (int)g; return (double)(t < 0 ? -t : t); }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }

__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---------------------------------------------------------------------------------------------
// Exclusive prefix sum of a 0/1 (or small-count) functor over [0, n): out has n+1 entries.
// Three phases (tile sums, scan of tile sums, rescan + write); tiles of 4096 items.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <int THREADS>
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total, u32* sh /* THREADS/32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = (lane < THREADS / 32) ? sh[lane] : 0;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < THREADS / 32) sh[lane] = winc - w;
        if (lane == THREADS / 32 - 1) sh[THREADS / 32] = winc;
    }
    __syncthreads();
    u32 res = sh[warp] + inc - v;
    *total = sh[THREADS / 32];
    __syncthreads();
    return res;
}

// the same for any integer type (e.g. several 16-bit counters packed into a u64)
template <class T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan_t(T v, T* total, T* sh /* THREADS/32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < THREADS / 32) ? sh[lane] : 0;
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < THREADS / 32) sh[lane] = winc - w;
        if (lane == THREADS / 32 - 1) sh[THREADS / 32] = winc;
    }
    __syncthreads();
    T res = sh[warp] + inc - v;
    *total = sh[THREADS / 32];
    __syncthreads();
    return res;
}

template <class F>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(F f, long long n, u32* partial) {
    __shared__ u32 sh[kScanThreads / 32 + 1];
    long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += f(base + k);
    u32 total;
    block_exclusive_scan<kScanThreads>(s, &total, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single CTA: in-place exclusive scan of the tile sums (any count)
__global__ void __launch_bounds__(1024) k_scan_partials(u32* partial, int nparts) {
    __shared__ u32 sh[1024 / 32 + 1];
    u32 carry = 0;
    for (int base = 0; base < nparts; base += 1024) {
        int i = base + threadIdx.x;
        u32 v = (i < nparts) ? partial[i] : 0;
        u32 total;
        u32 ex = block_exclusive_scan<1024>(v, &total, sh);
        if (i < nparts) partial[i] = carry + ex;
        carry += total;
    }
}

template <class F>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(F f, long long n, const u32* partial, u32* out) {
    __shared__ u32 sh[kScanThreads / 32 + 1];
    long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    u32 v[kScanItems];
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? f(base + k) : 0;
        s += v[k];
    }
    u32 total;
    u32 ex = block_exclusive_scan<kScanThreads>(s, &total, sh) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

struct FlagArray {
    const u32* a;
    __device__ __forceinline__ u32 operator()(long long i) const { return a[i]; }
};

}  // namespace vv
