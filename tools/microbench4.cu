// Scratch: inner-loop structure of the convective leaf-warp kernel (lanes = sources, <= 15 targets broadcast
// from shared memory, 2 x 15 accumulators in registers). Variants differ in sources per lane (1 / 2), in how
// the pair bodies of a target group are ordered in the source (nested vs explicitly staged around the MUFU),
// and in CTAs per SM. Prints pairs/s per variant; registers come from -Xptxas -v.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kT = 15;
constexpr int kIters = 32;     // iterations of 64 sources per leaf

__device__ __forceinline__ double rcp_approx(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
}
__device__ __forceinline__ void pair(const double2 p, const double4 s, double& ax, double& ay) {
    const double dx = p.x - s.x, dy = p.y - s.y;
    const double den = fma(dx, dx, fma(dy, dy, s.w));
    const double r0 = rcp_approx(den);
    const double e = fma(-den, r0, 1.0);
    const double gr = s.z * r0;
    const double w = fma(gr, e, gr);
    ax = fma(-dy, w, ax);
    ay = fma(dx, w, ay);
}
// STYLE 0: nested (what the kernel does today): for t: pair(s), pair(s2)
// STYLE 1: staged per group and source: all dx,dy,den of the group; all rcp; all tails
template <int STYLE, int NS, int BASE, int N>
__device__ __forceinline__ void group(const double2* txy, const double4& s, const double4& s2, double (&ax)[kT], double (&ay)[kT]) {
    if (STYLE == 0) {
#pragma unroll
        for (int t = 0; t < N; t++) {
            const double2 p = txy[BASE + t];
            pair(p, s, ax[BASE + t], ay[BASE + t]);
            if (NS == 2) pair(p, s2, ax[BASE + t], ay[BASE + t]);
        }
    } else {
        double dx[2 * N], dy[2 * N], den[2 * N], r0[2 * N];
#pragma unroll
        for (int t = 0; t < N; t++) {
            const double2 p = txy[BASE + t];
            dx[t] = p.x - s.x; dy[t] = p.y - s.y;
            den[t] = fma(dx[t], dx[t], fma(dy[t], dy[t], s.w));
            if (NS == 2) {
                dx[N + t] = p.x - s2.x; dy[N + t] = p.y - s2.y;
                den[N + t] = fma(dx[N + t], dx[N + t], fma(dy[N + t], dy[N + t], s2.w));
            }
        }
#pragma unroll
        for (int t = 0; t < NS * N; t++) r0[t] = rcp_approx(den[t]);
#pragma unroll
        for (int t = 0; t < NS * N; t++) {
            const double g = (t < N) ? s.z : s2.z;
            const double e = fma(-den[t], r0[t], 1.0);
            const double gr = g * r0[t];
            const double w = fma(gr, e, gr);
            const int a = BASE + (t < N ? t : t - N);
            ax[a] = fma(-dy[t], w, ax[a]);
            ay[a] = fma(dx[t], w, ay[a]);
        }
    }
}

template <int STYLE, int NS, int NG, int MINB, bool UNROLL2>
__global__ void __launch_bounds__(128, MINB) k_var(double* out, int leaves, const double4* src, int nsrc) {
    __shared__ double2 tgt[4][16];
    __shared__ int idx[4][kIters * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = lane; k < kIters * 64; k += 32) idx[warp][k] = (k * 7 + blockIdx.x * 131 + warp * 17) & (nsrc - 1);
    if (lane < 16) tgt[warp][lane] = make_double2(0.001 * lane + 0.01 * warp, 0.002 * lane + blockIdx.x * 1e-4);
    __syncwarp();
    const int* ip = idx[warp];
    const double2* txy = tgt[warp];
    double tot = 0;
    constexpr int STEP = 32 * NS;
    constexpr int NIT = kIters * 64 / STEP;
    for (int leaf = 0; leaf < leaves; leaf++) {
        double ax[kT], ay[kT];
#pragma unroll
        for (int t = 0; t < kT; t++) ax[t] = ay[t] = 0;
        double4 s = src[ip[lane]], s2 = s;
        if (NS == 2) s2 = src[ip[lane + 32]];
        if (!UNROLL2) {
            for (int it = 0; it < NIT; it++) {
                const int k = ((it + 1) * STEP + lane) % (kIters * 64);
                double4 nx = src[ip[k]], nx2 = nx;
                if (NS == 2) nx2 = src[ip[(k + 32) % (kIters * 64)]];
                group<STYLE, NS, 0, 5>(txy, s, s2, ax, ay);
                if (NG > 1) group<STYLE, NS, 5, 5>(txy, s, s2, ax, ay);
                if (NG > 2) group<STYLE, NS, 10, 5>(txy, s, s2, ax, ay);
                s = nx; s2 = nx2;
            }
        } else {
            // ping-pong source registers: no moves at the end of an iteration
            for (int it = 0; it < NIT; it += 2) {
                const int k = ((it + 1) * STEP + lane) % (kIters * 64);
                double4 b = src[ip[k]], b2 = b;
                if (NS == 2) b2 = src[ip[(k + 32) % (kIters * 64)]];
                group<STYLE, NS, 0, 5>(txy, s, s2, ax, ay);
                if (NG > 1) group<STYLE, NS, 5, 5>(txy, s, s2, ax, ay);
                if (NG > 2) group<STYLE, NS, 10, 5>(txy, s, s2, ax, ay);
                const int k2 = ((it + 2) * STEP + lane) % (kIters * 64);
                s = src[ip[k2]];
                if (NS == 2) s2 = src[ip[(k2 + 32) % (kIters * 64)]];
                group<STYLE, NS, 0, 5>(txy, b, b2, ax, ay);
                if (NG > 1) group<STYLE, NS, 5, 5>(txy, b, b2, ax, ay);
                if (NG > 2) group<STYLE, NS, 10, 5>(txy, b, b2, ax, ay);
            }
        }
#pragma unroll
        for (int t = 0; t < kT; t++) tot += ax[t] + ay[t];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <int STYLE, int NS, int NG, int MINB, bool UNROLL2>
void run(double* out, const double4* src, int nsrc) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_var<STYLE, NS, NG, MINB, UNROLL2>, 128, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_var<STYLE, NS, NG, MINB, UNROLL2>);
    const int blocks = 148 * nb, leaves = 40;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_var<STYLE, NS, NG, MINB, UNROLL2><<<blocks, 128>>>(out, 2, src, nsrc);
    cudaEventRecord(a);
    k_var<STYLE, NS, NG, MINB, UNROLL2><<<blocks, 128>>>(out, leaves, src, nsrc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double pairs = (double)blocks * 4 * leaves * (kIters * 64) * (NG * 5);
    printf("style=%d ns=%d groups=%d minb=%d unroll2=%d regs=%3d ctas/SM=%d warps/SMSP=%d : %7.1f Gpairs/s  (%s)\n", STYLE, NS, NG, MINB,
           (int)UNROLL2, fa.numRegs, nb, nb, pairs / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* out; double4* src; const int nsrc = 1 << 16;
    cudaMalloc(&out, 148 * 16 * 128 * 8); cudaMalloc(&src, nsrc * sizeof(double4));
    double4* h = new double4[nsrc];
    for (int i = 0; i < nsrc; i++) { h[i].x = 0.37 + 1e-5 * i; h[i].y = 1.0 + 0.001 * (i % 977); h[i].z = 1e-6; h[i].w = 1e-8; }
    cudaMemcpy(src, h, nsrc * sizeof(double4), cudaMemcpyHostToDevice);
#define RUN3(ST, NS, U) run<ST, NS, 3, 2, U>(out, src, nsrc); run<ST, NS, 3, 3, U>(out, src, nsrc); run<ST, NS, 3, 4, U>(out, src, nsrc); \
                        run<ST, NS, 2, 3, U>(out, src, nsrc); run<ST, NS, 2, 4, U>(out, src, nsrc); run<ST, NS, 2, 5, U>(out, src, nsrc);
    RUN3(0, 2, false)
    RUN3(0, 2, true)
    RUN3(1, 2, false)
    RUN3(1, 2, true)
    RUN3(0, 1, false)
    RUN3(0, 1, true)
    RUN3(1, 1, false)
    RUN3(1, 1, true)
    return 0;
}
