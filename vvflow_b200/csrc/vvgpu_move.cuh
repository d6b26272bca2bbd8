// K6 — MFlowmove::move_and_clean, particle part (libvvhd/src/MFlowmove.cpp:107-144,194-199) with
// TBody::isPointInvalid -> isPointInContour (libvvhd/src/TBody.cpp:217-220,238-281), plus the
// AoS <-> SoA marshalling kernels of the host boundary (TObj, libvvhd/headers/TObj.hpp:10-16).
#pragma once
#include "vvgpu_near.cuh"

namespace vv {

struct BodyGeom {
    int nseg, nbody;
    const double *rx, *ry, *cx, *cy;
    const int* bfirst;
    const double* bprop;  // 16 per body
};

// nearest segment of body ib if p is an invalid (in-body) point, else -1
__device__ __forceinline__ int point_invalid(const BodyGeom& B, int ib, double px, double py) {
    const double* bp = B.bprop + 16 * ib;
    bool in = bp[12] != 0;  // isInsideValid()
    if (!in) {
        double dx = VV_SUB(px, bp[2]), dy = VV_SUB(py, bp[3]);
        if (px < bp[4] || py < bp[5] || px > bp[6] || py > bp[7] ||
            VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy)) > bp[8]) return -1;
    }
    const int f = B.bfirst[ib], e = B.bfirst[ib + 1];
    double vjx = B.cx[e - 1], vjy = B.cy[e - 1];
    for (int i = f; i < e; i++) {
        double vix = B.cx[i], viy = B.cy[i];
        double lhs = VV_MUL(VV_SUB(vjy, viy), VV_SUB(px, vix));
        double rhs = VV_MUL(VV_SUB(vjx, vix), VV_SUB(py, viy));
        if (((viy < vjy) && (viy < py) && (py <= vjy) && (lhs > rhs)) ||
            ((viy > vjy) && (viy > py) && (py >= vjy) && (lhs < rhs)))
            in = !in;
        vjx = vix; vjy = viy;
    }
    if (!in) return -1;
    int nearest = -1;
    double nd = DBL_MAX;
    for (int s = f; s < e; s++) {
        double dx = VV_SUB(B.rx[s], px), dy = VV_SUB(B.ry[s], py);
        double d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
        if (d < nd) { nearest = s; nd = d; }
    }
    return nearest;
}

// advect (:107), decide removal (:113-117 small |g|, :126-143 in-body), accumulate the dead sums
__global__ void k_move_flag(int n, Particles P, double dt, double remove_eps, int remove, BodyGeom B, u32* keep,
                            double* fdt_dead, double* g_dead, double* gsum, unsigned long long* cleaned) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = VV_ADD(P.x[i], VV_MUL(P.vx[i], dt));
    double y = VV_ADD(P.y[i], VV_MUL(P.vy[i], dt));
    P.x[i] = x; P.y[i] = y;
    double g = P.g[i];
    bool k = !(fabs(g) < remove_eps);
    if (k && remove) {
        for (int ib = 0; ib < B.nbody; ib++) {
            int s = point_invalid(B, ib, x, y);
            if (s < 0) continue;
            const double* bp = B.bprop + 16 * ib;
            atomicAdd(&fdt_dead[3 * ib + 0], -y * g);   // rotl(r)*g
            atomicAdd(&fdt_dead[3 * ib + 1], x * g);
            double ax = x - bp[0], ay = y - bp[1];
            atomicAdd(&fdt_dead[3 * ib + 2], (ax * ax + ay * ay) * g);
            atomicAdd(&gsum[s], -g);
            atomicAdd(&g_dead[ib], g);
            atomicAdd(cleaned, 1ull);
            k = false;
            break;
        }
    }
    keep[i] = k ? 1u : 0u;
}

// stable compaction of the survivors; v is zeroed (:194-199)
__global__ void k_move_compact(int n, const u32* __restrict__ scan, Particles in, const int* __restrict__ orig_in,
                               Particles out, int* __restrict__ orig_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 d = scan[i];
    if (scan[i + 1] == d) return;
    out.x[d] = in.x[i]; out.y[d] = in.y[i]; out.g[d] = in.g[i];
    out.vx[d] = 0; out.vy[d] = 0; out.ie[d] = in.ie[i];
    orig_out[d] = orig_in[i];
}

// Space::gsum() (TSpace.hpp): total circulation of the list, for the SLAE's circulation equation when the list lives on
// the device. One CTA, fixed association order (deterministic; not the reference's sequential order: ~1e-16 relative).
__global__ void __launch_bounds__(1024) k_sum_g(int n, const double* __restrict__ g, double* out) {
    __shared__ double sh[1024];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 1024) s += g[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// ---- host boundary ---------------------------------------------------------------------------
__global__ void k_unpack48(int n, const double* __restrict__ rec, Particles P, int* orig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = rec + 6ll * i;
    P.x[i] = r[0]; P.y[i] = r[1]; P.g[i] = r[2]; P.vx[i] = r[3]; P.vy[i] = r[4]; P.ie[i] = r[5];
    orig[i] = i;
}
// records appended behind the `at` resident ones (vvgpu_append_particles)
__global__ void k_unpack48_at(int n, const double* __restrict__ rec, Particles P, int* orig, int at, int orig_base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = rec + 6ll * i;
    const int d = at + i;
    P.x[d] = r[0]; P.y[d] = r[1]; P.g[d] = r[2]; P.vx[d] = r[3]; P.vy[d] = r[4]; P.ie[d] = r[5];
    orig[d] = orig_base + i;
}
// records that arrived as one padded slice per rank (vvgpu_set_particles_slice): record i sits in the slice of the rank
// r with n r / P <= i < n (r + 1) / P, at (i - n r / P) of that slice
__global__ void k_unpack48_slices(int n, int P_, int per, const double* __restrict__ rec, Particles P, int* orig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = (int)(((long long)i * P_) / n);
    while ((long long)n * (r + 1) / P_ <= i) r++;
    while ((long long)n * r / P_ > i) r--;
    const double* q = rec + 6ll * ((long long)r * per + (i - (long long)n * r / P_));
    P.x[i] = q[0]; P.y[i] = q[1]; P.g[i] = q[2]; P.vx[i] = q[3]; P.vy[i] = q[4]; P.ie[i] = q[5];
    orig[i] = i;
}
__global__ void k_unpack24(int n, const double* __restrict__ rec, Particles P, int* orig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = rec + 3ll * i;
    P.x[i] = r[0]; P.y[i] = r[1]; P.g[i] = r[2]; P.vx[i] = 0; P.vy[i] = 0; P.ie[i] = 0;
    orig[i] = i;
}
__global__ void k_pack48(int n, Particles P, double* __restrict__ rec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* r = rec + 6ll * i;
    r[0] = P.x[i]; r[1] = P.y[i]; r[2] = P.g[i]; r[3] = P.vx[i]; r[4] = P.vy[i]; r[5] = P.ie[i];
}

// ---- FP64 pipe micro-benchmark (roofline denominator for K3-K5) -------------------------------
__global__ void k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int k = 0; k < iters; k++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace vv
