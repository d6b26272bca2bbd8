// Scratch micro-benchmark for the leaf-warp layout of the near-field kernels: lanes hold sources in
// registers, the <=15 targets of one leaf are broadcast from shared memory, accumulators stay in
// registers. Measures pairs/s per variant and the accuracy of rcp.approx.ftz.f64.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

constexpr int kT = 15;

// MODE 0: rcp + cubic (10 DP/pair)   MODE 1: rcp + one Newton step (9 DP/pair)   MODE 2: no rcp, 9 DP
template <int MODE, int NT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_leafwarp(double* out, int iters, const double4* src, int nsrc) {
    __shared__ double2 tgt[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 16) tgt[warp][lane] = make_double2(0.001 * lane + 0.01 * warp, 0.002 * lane + blockIdx.x * 1e-4);
    __syncwarp();
    double ax[kT], ay[kT];
#pragma unroll
    for (int t = 0; t < kT; t++) ax[t] = ay[t] = 0;
    double4 s = src[lane];
    for (int it = 0; it < iters; it++) {
        double4 nx = src[((it + 1) * 32 + lane) & (nsrc - 1)];   // prefetch the next source
#pragma unroll
        for (int t = 0; t < NT; t++) {
            double2 p = tgt[warp][t];
            double dx = p.x - s.x, dy = p.y - s.y;
            double den = fma(dx, dx, fma(dy, dy, s.w));
            double w;
            if (MODE == 0) {
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
                double e = fma(-den, r0, 1.0);
                double e2 = fma(e, e, e);
                double gr = s.z * r0;
                w = fma(gr, e2, gr);
            } else if (MODE == 1) {
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
                double e = fma(-den, r0, 1.0);
                double gr = s.z * r0;
                w = fma(gr, e, gr);
            } else {
                double e = fma(-den, den, 1.0);
                double gr = s.z * den;
                w = fma(gr, e, gr);
            }
            ax[t] = fma(-dy, w, ax[t]);
            ay[t] = fma(dx, w, ay[t]);
        }
        s = nx;
    }
    double r = 0;
#pragma unroll
    for (int t = 0; t < kT; t++) r += ax[t] + ay[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int NT, int MINB>
void run(const char* name, double* out, const double4* src, int nsrc, int blocks) {
    int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_leafwarp<MODE, NT, MINB><<<blocks, 256>>>(out, 16, src, nsrc);
    cudaEventRecord(a);
    k_leafwarp<MODE, NT, MINB><<<blocks, 256>>>(out, iters, src, nsrc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double pairs = (double)blocks * 256 * iters * NT;
    printf("%-34s minb=%d NT=%2d blocks=%d  %.3f ms  %.1f Gpairs/s\n", name, MINB, NT, blocks, ms, pairs / ms / 1e6);
}

__global__ void k_rcp_err(double* maxerr, double lo, double ratio, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double worst = 0;
    double x = lo * pow(ratio, (double)i);
    for (int k = 0; k < n; k++) {
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
        double e = fabs(fma(-x, r0, 1.0));
        worst = fmax(worst, e);
        x *= 1.0000001192092896;
    }
    atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(worst));
}

int main() {
    double* out; double4* src; double* me;
    const int nsrc = 1 << 16;
    cudaMalloc(&out, 148 * 64 * 256 * 8);
    cudaMalloc(&src, nsrc * sizeof(double4));
    cudaMalloc(&me, 8);
    double4* h = new double4[nsrc];
    for (int i = 0; i < nsrc; i++) { h[i].x = 0.37 + 1e-5 * i; h[i].y = 1.0 + 0.001 * (i % 977); h[i].z = 1e-6; h[i].w = 1e-8; }
    cudaMemcpy(src, h, nsrc * sizeof(double4), cudaMemcpyHostToDevice);
    cudaMemset(me, 0, 8);
    k_rcp_err<<<1024, 256>>>(me, 1e-12, 1.0001, 4096);
    double herr;
    cudaMemcpy(&herr, me, 8, cudaMemcpyDeviceToHost);
    printf("rcp.approx.ftz.f64 max |1 - x*r0| = %.3e = 2^%.2f\n", herr, log2(herr));
    for (int bpsm : {1, 2, 3, 4}) {
        int blocks = 148 * bpsm;
        run<0, 15, 1>("leafwarp rcp+cubic (10 DP)", out, src, nsrc, blocks);
        run<1, 15, 1>("leafwarp rcp+newton (9 DP)", out, src, nsrc, blocks);
        run<2, 15, 1>("leafwarp no rcp (9 DP)", out, src, nsrc, blocks);
        run<1, 10, 1>("leafwarp rcp+newton (9 DP)", out, src, nsrc, blocks);
        run<0, 15, 2>("leafwarp rcp+cubic (10 DP)", out, src, nsrc, blocks);
        run<1, 15, 2>("leafwarp rcp+newton (9 DP)", out, src, nsrc, blocks);
        run<1, 15, 3>("leafwarp rcp+newton (9 DP)", out, src, nsrc, blocks);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
