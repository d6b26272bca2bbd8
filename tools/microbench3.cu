// Scratch: FP64 pair-body throughput as a function of independent chains per warp (NT) and warps per
// SM sub-partition — how much ILP x TLP the Biot-Savart body needs to fill the B200 FP64 pipe.
#include <cstdio>
#include <cuda_runtime.h>
template <int NT>
__global__ void __launch_bounds__(128) k_chain(double* out, int iters, const double4* src, int nsrc) {
    __shared__ double2 tgt[4][16];
    __shared__ double4 ssrc[512];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) ssrc[k] = src[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 16) tgt[warp][lane] = make_double2(0.001 * lane + 0.01 * warp, 0.002 * lane + blockIdx.x * 1e-4);
    __syncwarp();
    double ax[NT], ay[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) ax[t] = ay[t] = 0;
    double4 s = ssrc[lane];
    for (int it = 0; it < iters; it++) {
        double4 nx = ssrc[((it + 1) * 32 + lane) & 511];
#pragma unroll
        for (int t = 0; t < NT; t++) {
            double2 p = tgt[warp][t];
            double dx = p.x - s.x, dy = p.y - s.y;
            double den = fma(dx, dx, fma(dy, dy, s.w));
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
            double e = fma(-den, r0, 1.0);
            double gr = s.z * r0;
            double w = fma(gr, e, gr);
            ax[t] = fma(-dy, w, ax[t]);
            ay[t] = fma(dx, w, ay[t]);
        }
        s = nx;
    }
    double r = 0;
#pragma unroll
    for (int t = 0; t < NT; t++) r += ax[t] + ay[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NT>
void run(double* out, const double4* src, int nsrc) {
    for (int bpsm = 1; bpsm <= 4; bpsm++) {
        int blocks = 148 * bpsm, iters = 20000 / NT;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k_chain<NT><<<blocks, 128>>>(out, 16, src, nsrc);
        cudaEventRecord(a);
        k_chain<NT><<<blocks, 128>>>(out, iters, src, nsrc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double pairs = (double)blocks * 128 * iters * NT;
        printf("NT=%2d warps/SMSP=%d  %.1f Gpairs/s\n", NT, bpsm, pairs / ms / 1e6);
    }
}
int main() {
    double* out; double4* src; const int nsrc = 1 << 16;
    cudaMalloc(&out, 148 * 8 * 128 * 8); cudaMalloc(&src, nsrc * sizeof(double4));
    double4* h = new double4[nsrc];
    for (int i = 0; i < nsrc; i++) { h[i].x = 0.37 + 1e-5 * i; h[i].y = 1.0 + 0.001 * (i % 977); h[i].z = 1e-6; h[i].w = 1e-8; }
    cudaMemcpy(src, h, nsrc * sizeof(double4), cudaMemcpyHostToDevice);
    run<1>(out, src, nsrc); run<2>(out, src, nsrc); run<3>(out, src, nsrc); run<4>(out, src, nsrc); run<5>(out, src, nsrc);
    run<6>(out, src, nsrc); run<8>(out, src, nsrc); run<10>(out, src, nsrc); run<15>(out, src, nsrc);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
