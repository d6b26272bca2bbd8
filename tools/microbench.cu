// Scratch micro-benchmarks for the K4 pair body: how many pairs/s can one SM sustain for this
// instruction mix, with and without MUFU.RCP64H, from registers and from shared-memory broadcast.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_pairs(double* out, int iters, const double2* src) {
    __shared__ double2 sxy[1024];
    __shared__ double2 sab[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) { sxy[k] = src[k]; sab[k] = src[1024 + k]; }
    __syncthreads();
    double tx = threadIdx.x * 1e-3, ty = blockIdx.x * 1e-3, rx = 0, ry = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll 4
        for (int k = 0; k < 1024; k++) {
            double2 p = sxy[k], q = sab[k];
            double dx = tx - p.x, dy = ty - p.y;
            double den = fma(dx, dx, fma(dy, dy, q.y));
            double w;
            if (MODE == 0) {  // rcp.approx + cubic correction
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
                double e = fma(-den, r0, 1.0);
                double e2 = fma(e, e, e);
                double gr = q.x * r0;
                w = fma(gr, e2, gr);
            } else if (MODE == 1) {  // IEEE division
                w = q.x / den;
            } else if (MODE == 2) {  // no reciprocal at all (pure FP64 pipe reference)
                double e = fma(-den, den, 1.0);
                double e2 = fma(e, e, e);
                double gr = q.x * den;
                w = fma(gr, e2, gr);
            } else {  // fp32 seed: rcp.approx.f32 + 2 Newton steps in fp64
                float rf;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)den));
                double r0 = (double)rf;
                double e = fma(-den, r0, 1.0);
                r0 = fma(r0, e, r0);
                e = fma(-den, r0, 1.0);
                double e2 = fma(e, e, e);
                double gr = q.x * r0;
                w = fma(gr, e2, gr);
            }
            rx = fma(-dy, w, rx);
            ry = fma(dx, w, ry);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = rx + ry;
}

template <int MODE>
void run(const char* name, double* out, const double2* src, int blocks, int threads) {
    int iters = 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_pairs<MODE><<<blocks, threads>>>(out, 1, src);
    cudaEventRecord(a);
    k_pairs<MODE><<<blocks, threads>>>(out, iters, src);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double pairs = (double)blocks * threads * iters * 1024;
    printf("%-28s blocks=%d threads=%d  %.3f ms  %.1f Gpairs/s\n", name, blocks, threads, ms, pairs / ms / 1e6);
}

int main() {
    double* out; double2* src;
    cudaMalloc(&out, 148 * 16 * 1024 * 8);
    cudaMalloc(&src, 2048 * 16);
    double2 h[2048];
    for (int i = 0; i < 2048; i++) { h[i].x = 0.37 * i + 0.1; h[i].y = 1.0 + 0.001 * i; }
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int threads : {256, 512}) {
        int blocks = 148 * (2048 / threads);
        run<0>("rcp.approx.f64+cubic", out, src, blocks, threads);
        run<1>("ieee div", out, src, blocks, threads);
        run<2>("no reciprocal (fp64 only)", out, src, blocks, threads);
        run<3>("rcp.f32 seed + newton", out, src, blocks, threads);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
