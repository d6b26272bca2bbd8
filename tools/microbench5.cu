// Scratch: cost of cooperative_groups grid.sync() for the tree-build launch shape (148 CTAs x 1024 / 512 / 256 threads)
// and of a hand-rolled sense-reversing barrier (one atomic per CTA + spin on a flag).
#include <cstdio>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void k_cg(int iters, int* out) {
    cg::grid_group grid = cg::this_grid();
    for (int i = 0; i < iters; i++) grid.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = iters;
}
__device__ __forceinline__ void my_sync(unsigned* count, volatile unsigned* gen, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(count, 1) == nblocks - 1) {
            *count = 0;
            __threadfence();
            *gen = g + 1;
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}
__global__ void k_my(int iters, unsigned* count, unsigned* gen, int* out) {
    for (int i = 0; i < iters; i++) my_sync(count, gen, gridDim.x);
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = iters;
}
int main() {
    int* out; unsigned* bar; cudaMalloc(&out, 4); cudaMalloc(&bar, 8); cudaMemset(bar, 0, 8);
    for (int threads : {1024, 512, 256}) {
        int iters = 2000;
        void* args[] = {&iters, &out};
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaLaunchCooperativeKernel((void*)k_cg, dim3(148), dim3(threads), args, 0, 0);
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void*)k_cg, dim3(148), dim3(threads), args, 0, 0);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("cg grid.sync  148 x %4d: %.2f us per sync (%s)\n", threads, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        unsigned* cnt = bar; unsigned* gen = bar + 1;
        void* args2[] = {&iters, &cnt, &gen, &out};
        cudaLaunchCooperativeKernel((void*)k_my, dim3(148), dim3(threads), args2, 0, 0);
        cudaEventRecord(a);
        cudaLaunchCooperativeKernel((void*)k_my, dim3(148), dim3(threads), args2, 0, 0);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("hand barrier  148 x %4d: %.2f us per sync (%s)\n", threads, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
