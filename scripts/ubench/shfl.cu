// micro-benchmark: do warp shuffles consume the L1/shared data pipe (l1tex__data_pipe_lsu_wavefronts)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void shfl_kernel(double* out, int iters) {
    double v = threadIdx.x * 1.0 + blockIdx.x, acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        v = __shfl_down_sync(0xffffffffu, v, 1) + 1.0;
        acc += v;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void lds_kernel(double* out, int iters) {
    __shared__ double s[256];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double acc = 0.0;
    int j = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        acc += s[j];
        j = (j + 33) & 255;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    double* d;
    cudaMalloc(&d, 148 * 8 * 256 * 8);
    for (int r = 0; r < 2; ++r) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        shfl_kernel<<<148 * 8, 256>>>(d, 4096);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        // warp-instructions: 148*8*8 warps * 4096 iters * 2 SHFL.32
        printf("shfl: %.3f ms, %.2f SHFL.32 warp-instr / clk / SM (at 1.965 GHz)\n", ms, 148.0 * 8 * 8 * 4096 * 2 / (ms * 1e-3 * 1.965e9 * 148));
        cudaEventRecord(a);
        lds_kernel<<<148 * 8, 256>>>(d, 4096);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("lds64: %.3f ms, %.2f LDS.64 warp-instr / clk / SM\n", ms, 148.0 * 8 * 8 * 4096 / (ms * 1e-3 * 1.965e9 * 148));
    }
    return 0;
}
