// dfma_latency.cu -- dependent-chain latency of DFMA / DMUL on this B200's FP64 pipe (1 warp per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS> __global__ void k(double* out, int iters, double s) {
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], s, 1e-9);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) t += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int CHAINS> void run(double* out, int sms, int warps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 100000; float ms;
    k<CHAINS><<<sms, warps * 32>>>(out, 100, 0.999);
    cudaEventRecord(e0); k<CHAINS><<<sms, warps * 32>>>(out, iters, 0.999); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("chains=%d warps/SM=%2d: %.1f clk per dependent step (%.2f clk per DFMA issued per warp)\n", CHAINS, warps,
           ms * 1e-3 * 1.965e9 / iters, ms * 1e-3 * 1.965e9 / iters / CHAINS);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    run<1>(out, p.multiProcessorCount, 4);
    run<2>(out, p.multiProcessorCount, 4);
    run<4>(out, p.multiProcessorCount, 4);
    run<8>(out, p.multiProcessorCount, 4);
    run<1>(out, p.multiProcessorCount, 16);
    run<4>(out, p.multiProcessorCount, 16);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
