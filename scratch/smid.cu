#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(int* out) {
    extern __shared__ char sm[];
    unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s));
    if (threadIdx.x == 0) out[blockIdx.x] = s;
    unsigned long long t0, t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 20000);
}
int main() {
    int n = 600; int* d; cudaMalloc(&d, n * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<n, 256, 82000>>>(d);
    int h[600]; cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 320; ++i) printf("%d%c", h[i], (i % 37 == 36) ? '\n' : ' ');
    printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
}
