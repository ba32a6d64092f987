// fp64_peak.cu -- measured FP64 issue peaks on this B200: DFMA (vector pipe) and DMMA m8n8k4 (tensor pipe).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double s) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], s, 1e-9);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void dmma_kernel(double* out, int iters, double s) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    double a = s, b = 1e-3 * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 16; warps *= 2) {
        int threads = 128, blocks = sms * warps / 4;
        int iters = 20000;
        for (int k = 0; k < 2; ++k) {
            float ms;
            dfma_kernel<<<blocks, threads>>>(out, 100, 0.999);
            cudaEventRecord(e0);
            dfma_kernel<<<blocks, threads>>>(out, iters, 0.999);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)blocks * threads * iters * 16;
            if (k) printf("DFMA  warps/SM=%2d: %.2f TFLOP/s (%.1f FMA/clk/SM @1965MHz)\n", warps, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
            dmma_kernel<<<blocks, threads>>>(out, 100, 0.999);
            cudaEventRecord(e0);
            dmma_kernel<<<blocks, threads>>>(out, iters, 0.999);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            fma = (double)blocks * (threads / 32) * iters * 8 * 256;
            if (k) printf("DMMA  warps/SM=%2d: %.2f TFLOP/s (%.1f FMA/clk/SM @1965MHz)\n", warps, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
