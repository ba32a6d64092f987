// dmma_pattern.cu -- throughput of the exact DMMA sequences the pass kernel issues (two chained m8n8k4.f64 per
// register pair, results in place), without any data movement around them.  Peak is 64 FMA/clk/SM (fp64_peak.cu).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NR = 32;
typedef double Regs[NR];
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int X> __device__ __forceinline__ void m_u2(Regs& a, double b0, double b1) {
    if (X == 0) {
#pragma unroll
        for (int p = 0; p < NR; p += 4) {
            double t0, t1, u0, u1;
            dmma884(t0, t1, a[p], b0, 0.0, 0.0);
            dmma884(u0, u1, a[p + 2], b0, 0.0, 0.0);
            dmma884(a[p], a[p + 1], a[p + 1], b1, t0, t1);
            dmma884(a[p + 2], a[p + 3], a[p + 3], b1, u0, u1);
        }
    } else {
#pragma unroll
        for (int o = 0; o < NR / 4; ++o) {
            const int lo = o & ((1 << (X - 1)) - 1), hi = o >> (X - 1);
            const int p00 = (hi << (X + 1)) | (lo << 1), p01 = p00 | 1, p10 = p00 | (1 << X), p11 = p10 | 1;
            double t0, t1, u0, u1;
            dmma884(t0, t1, a[p00], b0, 0.0, 0.0);
            dmma884(u0, u1, a[p01], b0, 0.0, 0.0);
            dmma884(a[p00], a[p01], a[p10], b1, t0, t1);
            dmma884(a[p10], a[p11], a[p11], b1, u0, u1);
        }
    }
}
// variant with 4 chains in flight
template <int X> __device__ __forceinline__ void m_u2_deep(Regs& a, double b0, double b1) {
#pragma unroll
    for (int p = 0; p < NR; p += 8) {
        double t[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma884(t[i][0], t[i][1], a[p + 2 * i], b0, 0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma884(a[p + 2 * i], a[p + 2 * i + 1], a[p + 2 * i + 1], b1, t[i][0], t[i][1]);
    }
}
template <int MODE> __global__ void __launch_bounds__(256, 2) k(double* out, int iters, double s) {
    if (MODE == 3) {  // latency: one dependent chain
        double c0 = threadIdx.x, c1 = 1.0, b = 1e-3 * (threadIdx.x & 3);
        for (int it = 0; it < iters * 64; ++it) dmma884(c0, c1, c0, b, c0, c1);
        out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
        return;
    }
    Regs a;
#pragma unroll
    for (int i = 0; i < NR; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b0 = s, b1 = 1e-3 * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) m_u2<0>(a, b0, b1);
        else if (MODE == 1) m_u2<2>(a, b0, b1);
        else m_u2_deep<0>(a, b0, b1);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < NR; ++i) t += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 2 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int ctas = warps > 8 ? 2 : 1, threads = warps > 8 ? 256 : warps * 32;
            int iters = 4000; float ms;
            auto launch = [&](int it) {
                if (mode == 0) k<0><<<sms * ctas, threads>>>(out, it, 0.5);
                else if (mode == 1) k<1><<<sms * ctas, threads>>>(out, it, 0.5);
                else if (mode == 2) k<2><<<sms * ctas, threads>>>(out, it, 0.5);
                else k<3><<<sms * ctas, threads>>>(out, it, 0.5);
            };
            launch(10);
            cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)sms * warps * iters * (mode == 3 ? 64 : 32) * 256;
            printf("mode %d (%s) warps/SM=%2d: %.1f FMA/clk/SM @1965MHz", mode,
                   mode == 0 ? "X=0, 2 chains" : mode == 1 ? "X=2 role swap, 2 chains" : mode == 2 ? "X=0, 4 chains" : "1 dependent chain", warps,
                   fma / (ms * 1e-3) / sms / 1.965e9);
            if (mode == 3) printf("  -> %.1f clk per dependent DMMA", ms * 1e-3 * 1.965e9 / (iters * 64.0));
            printf("\n");
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
