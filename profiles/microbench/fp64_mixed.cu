// fp64_mixed.cu -- are DMMA (m8n8k4.f64) and DFMA served by the same FP64 datapath on B200?
// Three kernels with the same total warps per SM: DMMA only, DFMA only, and a mix -- (a) alternate warps take one or the
// other, (b) every warp interleaves both streams.  If the pipes were separate the mixed FMA rate would approach the sum.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mixed fp64_mixed.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define DMMA(c0, c1, a, b) \
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

// mode 0: DMMA only; 1: DFMA only; 2: odd warps DFMA, even warps DMMA; 3: every warp both, interleaved
// per iteration a DMMA warp issues 8 DMMA (2048 FMA), a DFMA warp 64 DFMA (2048 FMA); an interleaving warp issues both
template <int MODE>
__global__ void mixed(double* out, int iters, double s) {
    double c[8][2], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    const double a = s, b = 1e-3 * (threadIdx.x & 3);
    const int warp = threadIdx.x >> 5;
    const bool do_mma = MODE == 0 || MODE == 3 || (MODE == 2 && !(warp & 1));
    const bool do_fma = MODE == 1 || MODE == 3 || (MODE == 2 && (warp & 1));
    for (int it = 0; it < iters; ++it) {
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                DMMA(c[i][0], c[i][1], a, b);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[(i * 8 + k) & 15] = fma(f[(i * 8 + k) & 15], s, 1e-9);
            }
        } else {
            if (do_mma) {
#pragma unroll
                for (int i = 0; i < 8; ++i) DMMA(c[i][0], c[i][1], a, b);
            }
            if (do_fma) {
#pragma unroll
                for (int k = 0; k < 64; ++k) f[k & 15] = fma(f[k & 15], s, 1e-9);
            }
        }
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) t += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
static void run(const char* name, double* out, int sms, int warps, double fma_per_warp_iter_avg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = warps * 32, iters = 20000;
    float ms;
    mixed<MODE><<<sms, threads>>>(out, 100, 0.999);
    cudaEventRecord(e0);
    mixed<MODE><<<sms, threads>>>(out, iters, 0.999);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * warps * iters * fma_per_warp_iter_avg;
    printf("%-28s warps/SM=%2d: %7.2f TFLOP/s (%.1f FMA/clk/SM @1965MHz)  %.3f ms\n", name, warps, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / 1.965e9, ms);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int warps = 8; warps <= 16; warps *= 2) {   // (32 warps of this kernel do not fit the register file)
        run<0>("DMMA only", out, sms, warps, 2048.0);
        run<1>("DFMA only", out, sms, warps, 2048.0);
        run<2>("alternate warps DMMA/DFMA", out, sms, warps, 2048.0);
        run<3>("every warp DMMA+DFMA", out, sms, warps, 4096.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
