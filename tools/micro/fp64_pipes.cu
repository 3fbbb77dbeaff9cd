// Micro-benchmark: DFMA vs DMMA.8x8x4 throughput on sm_100a, alone and mixed (are the pipes independent?)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA; 3: every warp interleaves both
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, double a, double b)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const int warp = threadIdx.x >> 5;
    const bool do_fma = MODE == 0 || (MODE == 2 && (warp & 1) == 0);
    const bool do_mma = MODE == 1 || (MODE == 2 && (warp & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
#pragma unroll
            for (int i = 8; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
        } else if (do_fma) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        } else if (do_mma) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double fma_per_thread_iter, double mma_per_warp_iter)
{
    double* out; cudaMalloc(&out, 148 * 512 * 8);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(out, 1000, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 512 threads
    double fma_flops = 2.0 * fma_per_thread_iter * 512 * 148 * iters;
    double mma_flops = 2.0 * 256 * mma_per_warp_iter * 16 * 148 * iters;
    printf("%-28s %.3f ms  DFMA %.2f TF  DMMA %.2f TF  total %.2f TF\n", name, ms, fma_flops / ms / 1e9, mma_flops / ms / 1e9,
           (fma_flops + mma_flops) / ms / 1e9);
    cudaFree(out);
}
int main()
{
    run<0>("all warps DFMA", 16, 0);
    run<1>("all warps DMMA", 0, 8);
    run<2>("even DFMA / odd DMMA", 8, 4);     // averaged over warps: half the warps do 16 fma, half do 8 mma
    run<3>("interleaved 16 DFMA + 4 DMMA", 16, 4);
    return 0;
}
