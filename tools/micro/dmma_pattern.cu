// DMMA.8x8x4 throughput with realistic operand patterns (distinct A/B registers, dependent pairs).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: 4x4 tiles, order (cb, i, {x,y}) dependent pairs adjacent (as in the kernel)
// MODE 1: 4x4 tiles, all .x first then all .y (dependent DMMAs 16 apart)
// MODE 2: same as 1 but operands refreshed from shared memory every iteration
template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k(double* out, const double* in, int iters)
{
    __shared__ double sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = in[i];
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    double2 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<double2*>(sh + 2 * threadIdx.x % 1024 + 8 * i);
        b[i] = *reinterpret_cast<double2*>(sh + 1024 + 2 * threadIdx.x % 512 + 8 * i);
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = *reinterpret_cast<double2*>(sh + (2 * threadIdx.x + 64 * it) % 1024 + 8 * i);
                b[i] = *reinterpret_cast<double2*>(sh + 1024 + (2 * threadIdx.x + 32 * it) % 512 + 8 * i);
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dmma(acc[i][cb][0], acc[i][cb][1], a[i].x, b[cb].x);
                    dmma(acc[i][cb][0], acc[i][cb][1], a[i].y, b[cb].y);
                }
        } else {
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(acc[i][cb][0], acc[i][cb][1], a[i].x, b[cb].x);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(acc[i][cb][0], acc[i][cb][1], a[i].y, b[cb].y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int NW>
void run(const char* name)
{
    double *out, *in; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 2048 * 8); cudaMemset(in, 0, 2048 * 8);
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NW><<<148, NW * 32>>>(out, in, 100);
    cudaEventRecord(e0);
    k<MODE, NW><<<148, NW * 32>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 32 * NW * 148 * (double)iters;
    printf("%-44s warps=%2d %.3f ms  %.2f TF\n", name, NW, ms, flops / ms / 1e9);
}
int main()
{
    run<0, 16>("dependent pairs adjacent"); run<0, 8>("dependent pairs adjacent"); run<0, 4>("dependent pairs adjacent");
    run<1, 16>("x pass then y pass"); run<1, 8>("x pass then y pass"); run<1, 4>("x pass then y pass");
    run<2, 16>("x/y passes + smem operand reload"); run<2, 4>("x/y passes + smem operand reload");
    return 0;
}
