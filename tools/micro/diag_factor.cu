// Micro-benchmark: 32x32 Cholesky of a diagonal block by one warp (variants), cycles per factorisation.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int D_LD = 33, LT_LD = 34;

// V0: blocked by 8 columns in registers + rank-8 trailing updates through shared memory (kernel code)
__device__ int factor_v0(double* __restrict__ D, double* __restrict__ LT, double* rdiag, int lane, double& mypiv)
{
    int bad = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < 32 && !bad; k0 += 8) {
        double x[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) x[kk] = D[lane * D_LD + k0 + kk];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = k0 + kk;
            const double dk = __shfl_sync(0xffffffffu, x[kk], k);
            if (!(dk > 0.0) || !(dk < 1.0e300)) { if (!bad) bad = k + 1; }
            const double rk = rsqrt(dk);
            const double lik = (lane == k) ? dk * rk : x[kk] * rk;
            if (lane == k) mypiv = dk;
            x[kk] = lik;
            if (lane == 0) rdiag[k] = rk;
#pragma unroll
            for (int k2 = kk + 1; k2 < 8; ++k2) x[k2] = fma(-lik, __shfl_sync(0xffffffffu, lik, k0 + k2), x[k2]);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const bool low = lane >= k0 + kk;
            D[lane * D_LD + k0 + kk] = low ? x[kk] : 0.0;
            LT[(k0 + kk) * LT_LD + lane] = low ? x[kk] : 0.0;
        }
        __syncwarp();
#pragma unroll 2
        for (int c = k0 + 8; c < 32; ++c) {
            double a = D[lane * D_LD + c];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) a = fma(-x[kk], LT[(k0 + kk) * LT_LD + c], a);
            D[lane * D_LD + c] = a;
        }
        __syncwarp();
    }
    return bad;
}

// fast reciprocal square root: float seed + two Newton steps in double (rel. error ~1e-16)
__device__ __forceinline__ double rsqrt_fast(double d)
{
    const float f = rsqrtf((float)d);
    double y = (double)f;
    double h = 0.5 * d;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    double e = fma(-d * y, y, 1.0);          // residual correction
    return fma(0.5 * y, e, y);
}

// V1: same as V0 with rsqrt_fast
__device__ int factor_v1(double* __restrict__ D, double* __restrict__ LT, double* rdiag, int lane, double& mypiv)
{
    int bad = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < 32; k0 += 8) {
        double x[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) x[kk] = D[lane * D_LD + k0 + kk];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = k0 + kk;
            const double dk = __shfl_sync(0xffffffffu, x[kk], k);
            if (!(dk > 0.0) || !(dk < 1.0e300)) { if (!bad) bad = k + 1; }
            const double rk = rsqrt_fast(dk);
            const double lik = (lane == k) ? dk * rk : x[kk] * rk;
            if (lane == k) mypiv = dk;
            x[kk] = lik;
            if (lane == 0) rdiag[k] = rk;
#pragma unroll
            for (int k2 = kk + 1; k2 < 8; ++k2) x[k2] = fma(-lik, __shfl_sync(0xffffffffu, lik, k0 + k2), x[k2]);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const bool low = lane >= k0 + kk;
            D[lane * D_LD + k0 + kk] = low ? x[kk] : 0.0;
            LT[(k0 + kk) * LT_LD + lane] = low ? x[kk] : 0.0;
        }
        __syncwarp();
#pragma unroll 4
        for (int c = k0 + 8; c < 32; ++c) {
            double a0 = D[lane * D_LD + c], a1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; kk += 2) {
                a0 = fma(-x[kk], LT[(k0 + kk) * LT_LD + c], a0);
                a1 = fma(-x[kk + 1], LT[(k0 + kk + 1) * LT_LD + c], a1);
            }
            D[lane * D_LD + c] = a0 + a1;
        }
        __syncwarp();
    }
    return bad;
}

template <int V>
__global__ void __launch_bounds__(256, 3) k(const double* A, double* out, long long* cyc, int reps)
{
    __shared__ double D[32 * D_LD], LT[32 * LT_LD], rdiag[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long tot = 0;
    double mypiv = 1.0;
    int bad = 0;
    for (int r = 0; r < reps; ++r) {
        if (warp == 0) {
            for (int c = 0; c < 32; ++c) D[lane * D_LD + c] = A[lane * 32 + c];
            __syncwarp();
            long long t0 = clock64();
            bad |= (V == 0) ? factor_v0(D, LT, rdiag, lane, mypiv) : factor_v1(D, LT, rdiag, lane, mypiv);
            __syncwarp();
            tot += clock64() - t0;
        }
        __syncthreads();      // other warps wait here, as in the real kernel
    }
    if (warp != 0) return;
    for (int c = 0; c < 32; ++c) out[lane * 32 + c] = D[lane * D_LD + c];
    if (lane == 0) { cyc[0] = tot / reps; cyc[1] = bad; }
    out[1024 + lane] = mypiv + rdiag[lane];
}

int main(int argc, char** argv)
{
    int NT_ = argc > 1 ? atoi(argv[1]) : 32, NB_ = argc > 2 ? atoi(argv[2]) : 1;
    printf("threads %d blocks %d\n", NT_, NB_);
    double h[1024], o[1100];
    // SPD test matrix: exp(-|i-j|^2 / 50) + 0.05 I
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) h[i * 32 + j] = exp(-(i - j) * (i - j) / 50.0) + (i == j ? 0.05 : 0.0);
    double *A, *out; long long* cyc;
    cudaMalloc(&A, sizeof(h)); cudaMalloc(&out, sizeof(o)); cudaMalloc(&cyc, 16);
    cudaMemcpy(A, h, sizeof(h), cudaMemcpyHostToDevice);
    long long c[2];
    for (int v = 0; v < 2; ++v) {
        if (v == 0) k<0><<<NB_, NT_>>>(A, out, cyc, 20); else k<1><<<NB_, NT_>>>(A, out, cyc, 20);
        cudaDeviceSynchronize();
        cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
        // check L L^T = A
        double err = 0;
        for (int i = 0; i < 32; ++i) for (int j = 0; j <= i; ++j) {
            double s = 0; for (int t = 0; t <= j; ++t) s += o[i * 32 + t] * o[j * 32 + t];
            err = fmax(err, fabs(s - h[i * 32 + j]));
        }
        printf("variant %d: %lld cycles per 32x32 factorisation, bad=%lld, max |LL^T - A| = %.3e  (%s)\n", v, c[0], c[1], err, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
