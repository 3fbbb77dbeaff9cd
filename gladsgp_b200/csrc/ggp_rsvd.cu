// Streaming passes of the randomized SVD over the (m x n) float32 ensemble.
//
// Replaces the three products of /root/reference/src/svd.py:52-60
//   Y = X @ omega ;  Y = X @ X.T @ Y ;  B = Q.T @ X
// as two kernels, each one pass over X (HBM-read bound, 4*m*n bytes):
//   sketch : Y[m][r]  = X * Omega      (Omega given transposed, [r][n])
//   xty    : Bt[r][n] = Y^T * X
// The power iteration X X^T Y is evaluated as X (X^T Y) = sketch(xty(Y)) -- the same product with
// 2*(2 m n r) instead of 2 m^2 n flops (the reference forms the m x m Gram matrix first,
// src/svd.py:56); results agree to FP32 round-off.  The small dense steps (QR of m x r, the
// r x r eigen-problem) stay in torch/cuSOLVER on the device.
// Lanes own columns (coalesced 128-byte row segments, alignment free: n is odd in the reference,
// 3693*365), partial sums across CTAs are combined in a fixed order in FP64 (deterministic).
#include "ggp_common.cuh"
#include "../../include/gladsgp_b200.h"

namespace ggp {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CPL = 4;                 // columns per lane
constexpr int RS_CHUNK = 32 * RS_CPL;     // columns per warp chunk

// ---- Y = X * Omega : CTA owns column chunks, accumulates Y in shared memory ---------------------
template <int RG>
__global__ void __launch_bounds__(RS_THREADS, 1)
sketch_kernel(const float* __restrict__ X, int m, long long n, const float* __restrict__ OmT, int r, int k0,
              float* __restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ys = reinterpret_cast<float*>(smem_raw);          // [m][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < m * 32; idx += RS_THREADS) Ys[idx] = 0.f;
    __syncthreads();
    const long long nchunk = (n + RS_CHUNK - 1) / RS_CHUNK;
    for (long long ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
        const long long c0 = ch * RS_CHUNK;
        float om[RS_CPL][RG];
        bool inb[RS_CPL];
#pragma unroll
        for (int i = 0; i < RS_CPL; ++i) {
            const long long c = c0 + lane + 32 * i;
            inb[i] = c < n;
#pragma unroll
            for (int k = 0; k < RG; ++k)
                om[i][k] = (inb[i] && k0 + k < r) ? OmT[(size_t)(k0 + k) * n + c] : 0.f;
        }
        for (int row = warp; row < m; row += RS_WARPS) {
            const float* xr = X + (size_t)row * n + c0 + lane;
            float x[RS_CPL];
#pragma unroll
            for (int i = 0; i < RS_CPL; ++i) x[i] = inb[i] ? __ldcs(xr + 32 * i) : 0.f;
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < RG) {
                    float s = x[0] * om[0][k];
#pragma unroll
                    for (int i = 1; i < RS_CPL; ++i) s = fmaf(x[i], om[i][k], s);
                    v[k] = s;
                } else {
                    v[k] = 0.f;
                }
            }
            // transpose-reduce: after the 5 rounds lane l holds sum over lanes of v[l]
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const bool up = (lane & step) != 0;
#pragma unroll
                for (int j = 0; j < step; ++j) {
                    const float send = up ? v[j] : v[j + step];
                    const float keep = up ? v[j + step] : v[j];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                }
            }
            Ys[row * 32 + lane] += v[0];       // rows are warp-private: no race
        }
    }
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x * m * 32;
    for (int idx = tid; idx < m * 32; idx += RS_THREADS) out[idx] = Ys[idx];
}

__global__ void sketch_reduce_kernel(const float* __restrict__ partial, int nparts, int m, int r, int k0,
                                     float* __restrict__ Y)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * 32) return;
    const int row = idx >> 5, k = idx & 31;
    if (k0 + k >= r) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * m * 32 + idx];
    Y[(size_t)row * r + k0 + k] = (float)s;
}

// ---- Bt = Y^T X : lanes own columns, loop over all rows, Y broadcast from shared memory ----------
template <int RG>
__global__ void __launch_bounds__(RS_THREADS, 1)
xty_kernel(const float* __restrict__ X, int m, long long n, const float* __restrict__ Y, int r, int k0,
           float* __restrict__ Bt)
{
    constexpr int RGP = (RG + 3) & ~3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ys = reinterpret_cast<float*>(smem_raw);          // [m][RGP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < m * RGP; idx += RS_THREADS) {
        const int row = idx / RGP, k = idx - row * RGP;
        Ys[idx] = (k < RG && k0 + k < r) ? Y[(size_t)row * r + k0 + k] : 0.f;
    }
    __syncthreads();
    const long long nchunk = (n + RS_CHUNK - 1) / RS_CHUNK;
    for (long long ch = (long long)blockIdx.x * RS_WARPS + warp; ch < nchunk; ch += (long long)gridDim.x * RS_WARPS) {
        const long long c0 = ch * RS_CHUNK;
        bool inb[RS_CPL];
#pragma unroll
        for (int i = 0; i < RS_CPL; ++i) inb[i] = c0 + lane + 32 * i < n;
        float acc[RS_CPL][RGP];
#pragma unroll
        for (int i = 0; i < RS_CPL; ++i)
#pragma unroll
            for (int k = 0; k < RGP; ++k) acc[i][k] = 0.f;
        const float* xc = X + c0 + lane;
#pragma unroll 2
        for (int row = 0; row < m; ++row) {
            const float* xr = xc + (size_t)row * n;
            float x[RS_CPL];
#pragma unroll
            for (int i = 0; i < RS_CPL; ++i) x[i] = inb[i] ? __ldcs(xr + 32 * i) : 0.f;
#pragma unroll
            for (int k4 = 0; k4 < RGP; k4 += 4) {
                const float4 y4 = *reinterpret_cast<const float4*>(Ys + row * RGP + k4);
#pragma unroll
                for (int i = 0; i < RS_CPL; ++i) {
                    acc[i][k4 + 0] = fmaf(x[i], y4.x, acc[i][k4 + 0]);
                    acc[i][k4 + 1] = fmaf(x[i], y4.y, acc[i][k4 + 1]);
                    acc[i][k4 + 2] = fmaf(x[i], y4.z, acc[i][k4 + 2]);
                    acc[i][k4 + 3] = fmaf(x[i], y4.w, acc[i][k4 + 3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            if (k0 + k < r) {
#pragma unroll
                for (int i = 0; i < RS_CPL; ++i)
                    if (inb[i]) Bt[(size_t)(k0 + k) * n + c0 + lane + 32 * i] = acc[i][k];
            }
        }
    }
}

static int sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace ggp

using namespace ggp;

extern "C" {

long long ggp_rsvd_workspace_bytes(int m)
{
    if (m <= 0) return -1;
    return (long long)sm_count() * m * 32 * (long long)sizeof(float);
}

int ggp_rsvd_sketch_f32(const float* X, int m, long long n, const float* OmegaT, int r, float* Y_out,
                        void* workspace, long long workspace_bytes, void* stream)
{
    GGP_ARG(X && OmegaT && Y_out && workspace, "null pointer");
    GGP_ARG(m > 0 && n > 0 && r > 0, "m, n, r must be positive");
    const size_t smem = (size_t)m * 32 * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("ggp_rsvd_sketch_f32: m=%d too large (needs %zu B shared memory)", m, smem);
        return GGP_ERR_UNSUPPORTED;
    }
    const int grid = sm_count();
    if (workspace_bytes < ggp_rsvd_workspace_bytes(m)) {
        set_error("ggp_rsvd_sketch_f32: workspace too small");
        return GGP_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    for (int k0 = 0; k0 < r; k0 += 32) {
        if (r - k0 <= 25) {
            GGP_CUDA(cudaFuncSetAttribute(sketch_kernel<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sketch_kernel<25><<<grid, RS_THREADS, smem, st>>>(X, m, n, OmegaT, r, k0, partial);
        } else {
            GGP_CUDA(cudaFuncSetAttribute(sketch_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sketch_kernel<32><<<grid, RS_THREADS, smem, st>>>(X, m, n, OmegaT, r, k0, partial);
        }
        sketch_reduce_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(partial, grid, m, r, k0, Y_out);
    }
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

int ggp_rsvd_xty_f32(const float* X, int m, long long n, const float* Y, int r, float* Bt_out, void* stream)
{
    GGP_ARG(X && Y && Bt_out, "null pointer");
    GGP_ARG(m > 0 && n > 0 && r > 0, "m, n, r must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nchunk = (n + RS_CHUNK - 1) / RS_CHUNK;
    long long want = (nchunk + RS_WARPS - 1) / RS_WARPS;
    const int grid = (int)(want < 4LL * sm_count() ? want : 4LL * sm_count());
    for (int k0 = 0; k0 < r; k0 += 32) {
        if (r - k0 <= 25) {
            const size_t smem = (size_t)m * 28 * sizeof(float);
            GGP_ARG(smem <= 200 * 1024, "m too large for xty");
            GGP_CUDA(cudaFuncSetAttribute(xty_kernel<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            xty_kernel<25><<<grid, RS_THREADS, smem, st>>>(X, m, n, Y, r, k0, Bt_out);
        } else {
            const size_t smem = (size_t)m * 32 * sizeof(float);
            GGP_ARG(smem <= 200 * 1024, "m too large for xty");
            GGP_CUDA(cudaFuncSetAttribute(xty_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            xty_kernel<32><<<grid, RS_THREADS, smem, st>>>(X, m, n, Y, r, k0, Bt_out);
        }
    }
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

}  // extern "C"
