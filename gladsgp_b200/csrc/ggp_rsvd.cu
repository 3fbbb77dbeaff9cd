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

// ---- Y = X * Omega : the CTA streams 32-column chunks of all m rows through a padded shared-memory tile
// (coalesced 128-byte row segments in, conflict-free column reads out); each thread owns RPT rows and keeps
// their RG partial outputs in registers, Omega rows are broadcast LDS.128.  FMA-bound at ~20 B/clk/SM.
constexpr int SK_CK = 32;

template <int RG, int RPT>
__global__ void __launch_bounds__(RS_THREADS, 2)
sketch_kernel(const float* __restrict__ X, int m, long long n, const float* __restrict__ OmT, int r, int k0,
              float* __restrict__ partial)
{
    constexpr int RGP = (RG + 3) & ~3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Xs = reinterpret_cast<float*>(smem_raw);                 // [RPT*256][33]
    float* Os = Xs + (size_t)RPT * RS_THREADS * 33;                 // [32][RGP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc[RPT][RGP];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int k = 0; k < RGP; ++k) acc[i][k] = 0.f;
    const int mrows = RPT * RS_THREADS;                             // padded row count handled by the CTA
    const long long nchunk = (n + SK_CK - 1) / SK_CK;
    for (long long ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
        const long long c0 = ch * SK_CK;
        const bool cin = c0 + lane < n;
        __syncthreads();                                            // previous chunk consumed
        // tile load: warp w takes rows w, w+8, ...; 16 loads in flight per thread
        for (int rb = warp; rb < mrows; rb += RS_WARPS * 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int row = rb + RS_WARPS * u;
                v[u] = (row < m && cin) ? __ldcs(X + (size_t)row * n + c0 + lane) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int row = rb + RS_WARPS * u;
                if (row < mrows) Xs[row * 33 + lane] = v[u];
            }
        }
        for (int idx = tid; idx < RGP * 32; idx += RS_THREADS) {
            const int k = idx >> 5, c = idx & 31;
            Os[c * RGP + k] = (k < RG && k0 + k < r && c0 + c < n) ? OmT[(size_t)(k0 + k) * n + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < SK_CK; ++c) {
            float x[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) x[i] = Xs[(tid + RS_THREADS * i) * 33 + c];
#pragma unroll
            for (int k4 = 0; k4 < RGP; k4 += 4) {
                const float4 o4 = *reinterpret_cast<const float4*>(Os + c * RGP + k4);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    acc[i][k4 + 0] = fmaf(x[i], o4.x, acc[i][k4 + 0]);
                    acc[i][k4 + 1] = fmaf(x[i], o4.y, acc[i][k4 + 1]);
                    acc[i][k4 + 2] = fmaf(x[i], o4.z, acc[i][k4 + 2]);
                    acc[i][k4 + 3] = fmaf(x[i], o4.w, acc[i][k4 + 3]);
                }
            }
        }
    }
    float* out = partial + (size_t)blockIdx.x * m * 32;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int row = tid + RS_THREADS * i;
        if (row < m) {
#pragma unroll
            for (int k = 0; k < 32; ++k) out[row * 32 + k] = (k < RG) ? acc[i][k < RGP ? k : 0] : 0.f;
        }
    }
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
#pragma unroll 8
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

template <int RG, int RPT>
static int launch_sketch(const float* X, int m, long long n, const float* OmT, int r, int k0, float* partial, int grid,
                         cudaStream_t st)
{
    constexpr int RGP = (RG + 3) & ~3;
    const size_t smem = ((size_t)RPT * RS_THREADS * 33 + 32 * RGP) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(sketch_kernel<RG, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(sketch_kernel)");
    sketch_kernel<RG, RPT><<<grid, RS_THREADS, smem, st>>>(X, m, n, OmT, r, k0, partial);
    return GGP_OK;
}

extern "C" {

long long ggp_rsvd_workspace_bytes(int m)
{
    if (m <= 0) return -1;
    return 2LL * sm_count() * m * 32 * (long long)sizeof(float);
}

int ggp_rsvd_sketch_f32(const float* X, int m, long long n, const float* OmegaT, int r, float* Y_out,
                        void* workspace, long long workspace_bytes, void* stream)
{
    GGP_ARG(X && OmegaT && Y_out && workspace, "null pointer");
    GGP_ARG(m > 0 && n > 0 && r > 0, "m, n, r must be positive");
    if (m > 4 * RS_THREADS) {
        set_error("ggp_rsvd_sketch_f32: m=%d > %d not supported", m, 4 * RS_THREADS);
        return GGP_ERR_UNSUPPORTED;
    }
    const int grid = 2 * sm_count();
    if (workspace_bytes < ggp_rsvd_workspace_bytes(m)) {
        set_error("ggp_rsvd_sketch_f32: workspace too small");
        return GGP_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    const int rpt = (m + RS_THREADS - 1) / RS_THREADS;
    for (int k0 = 0; k0 < r; k0 += 32) {
        const bool small = (r - k0 <= 25);
        int rc;
        if (rpt <= 1) rc = small ? launch_sketch<25, 1>(X, m, n, OmegaT, r, k0, partial, grid, st)
                                 : launch_sketch<32, 1>(X, m, n, OmegaT, r, k0, partial, grid, st);
        else if (rpt == 2) rc = small ? launch_sketch<25, 2>(X, m, n, OmegaT, r, k0, partial, grid, st)
                                      : launch_sketch<32, 2>(X, m, n, OmegaT, r, k0, partial, grid, st);
        else rc = small ? launch_sketch<25, 4>(X, m, n, OmegaT, r, k0, partial, grid, st)
                        : launch_sketch<32, 4>(X, m, n, OmegaT, r, k0, partial, grid, st);
        if (rc != GGP_OK) return rc;
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
