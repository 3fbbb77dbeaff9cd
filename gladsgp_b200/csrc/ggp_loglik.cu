// Covariance build, cross-covariance and batched fused log-likelihood entry points.
// C-ABI declared in include/gladsgp_b200.h.
#include <cstdlib>
#include "ggp_chol.cuh"
#include "../../include/gladsgp_b200.h"

namespace ggp {

// ---------------------------------------------------------------------------------------------
// (1) product squared-exponential covariance, materialised (SepiaDistCov type 1).
// 8*m*m bytes written per matrix.  One CTA per (64x64 tile pair, matrix): the tile (bi >= bj) is
// computed once (half the exps; 4x4 distances per thread in registers, in-kernel exp) and written
// twice -- directly and transposed through shared memory, both as 16-byte streaming stores.
// Measured 2.8 TB/s: the double-precision exponential, not HBM, is the bound (profiles/r1_cov_build_summary.txt).
// ---------------------------------------------------------------------------------------------
constexpr int CT = 64;

__global__ void __launch_bounds__(256, 4)
cov_build_kernel(const double* __restrict__ X, int m, int d, const double* __restrict__ beta,
                 const double* __restrict__ lamz, const double* __restrict__ diag_add,
                 double* __restrict__ C, int ntile)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Sr = reinterpret_cast<double*>(smem_raw);      // [d][CT] scaled row coordinates (transposed)
    double* Sc = Sr + CT * d;                              // [d][CT] scaled column coordinates
    double* T = Sc + CT * d;                               // [CT][CT+1] transposed tile
    double* etab = T + CT * (CT + 1);                      // [32]
    const int b = blockIdx.y;
    // decode lower-triangular tile index
    int t = blockIdx.x;
    int bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - bi * (bi + 1) / 2;
    const double* be = beta + (size_t)b * d;
    const double il = 1.0 / lamz[b];
    const double dg = il + diag_add[b];
    const int r0 = bi * CT, c0 = bj * CT;
    fill_exp_table(etab);
    for (int idx = threadIdx.x; idx < CT * d; idx += blockDim.x) {
        const int k = idx / CT, r = idx - k * CT;
        const double sb = sqrt(be[k]);
        Sr[idx] = (r0 + r < m) ? X[(size_t)(r0 + r) * d + k] * sb : 0.0;
        Sc[idx] = (c0 + r < m) ? X[(size_t)(c0 + r) * d + k] * sb : 0.0;
    }
    __syncthreads();
    double* Cb = C + (size_t)b * m * m;
    const bool vec2 = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    // each thread: 4 rows (ty + 16 i) x 4 consecutive columns (4 tx + j), distances in registers
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double dist[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dist[i][j] = 0.0;
    for (int k = 0; k < d; ++k) {
        double rv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[i] = Sr[k * CT + ty + 16 * i];
        const double2 c01 = *reinterpret_cast<const double2*>(Sc + k * CT + 4 * tx);
        const double2 c23 = *reinterpret_cast<const double2*>(Sc + k * CT + 4 * tx + 2);
        const double cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double df = rv[i] - cv[j];
                dist[i][j] = fma(df, df, dist[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int lr = ty + 16 * i;
        const int r = r0 + lr;
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int lc = 4 * tx + j;
            v[j] = (r == c0 + lc) ? dg : exp_neg(-dist[i][j], etab) * il;
            T[lc * (CT + 1) + lr] = v[j];
        }
        if (r < m) {
            const int c = c0 + 4 * tx;
            if (vec2 && c + 3 < m) {
                __stcs(reinterpret_cast<double2*>(Cb + (size_t)r * m + c), make_double2(v[0], v[1]));
                __stcs(reinterpret_cast<double2*>(Cb + (size_t)r * m + c + 2), make_double2(v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < m) Cb[(size_t)r * m + c + j] = v[j];
            }
        }
    }
    if (bi != bj) {
        __syncthreads();
        for (int rr = 0; rr < 4; ++rr) {
            const int lr = ty + 16 * rr;          // row of the transposed tile = column of the original
            const int r = c0 + lr;
            if (r < m) {
                const int c = r0 + 4 * tx;
                const double* tp = T + lr * (CT + 1) + 4 * tx;
                if (vec2 && c + 3 < m) {
                    __stcs(reinterpret_cast<double2*>(Cb + (size_t)r * m + c), make_double2(tp[0], tp[1]));
                    __stcs(reinterpret_cast<double2*>(Cb + (size_t)r * m + c + 2), make_double2(tp[2], tp[3]));
                } else {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        if (c + cc < m) Cb[(size_t)r * m + c + cc] = tp[cc];
                }
            }
        }
    }
}

// SepiaDistCov type 2: S21[b][i][t] = exp(-sum_k beta_k (x_ik - xp_tk)^2) / lamz, (m x n) row-major.
__global__ void __launch_bounds__(256)
cross_cov_kernel(const double* __restrict__ X, int m, const double* __restrict__ Xp, int n, int d,
                 const double* __restrict__ beta, const double* __restrict__ lamz,
                 double* __restrict__ S21)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Sr = reinterpret_cast<double*>(smem_raw);      // [CT][d]
    double* Sc = Sr + CT * d;                              // [CT][d]
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * CT, c0 = blockIdx.x * CT;
    const double* be = beta + (size_t)b * d;
    const double il = 1.0 / lamz[b];
    for (int idx = threadIdx.x; idx < CT * d; idx += blockDim.x) {
        int r = idx / d, k = idx - r * d;
        double sb = sqrt(be[k]);
        Sr[idx] = (r0 + r < m) ? X[(size_t)(r0 + r) * d + k] * sb : 0.0;
        Sc[idx] = (c0 + r < n) ? Xp[(size_t)(c0 + r) * d + k] * sb : 0.0;
    }
    __syncthreads();
    double* Sb = S21 + (size_t)b * m * n;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int rr = 0; rr < 4; ++rr) {
        const int lr = ty + 16 * rr;
        const int r = r0 + lr;
        if (r >= m) continue;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int lc = 4 * tx + cc;
            const int c = c0 + lc;
            if (c >= n) continue;
            double dist = 0.0;
            for (int k = 0; k < d; ++k) {
                double df = Sr[lr * d + k] - Sc[lc * d + k];
                dist = fma(df, df, dist);
            }
            Sb[(size_t)r * n + c] = exp(-dist) * il;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (2) batched fused log-likelihood: one CTA per matrix.
// ---------------------------------------------------------------------------------------------
template <bool CL, bool LA = false, int RA = GGP_RA>
__global__ void __launch_bounds__(NT, CL ? GGP_CL_CTAS_PER_SM : GGP_CTAS_PER_SM)
loglik_batched_kernel(const double* __restrict__ X, int m, int Mp, int d, const double* __restrict__ W,
                      long long w_stride, const double* __restrict__ beta, const double* __restrict__ lamz,
                      const double* __restrict__ diag_add, double* __restrict__ Lws, long long l_stride,
                      double* __restrict__ u_out, double* __restrict__ loglik, int* __restrict__ info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = CL ? blockIdx.x / cluster_nctarank() : blockIdx.x;
    double ll;
    if constexpr (LA) {
        LaSmem sm = carve_la_smem(smem_raw, Mp, d);
        ll = eval_block_loglik_la(sm, X, m, Mp, d, beta + (size_t)b * d, lamz[b], diag_add[b],
                                  W + (size_t)b * w_stride, Lws + (size_t)b * l_stride,
                                  u_out ? u_out + (size_t)b * Mp : nullptr, info ? info + b : nullptr);
    } else {
        EvalSmem sm = carve_eval_smem(smem_raw, Mp, d);
        ll = eval_block_loglik<CL, RA>(sm, X, m, Mp, d, beta + (size_t)b * d, lamz[b], diag_add[b],
                                   W + (size_t)b * w_stride, Lws + (size_t)b * l_stride,
                                   u_out ? u_out + (size_t)b * Mp : nullptr, info ? info + b : nullptr);
    }
    if (threadIdx.x == 0 && (!CL || cluster_ctarank() == 0)) loglik[b] = ll;
}

// packed factor -> dense lower-triangular (m x m row-major), for tests / inspection
__global__ void unpack_factor_kernel(const double* __restrict__ Lws, long long l_stride, int m, int Mp,
                                     double* __restrict__ Ld)
{
    const int b = blockIdx.y;
    const double* Lp = Lws + (size_t)b * l_stride;
    double* out = Ld + (size_t)b * m * m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)m * m;
         idx += (long long)gridDim.x * blockDim.x) {
        int r = (int)(idx / m), c = (int)(idx - (long long)r * m);
        double v = 0.0;
        if (c <= r) {
            int kb = c >> 5, ks = (c >> 3) & 3, cc = c & 7;
            v = Lp[panel_off(kb, Mp) + (long long)ks * (Mp - 32 * kb) * 8 + (long long)(r - 32 * kb) * 8 + cc];
        }
        out[idx] = v;
    }
}

// test hook for the in-kernel exponential
__global__ void exp_neg_kernel(const double* __restrict__ y, double* __restrict__ out, int n)
{
    __shared__ double etab[32];
    fill_exp_table(etab);
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = exp_neg(y[i], etab);
}

}  // namespace ggp

using namespace ggp;

extern "C" {

int ggp_cov_build_f64(const double* X, int m, int d, const double* beta, const double* lamz,
                      const double* diag_add, int B, double* C_out, void* stream)
{
    GGP_ARG(X && beta && lamz && diag_add && C_out, "null pointer");
    GGP_ARG(m > 0 && d > 0 && B > 0, "m, d, B must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    int nt = (m + CT - 1) / CT;
    int ntile = nt * (nt + 1) / 2;
    size_t smem = (size_t)(2 * CT * d + CT * (CT + 1) + 32) * sizeof(double);
    GGP_ARG(smem <= 200 * 1024, "d too large for cov_build");
    GGP_CUDA(cudaFuncSetAttribute(cov_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cov_build_kernel<<<dim3(ntile, B), 256, smem, st>>>(X, m, d, beta, lamz, diag_add, C_out, ntile);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

int ggp_cross_cov_f64(const double* X, int m, const double* Xp, int n, int d, const double* beta,
                      const double* lamz, int B, double* S21_out, void* stream)
{
    GGP_ARG(X && Xp && beta && lamz && S21_out, "null pointer");
    GGP_ARG(m > 0 && n > 0 && d > 0 && B > 0, "m, n, d, B must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)(2 * CT * d) * sizeof(double);
    GGP_ARG(smem <= 200 * 1024, "d too large for cross_cov");
    GGP_CUDA(cudaFuncSetAttribute(cross_cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + CT - 1) / CT, (m + CT - 1) / CT, B);
    cross_cov_kernel<<<grid, 256, smem, st>>>(X, m, Xp, n, d, beta, lamz, S21_out);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

#ifdef GGP_PHASES
int ggp_debug_phase_cycles(unsigned long long* out_host, int reset)
{
    unsigned long long z[32] = {0};
    GGP_CUDA(cudaDeviceSynchronize());
    GGP_CUDA(cudaMemcpyFromSymbol(out_host, ggp::g_phase, sizeof(z)));
    if (reset) GGP_CUDA(cudaMemcpyToSymbol(ggp::g_phase, z, sizeof(z)));
    return GGP_OK;
}
int ggp_debug_phase2_cycles(unsigned long long* out_host, int reset)
{
    unsigned long long z[128] = {0};
    GGP_CUDA(cudaDeviceSynchronize());
    GGP_CUDA(cudaMemcpyFromSymbol(out_host, ggp::g_phase2, sizeof(z)));
    if (reset) GGP_CUDA(cudaMemcpyToSymbol(ggp::g_phase2, z, sizeof(z)));
    return GGP_OK;
}
#endif

int ggp_debug_exp_neg_f64(const double* y, double* out, int n, void* stream)
{
    GGP_ARG(y && out && n > 0, "bad argument");
    exp_neg_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(y, out, n);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

long long ggp_factor_doubles(int m) { return packed_doubles(round_up32(m)); }

int ggp_set_lookahead(int on)
{
    const int old = lookahead_flag();
    lookahead_flag() = on ? 1 : 0;
    return old;
}

int ggp_padded_m(int m) { return round_up32(m); }

int ggp_loglik_batched_f64(const double* X, int m, int d, const double* W, long long w_stride,
                           const double* beta, const double* lamz, const double* diag_add, int B,
                           double* factor_ws, double* u_out, double* loglik_out, int* info_out, void* stream)
{
    GGP_ARG(X && W && beta && lamz && diag_add && factor_ws && loglik_out, "null pointer");
    GGP_ARG(m > 0 && d > 0 && B > 0, "m, d, B must be positive");
    const int Mp = round_up32(m);
    size_t smem = eval_smem_bytes(Mp, d);
    if (smem > 227 * 1024) {
        set_error("ggp_loglik_batched_f64: m=%d d=%d needs %zu B of shared memory (> 227 KB)", m, d, smem);
        return GGP_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int G = choose_cluster(B, round_up32(m));
    const long long ls = packed_doubles(Mp);
    if (G > 1) {
        auto run = [&](auto kern) -> int {
            GGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int Gc = checked_cluster(kern, G, smem);
            GGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, eval_carveout_pct(smem)));
            GGP_CUDA(launch_maybe_cluster(kern, dim3(B * Gc), dim3(NT), smem, st, Gc, X, m, Mp, d, W, w_stride,
                                          beta, lamz, diag_add, factor_ws, ls, u_out, loglik_out, info_out));
            return GGP_OK;
        };
        const int rc = (Mp <= 1024) ? run(loglik_batched_kernel<true, false, 4>) : run(loglik_batched_kernel<true, false, 2>);
        if (rc != GGP_OK) return rc;
    } else if (use_lookahead()) {
        const size_t sla = la_smem_bytes(Mp, d);
        if (sla > 227 * 1024) {
            set_error("ggp_loglik_batched_f64: m=%d d=%d needs %zu B of shared memory (> 227 KB)", m, d, sla);
            return GGP_ERR_UNSUPPORTED;
        }
        GGP_CUDA(cudaFuncSetAttribute(loglik_batched_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sla));
        GGP_CUDA(cudaFuncSetAttribute(loglik_batched_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, eval_carveout_pct(sla)));
        loglik_batched_kernel<false, true><<<B, NT, sla, st>>>(X, m, Mp, d, W, w_stride, beta, lamz, diag_add, factor_ws, ls, u_out,
                                                               loglik_out, info_out);
    } else {
        GGP_CUDA(cudaFuncSetAttribute(loglik_batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GGP_CUDA(cudaFuncSetAttribute(loglik_batched_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, eval_carveout_pct(smem)));
        loglik_batched_kernel<false><<<B, NT, smem, st>>>(X, m, Mp, d, W, w_stride, beta, lamz, diag_add, factor_ws, ls, u_out,
                                                          loglik_out, info_out);
    }
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

int ggp_factor_unpack_f64(const double* factor_ws, int m, int B, double* L_dense, void* stream)
{
    GGP_ARG(factor_ws && L_dense && m > 0 && B > 0, "bad argument");
    const int Mp = round_up32(m);
    unpack_factor_kernel<<<dim3(64, B), 256, 0, (cudaStream_t)stream>>>(factor_ws, packed_doubles(Mp), m, Mp, L_dense);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

}  // extern "C"
