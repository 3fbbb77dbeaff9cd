// Posterior prediction kernels: predictive mean / variance through the cached factor, predictive
// covariance for joint draws, and the PC-basis reconstruction.
//
// Replaces SepiaPredict.wPred and SepiaEmulatorPrediction.get_y (SURVEY.md 8a rows a7, a8;
// Appendix A.7, A.10 w_pred / get_y; callers e.g.
// /root/reference/experiments/synthetic/analysis/assess_all_models.py:489-492).
//
// The reference solves S22 W = S21 from scratch for every (sample, PC, call).  Here the packed
// Cholesky factor L and u = L^-1 w of every (sample, PC) are computed once
// (ggp_loglik_batched_f64) and a block of test designs is pushed through
//   V = S21^T L^-T  (rows = designs; same left-looking DMMA panel loop as the factorisation),
//   mean = V u,  var = s11 - rowsum(V^2),  Sigma = S11 - V V^T.
#include "ggp_chol.cuh"
#include "../../include/gladsgp_b200.h"

namespace ggp {

constexpr int PB = PASS_ROWS;     // designs per task (one TRSM row per thread)

struct PredSmem {
    double* LT;     // [32][LT_LD]
    double* rdiag;  // [32]
    double* uj;     // [32]
    double* Ps;     // [PB][PS_LD]
    double* S;      // [Mp][d] scaled training coordinates
    double* Sp;     // [PB][d] scaled design coordinates
};

__host__ __device__ inline size_t pred_smem_bytes(int Mp, int d) {
    return (size_t)(32 * LT_LD + 32 + 32 + PB * PS_LD + (size_t)Mp * d + (size_t)PB * d) * sizeof(double);
}

__global__ void __launch_bounds__(NT, 1)
predict_kernel(const double* __restrict__ X, int m, int Mp, int d, const double* __restrict__ factor,
               long long l_stride, const double* __restrict__ U, const double* __restrict__ beta,
               const double* __restrict__ lamz, const double* __restrict__ s11,
               const double* __restrict__ Xp, int n, int B, double* __restrict__ mean_out,
               double* __restrict__ var_out, double* __restrict__ V_out, double* __restrict__ Vws)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PredSmem sm;
    {
        double* p = reinterpret_cast<double*>(smem_raw);
        sm.LT = p;    p += 32 * LT_LD;
        sm.rdiag = p; p += 32;
        sm.uj = p;    p += 32;
        sm.Ps = p;    p += PB * PS_LD;
        sm.S = p;     p += (size_t)Mp * d;
        sm.Sp = p;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nP = Mp >> 5;
    __shared__ int soff[1024];
    fill_slab_offsets(soff, Mp);
    const int nblk = (n + PB - 1) / PB;
    const long long ntask = (long long)B * nblk;
    double* __restrict__ Vw = Vws + (size_t)blockIdx.x * Mp * PB;   // [nP*4][PB][8]
    int cur_b = -1;

    for (long long task = blockIdx.x; task < ntask; task += gridDim.x) {
        const int b = (int)(task / nblk);
        const int t0 = (int)(task - (long long)b * nblk) * PB;
        const int nt = min(PB, n - t0);
        const double* be = beta + (size_t)b * d;
        const double inv_lamz = 1.0 / lamz[b];
        const double* __restrict__ Lp = factor + (size_t)b * l_stride;
        const double* __restrict__ ub = U + (size_t)b * Mp;
        __syncthreads();
        if (b != cur_b) {
            for (int idx = tid; idx < Mp * d; idx += NT) {
                int r = idx / d, k = idx - r * d;
                sm.S[idx] = (r < m) ? X[(size_t)r * d + k] * sqrt(be[k]) : 0.0;
            }
            cur_b = b;
        }
        for (int idx = tid; idx < PB * d; idx += NT) {
            int r = idx / d, k = idx - r * d;
            sm.Sp[idx] = (r < nt) ? Xp[(size_t)(t0 + r) * d + k] * sqrt(be[k]) : 0.0;
        }
        double mean = 0.0, vsum = 0.0;
        __syncthreads();

        for (int j = 0; j < nP; ++j) {
            const int row0 = j << 5;
            const int Rj = Mp - row0;
            const double* __restrict__ Lpj = Lp + panel_off(j, Mp);
            // ---- 1. S = V[:, 0:32j] * L[panel rows, 0:32j]^T on the FP64 tensor cores
            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { acc[i][cb][0] = 0.0; acc[i][cb][1] = 0.0; }
            if (j > 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int rb[2] = {8 * (warp + NWARP * (2 * h)), 8 * (warp + NWARP * (2 * h + 1))};
                    panel_gemm<2>(*reinterpret_cast<double (*)[2][4][2]>(&acc[2 * h]), Vw, Lp, soff, j, row0, rb, g, q, PB);
                }
            }
            // ---- 2. cross-covariance entries, P = S21^T - S
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 8 * (warp + NWARP * i) + g;          // local design index
                const double* Sr = sm.Sp + (size_t)r * d;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = row0 + 8 * cb + 2 * q + e;   // training point
                        double v = 0.0;
                        if (r < nt && c < m) {
                            const double* Sc = sm.S + (size_t)c * d;
                            double dist = 0.0;
                            for (int k = 0; k < d; ++k) {
                                double t = Sr[k] - Sc[k];
                                dist = fma(t, t, dist);
                            }
                            v = exp(-dist) * inv_lamz;
                        }
                        sm.Ps[r * PS_LD + (c - row0)] = v - acc[i][cb][e];
                    }
                }
            }
            // diagonal block of the factor -> LT, rdiag ; u block
            for (int idx = tid; idx < 1024; idx += NT) {
                const int i = idx >> 5, k = idx & 31;              // L[row0+i][row0+k]
                const double v = Lpj[(long long)(k >> 3) * Rj * 8 + (long long)i * 8 + (k & 7)];
                sm.LT[k * LT_LD + i] = v;
                if (i == k) sm.rdiag[k] = 1.0 / v;
            }
            if (tid < 32) sm.uj[tid] = ub[row0 + tid];
            __syncthreads();

            // ---- 3. one design per thread: x <- x * Ljj^-T ; accumulate mean and sum of squares
            double x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = sm.Ps[tid * PS_LD + c];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const double xc = x[c] * sm.rdiag[c];
                x[c] = xc;
                const double* lt = sm.LT + c * LT_LD;
                if (((c + 1) & 1) && c + 1 < 32) x[c + 1] = fma(-xc, lt[c + 1], x[c + 1]);
#pragma unroll
                for (int cp = (c + 2) & ~1; cp < 32; cp += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(lt + cp);
                    x[cp] = fma(-xc, l2.x, x[cp]);
                    x[cp + 1] = fma(-xc, l2.y, x[cp + 1]);
                }
            }
            {
                double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const double2 u2 = *reinterpret_cast<const double2*>(sm.uj + c);
                    s0 = fma(x[c], u2.x, s0);
                    s1 = fma(x[c + 1], u2.y, s1);
                    q0 = fma(x[c], x[c], q0);
                    q1 = fma(x[c + 1], x[c + 1], q1);
                }
                mean += s0 + s1;
                vsum += q0 + q1;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double* dst = Vw + ((size_t)(4 * j + ks) * PB + tid) * 8;
#pragma unroll
                for (int c = 0; c < 8; c += 2)
                    *reinterpret_cast<double2*>(dst + c) = make_double2(x[8 * ks + c], x[8 * ks + c + 1]);
            }
            if (V_out && tid < nt) {
                double* vo = V_out + ((size_t)b * n + t0 + tid) * Mp + row0;
#pragma unroll
                for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(vo + c) = make_double2(x[c], x[c + 1]);
            }
            __syncthreads();
        }
        if (tid < nt) {
            mean_out[(size_t)b * n + t0 + tid] = mean;
            var_out[(size_t)b * n + t0 + tid] = s11[b] - vsum;
        }
    }
}

// Sigma[b][t1][t2] = S11 - V V^T ; S11 diagonal = s11[b], off-diagonal = cov(xp_t1, xp_t2).
__global__ void __launch_bounds__(256)
pred_cov_kernel(const double* __restrict__ Xp, int n, int d, const double* __restrict__ beta,
                const double* __restrict__ lamz, const double* __restrict__ s11,
                const double* __restrict__ V, int Mp, double* __restrict__ Sigma)
{
    const int b = blockIdx.z;
    const int t1 = blockIdx.y * 16 + (threadIdx.x >> 4);
    const int t2 = blockIdx.x * 16 + (threadIdx.x & 15);
    if (t1 >= n || t2 >= n) return;
    const double* v1 = V + ((size_t)b * n + t1) * Mp;
    const double* v2 = V + ((size_t)b * n + t2) * Mp;
    double s0 = 0.0, s1 = 0.0;
    for (int k = 0; k < Mp; k += 2) {
        const double2 a = *reinterpret_cast<const double2*>(v1 + k);
        const double2 c = *reinterpret_cast<const double2*>(v2 + k);
        s0 = fma(a.x, c.x, s0);
        s1 = fma(a.y, c.y, s1);
    }
    double c11;
    if (t1 == t2) {
        c11 = s11[b];
    } else {
        const double* be = beta + (size_t)b * d;
        double dist = 0.0;
        for (int k = 0; k < d; ++k) {
            double df = Xp[(size_t)t1 * d + k] - Xp[(size_t)t2 * d + k];
            dist = fma(be[k] * df, df, dist);
        }
        c11 = exp(-dist) / lamz[b];
    }
    Sigma[((size_t)b * n + t1) * n + t2] = c11 - (s0 + s1);
}

// ---------------------------------------------------------------------------------------------
// PC-basis reconstruction  y[r][c] = (sum_p w[r][p] K[p][c]) * sd[c] + mean[c]   (get_y), float32.
// HBM-write bound: 4*R*n_y bytes.  One CTA = 1024 columns x RT rows; the K tile and the w rows sit
// in shared memory, each thread owns 4 columns and 4 rows at a time, stores are streaming.
// ---------------------------------------------------------------------------------------------
constexpr int RC_COLS = 1024;
constexpr int RC_RT = 32;

template <bool VEC>
__global__ void __launch_bounds__(256)
reconstruct_kernel(const float* __restrict__ w, const float* __restrict__ K, const float* __restrict__ sd,
                   int sd_len, const float* __restrict__ mu, int mu_len, int R, int pu, long long n_y,
                   float* __restrict__ y)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Ks = reinterpret_cast<float*>(smem_raw);           // [pu][RC_COLS]
    float* ws = Ks + (size_t)pu * RC_COLS;                    // [RC_RT][pu4]
    const int pu4 = (pu + 3) & ~3;
    const long long c0 = (long long)blockIdx.x * RC_COLS;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < pu * RC_COLS; idx += 256) {
        const int p = idx / RC_COLS, c = idx - p * RC_COLS;
        Ks[idx] = (c0 + c < n_y) ? K[(size_t)p * n_y + c0 + c] : 0.f;
    }
    long long col[4];
    float sdv[4], muv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        col[i] = VEC ? c0 + 4 * tid + i : c0 + tid + 256 * i;
        const bool in = col[i] < n_y;
        sdv[i] = in ? (sd_len == 1 ? sd[0] : sd[col[i]]) : 0.f;
        muv[i] = in ? (mu_len == 1 ? mu[0] : mu[col[i]]) : 0.f;
    }
    for (int r0 = blockIdx.y * RC_RT; r0 < R; r0 += gridDim.y * RC_RT) {
        __syncthreads();
        for (int idx = tid; idx < RC_RT * pu4; idx += 256) {
            const int rr = idx / pu4, p = idx - rr * pu4;
            ws[idx] = (r0 + rr < R && p < pu) ? w[(size_t)(r0 + rr) * pu + p] : 0.f;
        }
        __syncthreads();
        for (int rg = 0; rg < RC_RT; rg += 4) {
            if (r0 + rg >= R) break;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
            for (int p = 0; p < pu; ++p) {
                float kv[4];
                if (VEC) {
                    const float4 k4 = *reinterpret_cast<const float4*>(Ks + (size_t)p * RC_COLS + 4 * tid);
                    kv[0] = k4.x; kv[1] = k4.y; kv[2] = k4.z; kv[3] = k4.w;
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) kv[c] = Ks[(size_t)p * RC_COLS + tid + 256 * c];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float wv = ws[(rg + i) * pu4 + p];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(wv, kv[c], acc[i][c]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + rg + i;
                if (r >= R) break;
                float* yr = y + (size_t)r * n_y;
                if (VEC) {
                    if (col[3] < n_y) {
                        float4 o = make_float4(acc[i][0] * sdv[0] + muv[0], acc[i][1] * sdv[1] + muv[1],
                                               acc[i][2] * sdv[2] + muv[2], acc[i][3] * sdv[3] + muv[3]);
                        __stcs(reinterpret_cast<float4*>(yr + col[0]), o);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (col[c] < n_y) __stcs(yr + col[c], acc[i][c] * sdv[c] + muv[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (col[c] < n_y) __stcs(yr + col[c], acc[i][c] * sdv[c] + muv[c]);
                }
            }
        }
    }
}

}  // namespace ggp

using namespace ggp;

extern "C" {

static int predict_grid(int B, int n)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long ntask = (long long)B * ((n + PB - 1) / PB);
    return (int)(ntask < sms ? ntask : sms);
}

long long ggp_predict_workspace_bytes(int m, int n, int B)
{
    if (m <= 0 || n <= 0 || B <= 0) return -1;
    const int Mp = round_up32(m);
    return (long long)predict_grid(B, n) * Mp * PB * (long long)sizeof(double);
}

int ggp_predict_f64(const double* X, int m, int d, const double* factor, const double* u, const double* beta,
                    const double* lamz, const double* s11_diag, const double* Xp, int n, int B,
                    double* mean_out, double* var_out, double* V_out, void* workspace, long long workspace_bytes,
                    void* stream)
{
    GGP_ARG(X && factor && u && beta && lamz && s11_diag && Xp && mean_out && var_out && workspace, "null pointer");
    GGP_ARG(m > 0 && d > 0 && n > 0 && B > 0, "m, d, n, B must be positive");
    GGP_ARG(m <= 8192, "m must be <= 8192");
    const int Mp = round_up32(m);
    const size_t smem = pred_smem_bytes(Mp, d);
    if (smem > 227 * 1024) {
        set_error("ggp_predict_f64: m=%d d=%d needs %zu B of shared memory (> 227 KB)", m, d, smem);
        return GGP_ERR_UNSUPPORTED;
    }
    const long long need = ggp_predict_workspace_bytes(m, n, B);
    if (workspace_bytes < need) {
        set_error("ggp_predict_f64: workspace too small (%lld < %lld)", workspace_bytes, need);
        return GGP_ERR_WORKSPACE;
    }
    GGP_CUDA(cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    predict_kernel<<<predict_grid(B, n), NT, smem, (cudaStream_t)stream>>>(
        X, m, Mp, d, factor, packed_doubles(Mp), u, beta, lamz, s11_diag, Xp, n, B, mean_out, var_out, V_out,
        reinterpret_cast<double*>(workspace));
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

int ggp_pred_cov_f64(const double* Xp, int n, int d, const double* beta, const double* lamz,
                     const double* s11_diag, const double* V, int m, int B, double* Sigma_out, void* stream)
{
    GGP_ARG(Xp && beta && lamz && s11_diag && V && Sigma_out, "null pointer");
    GGP_ARG(n > 0 && d > 0 && m > 0 && B > 0, "n, d, m, B must be positive");
    dim3 grid((n + 15) / 16, (n + 15) / 16, B);
    pred_cov_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Xp, n, d, beta, lamz, s11_diag, V, round_up32(m), Sigma_out);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

int ggp_reconstruct_f32(const float* w, const float* K, const float* sd, int sd_len, const float* mean,
                        int mean_len, int R, int pu, long long n_y, float* y_out, void* stream)
{
    GGP_ARG(w && K && sd && mean && y_out, "null pointer");
    GGP_ARG(R > 0 && pu > 0 && n_y > 0, "R, pu, n_y must be positive");
    GGP_ARG((sd_len == 1 || sd_len == n_y) && (mean_len == 1 || mean_len == n_y), "sd/mean length must be 1 or n_y");
    const int pu4 = (pu + 3) & ~3;
    const size_t smem = ((size_t)pu * RC_COLS + (size_t)RC_RT * pu4) * sizeof(float);
    GGP_ARG(smem <= 200 * 1024, "pu too large for reconstruct");
    const unsigned gx = (unsigned)((n_y + RC_COLS - 1) / RC_COLS);
    unsigned gy = (unsigned)((R + RC_RT - 1) / RC_RT);
    if (gy > 65535u) gy = 65535u;
    const bool vec = (n_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(y_out) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) {
        GGP_CUDA(cudaFuncSetAttribute(reconstruct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        reconstruct_kernel<true><<<dim3(gx, gy), 256, smem, st>>>(w, K, sd, sd_len, mean, mean_len, R, pu, n_y, y_out);
    } else {
        GGP_CUDA(cudaFuncSetAttribute(reconstruct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        reconstruct_kernel<false><<<dim3(gx, gy), 256, smem, st>>>(w, K, sd, sd_len, mean, mean_len, R, pu, n_y, y_out);
    }
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

}  // extern "C"
