// Fused covariance build + blocked Cholesky + log-determinant + forward solve for ONE matrix per CTA.
//
// Replaces, for one (chain, PC) block evaluation, the reference chain
//   SepiaDistCov.compute_cov_mat -> np.fill_diagonal(nuggets) -> scipy.linalg.cholesky ->
//   scipy.linalg.solve_triangular -> -sum(log diag L) - 0.5 ||L^-1 w||^2
// (SURVEY.md 8a rows a3+a4 / Appendix A.10 cov_self, do_loglik, log_lik; called from
//  /root/reference/src/model.py:234-235 through sepia).
//
// Algorithm: left-looking blocked Cholesky, panel width 32.  For panel j the CTA
//   1. accumulates S = L[rows, 0:32j] * L[panel rows, 0:32j]^T on the FP64 tensor cores
//      (DMMA.8x8x4, accumulators in registers; A/B fragments are 16-byte loads straight from the
//      packed factor, which lives in global memory and stays L2 resident),
//   2. builds the covariance entries of the panel on the fly (exp fused in, nothing read from HBM)
//      and forms P = C - S in shared memory,
//   3. factors the 32x32 diagonal block with one warp (row per lane, shuffle broadcast of pivots),
//   4. solves the rows below with one thread per row, and updates the running forward solve of w,
//   5. writes the finished panel into the packed factor.
// The covariance matrix itself never exists in memory.
#pragma once
#include "ggp_common.cuh"

namespace ggp {

struct EvalSmem {
    double* LT;     // [32][LT_LD] transposed diagonal factor: LT[k][i] = L[i][k]
    double* rdiag;  // [32] reciprocal pivots
    double* uj;     // [32] forward-solve block
    double* red;    // [8]  scratch / broadcast
    double* Ps;     // [PASS_ROWS][PS_LD]
    double* wres;   // [Mp]
    double* S;      // [Mp][d] sqrt(beta)-scaled coordinates
    int* flag;      // [4]
};

__host__ __device__ inline size_t eval_smem_bytes(int Mp, int d) {
    return (size_t)(32 * LT_LD + 32 + 32 + 8 + PASS_ROWS * PS_LD + Mp + (size_t)Mp * d) * sizeof(double) + 16;
}

__device__ inline EvalSmem carve_eval_smem(unsigned char* base, int Mp, int d) {
    EvalSmem s;
    double* p = reinterpret_cast<double*>(base);
    s.LT = p;       p += 32 * LT_LD;
    s.rdiag = p;    p += 32;
    s.uj = p;       p += 32;
    s.red = p;      p += 8;
    s.Ps = p;       p += PASS_ROWS * PS_LD;
    s.wres = p;     p += Mp;
    s.S = p;        p += (size_t)Mp * d;
    s.flag = reinterpret_cast<int*>(p);
    return s;
}

// S += A[rows, 0:32j] * Lb[panel rows, 0:32j]^T for NU 8-row units of one warp (DMMA.8x8x4).
// A-operand rows come from `Ap` (packed factor layout when a_ld == 0, else a [slab][a_ld][8] layout
// indexed by local row), B-operand rows from the packed factor `Lb`.  NU is a template parameter so
// that no predicated-off DMMA is ever issued (a predicated-off DMMA still occupies the pipe).
template <int NU>
static __device__ __forceinline__ void panel_gemm(double (&acc)[4][4][2], const double* __restrict__ Ap,
                                                  const double* __restrict__ Lb, int Mp, int j, int row0,
                                                  const int (&rb)[4], int g, int q, int a_ld)
{
    const int nsl = 4 * j;
    auto a_slab = [&](int s) -> const double* {
        if (a_ld) return Ap + (size_t)s * a_ld * 8;
        int kb = s >> 2, ks = s & 3;
        return Ap + panel_off(kb, Mp) + (long long)ks * (Mp - 32 * kb) * 8 - 32LL * kb * 8;
    };
    auto b_slab = [&](int s) -> const double* {
        int kb = s >> 2, ks = s & 3;
        return Lb + panel_off(kb, Mp) + (long long)ks * (Mp - 32 * kb) * 8 - 32LL * kb * 8;
    };
    double2 an[NU];
    {
        const double* sl = a_slab(0);
#pragma unroll
        for (int i = 0; i < NU; ++i) an[i] = ldcg2(sl + (rb[i] + g) * 8 + 2 * q);
    }
    for (int s = 0; s < nsl; ++s) {
        const double* sl = b_slab(s);
        double2 a[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) a[i] = an[i];
        if (s + 1 < nsl) {
            const double* sn = a_slab(s + 1);
#pragma unroll
            for (int i = 0; i < NU; ++i) an[i] = ldcg2(sn + (rb[i] + g) * 8 + 2 * q);
        }
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const double2 b = *reinterpret_cast<const double2*>(sl + (row0 + 8 * cb + g) * 8 + 2 * q);
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                dmma884(acc[i][cb][0], acc[i][cb][1], a[i].x, b.x);
                dmma884(acc[i][cb][0], acc[i][cb][1], a[i].y, b.y);
            }
        }
    }
}

// Whole-CTA evaluation (NT threads).  Returns the block log-likelihood term to every thread;
// *info (if non-null, written by thread 0) = 0 or 1-based index of the failing pivot.
// beta may point to global or shared memory.  Lp is the packed-factor workspace (packed_doubles(Mp)).
// u_out (nullable): receives L^-1 w (Mp entries, zero padded).
static __device__ __forceinline__ double eval_block_loglik(const EvalSmem& sm, const double* __restrict__ X, int m, int Mp, int d,
                                    const double* beta, double lamz, double diag_add,
                                    const double* __restrict__ w, double* __restrict__ Lp,
                                    double* __restrict__ u_out, int* info)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nP = Mp >> 5;
    const double inv_lamz = 1.0 / lamz;
    const double diag = inv_lamz + diag_add;

    __syncthreads();   // previous user of the shared buffers is done
    for (int idx = tid; idx < Mp * d; idx += NT) {
        int r = idx / d, k = idx - r * d;
        sm.S[idx] = (r < m) ? X[(size_t)r * d + k] * sqrt(beta[k]) : 0.0;
    }
    for (int r = tid; r < Mp; r += NT) sm.wres[r] = (r < m) ? w[r] : 0.0;
    if (tid == 0) sm.flag[0] = 0;
    double logdet = 0.0, quad = 0.0;    // partial sums, live in warp 0
    __syncthreads();

    for (int j = 0; j < nP; ++j) {
        const int row0 = j << 5;
        const int Rj = Mp - row0;
        double* __restrict__ Lpj = Lp + panel_off(j, Mp);
        const int npass = (Rj + PASS_ROWS - 1) / PASS_ROWS;

        for (int ps = 0; ps < npass; ++ps) {
            const int r_lo = row0 + ps * PASS_ROWS;
            const int r_hi = min(Mp, r_lo + PASS_ROWS);
            const int nrows = r_hi - r_lo;
            const int nun = nrows >> 3;

            // ------------------------------------------------------------------ 1. DMMA update
            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) { acc[i][cb][0] = 0.0; acc[i][cb][1] = 0.0; }

            bool act[4];
            int rb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int u = warp + NWARP * i;
                act[i] = u < nun;
                rb[i] = r_lo + 8 * (act[i] ? u : 0);
            }

            if (act[0] && j > 0) {
                const int nmine = (nun - warp + NWARP - 1) / NWARP;     // units owned by this warp (1..4)
                switch (nmine) {
                    case 1: panel_gemm<1>(acc, Lp, Lp, Mp, j, row0, rb, g, q, 0); break;
                    case 2: panel_gemm<2>(acc, Lp, Lp, Mp, j, row0, rb, g, q, 0); break;
                    case 3: panel_gemm<3>(acc, Lp, Lp, Mp, j, row0, rb, g, q, 0); break;
                    default: panel_gemm<4>(acc, Lp, Lp, Mp, j, row0, rb, g, q, 0); break;
                }
            }

            // ------------------------------------------------------------------ 2. covariance, P = C - S
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (act[i]) {
                    const int r = rb[i] + g;
                    const double* Sr = sm.S + (size_t)r * d;
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = row0 + 8 * cb + 2 * q + e;
                            double v;
                            if (r == c) {
                                v = (r < m) ? diag : 1.0;
                            } else if (r < m && c < m) {
                                const double* Sc = sm.S + (size_t)c * d;
                                double dist = 0.0;
                                for (int k = 0; k < d; ++k) {
                                    double t = Sr[k] - Sc[k];
                                    dist = fma(t, t, dist);
                                }
                                v = exp(-dist) * inv_lamz;
                            } else {
                                v = 0.0;
                            }
                            sm.Ps[(r - r_lo) * PS_LD + (c - row0)] = v - acc[i][cb][e];
                        }
                    }
                }
            }
            __syncthreads();                                                     // #1

            // ------------------------------------------------------------------ 3. diagonal block
            if (ps == 0) {
                if (warp == 0) {
                    double x[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) x[c] = sm.Ps[lane * PS_LD + c];
                    double mypiv = 1.0;
                    bool ok = true;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        if (ok) {
                            const double dk = __shfl_sync(0xffffffffu, x[k], k);
                            if (!(dk > 0.0) || !(dk < 1.0e300)) {
                                ok = false;
                                if (lane == 0) sm.flag[0] = row0 + k + 1;
                            } else {
                                if (lane == k) mypiv = dk;
                                const double rk = rsqrt(dk);
                                double lik = x[k] * rk;
                                if (lane == k) lik = dk * rk;
                                x[k] = lik;
                                sm.LT[k * LT_LD + lane] = (lane >= k) ? lik : 0.0;
                                if (lane == 0) sm.rdiag[k] = rk;
                                __syncwarp();
#pragma unroll
                                for (int c = k + 1; c < 32; ++c) x[c] = fma(-lik, sm.LT[k * LT_LD + c], x[c]);
                            }
                        }
                    }
                    if (ok) logdet += 0.5 * log(mypiv);
#pragma unroll
                    for (int c = 0; c < 32; ++c) sm.Ps[lane * PS_LD + c] = (c > lane) ? 0.0 : x[c];
                }
                __syncthreads();                                                 // #2
                if (sm.flag[0] != 0) {
                    if (tid == 0 && info) *info = sm.flag[0];
                    return -INFINITY;
                }
            }

            // ------------------------------------------------------------------ 4. TRSM rows / forward solve of w
            const bool isdiag = (ps == 0) && (tid < 32);
            const bool myrow = tid < nrows;
            double x[32];                   // row of the panel owned by this thread
            if (myrow) {
#pragma unroll
                for (int c = 0; c < 32; ++c) x[c] = sm.Ps[tid * PS_LD + c];
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) x[c] = 0.0;
            }
            if (isdiag) {
                double b = sm.wres[row0 + lane];
                double myu = 0.0;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double uc = __shfl_sync(0xffffffffu, b, c) * sm.rdiag[c];
                    if (lane == c) myu = uc;
                    if (lane > c) b = fma(-x[c], uc, b);
                }
                sm.uj[lane] = myu;
                quad += myu * myu;
                if (u_out) u_out[row0 + lane] = myu;
            } else if (myrow) {
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
            }
            __syncthreads();                                                     // #3

            // ------------------------------------------------------------------ 5. store panel rows, update w
            if (myrow) {
                const int r = r_lo + tid;
                if (!isdiag) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        const double2 u2 = *reinterpret_cast<const double2*>(sm.uj + c);
                        s0 = fma(x[c], u2.x, s0);
                        s1 = fma(x[c + 1], u2.y, s1);
                    }
                    sm.wres[r] -= (s0 + s1);
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    double* dst = Lpj + (long long)ks * Rj * 8 + (long long)(r - row0) * 8;
#pragma unroll
                    for (int c = 0; c < 8; c += 2)
                        *reinterpret_cast<double2*>(dst + c) = make_double2(x[8 * ks + c], x[8 * ks + c + 1]);
                }
            }
            __syncthreads();                                                     // #4
        }
    }

    if (warp == 0) {
        double ld = warp_sum(logdet);
        double qd = warp_sum(quad);
        if (lane == 0) sm.red[0] = -ld - 0.5 * qd;
    }
    __syncthreads();
    if (tid == 0 && info) *info = 0;
    return sm.red[0];
}

}  // namespace ggp
