// Saltelli first-order / total sensitivity statistics for a batch of index sets.
//
// Replaces the per-resample NumPy evaluation inside the reference's bootstrap of the Sobol' indices
//   /root/reference/src/utils.py:80-92    point estimates   V_i = mean(f_A (f_C - f_B)), E_i = mean((f_B - f_C)^2)/2, var = var([f_A, f_B])
//   /root/reference/src/utils.py:97-118   first_order_statistic / total_index_statistic (scalar outputs)
//   /root/reference/src/utils.py:213-243  the same for PC weights (callers experiments/synthetic/analysis/sensitivity_indices.py:96,214)
// which scipy.stats.bootstrap calls 9999 times per statistic plus N times for the BCa jackknife, each time gathering
// N rows of f_A, f_B and f_AB with fancy indexing.  One CTA per index set; the function values (a few hundred KB) stay
// in L2, the index set sits in shared memory.  C-ABI in include/gladsgp_b200.h.
#include "ggp_common.cuh"
#include "../../include/gladsgp_b200.h"

namespace ggp {

constexpr int SOB_NT = 128;

// fA, fB [N][p]; fAB [nd][N][p]; idx [R][idx_stride] (entries in [0, N)) or null = the identity set 0..n-1.
// first / total [R][p][nd]:  first = V / var, total = E / var with V, E clamped at 0 when `clamp` (the reference clamps inside
// its bootstrap statistics, not in the point estimates).
__global__ void __launch_bounds__(SOB_NT)
sobol_stats_kernel(const double* __restrict__ fA, const double* __restrict__ fB, const double* __restrict__ fAB,
                   const int* __restrict__ idx, long long idx_stride, int n, int N, int p, int nd, int clamp,
                   double* __restrict__ first, double* __restrict__ total)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* ids = reinterpret_cast<int*>(smem_raw);                               // [n]
    double* var = reinterpret_cast<double*>(smem_raw + (((size_t)n * sizeof(int) + 15) & ~(size_t)15));   // [p]
    const int r = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += SOB_NT) ids[i] = idx ? idx[(size_t)r * idx_stride + i] : i;
    __syncthreads();
    // population variance of the 2n values [f_A[ids], f_B[ids]] per output (np.var(..., axis=(0, 1)): two passes)
    for (int pc = threadIdx.x; pc < p; pc += SOB_NT) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const size_t o = (size_t)ids[i] * p + pc;
            s += fA[o];
            s += fB[o];
        }
        const double mu = s / (2.0 * n);
        double ss = 0.0;
        for (int i = 0; i < n; ++i) {
            const size_t o = (size_t)ids[i] * p + pc;
            const double da = fA[o] - mu, db = fB[o] - mu;
            ss += da * da;
            ss += db * db;
        }
        var[pc] = ss / (2.0 * n);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nd * p; t += SOB_NT) {
        const int ix = t / p, pc = t - ix * p;
        const double* fC = fAB + (size_t)ix * N * p;
        double V = 0.0, E = 0.0;
        for (int i = 0; i < n; ++i) {
            const size_t o = (size_t)ids[i] * p + pc;
            const double a = fA[o], b = fB[o], c = fC[o];
            V += a * (c - b);
            const double e = b - c;
            E += e * e;
        }
        V = V / n;
        E = 0.5 * (E / n);
        if (clamp) {
            if (V < 0.0) V = 0.0;
            if (E < 0.0) E = 0.0;
        }
        const size_t oo = ((size_t)r * p + pc) * nd + ix;
        first[oo] = V / var[pc];
        total[oo] = E / var[pc];
    }
}

}  // namespace ggp

using namespace ggp;

extern "C" {

int ggp_sobol_stats_f64(const double* fA, const double* fB, const double* fAB, int N, int p, int n_dim,
                        const int* idx, long long idx_stride, int n, int R, int clamp,
                        double* first_out, double* total_out, void* stream)
{
    GGP_ARG(fA && fB && fAB && first_out && total_out, "null pointer");
    GGP_ARG(N > 0 && p > 0 && n_dim > 0 && n > 0 && R > 0, "sizes must be positive");
    GGP_ARG(idx || n <= N, "identity index set longer than the sample");
    GGP_ARG(!idx || idx_stride >= n, "idx_stride < n");
    const size_t smem = (((size_t)n * sizeof(int) + 15) & ~(size_t)15) + (size_t)p * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("ggp_sobol_stats_f64: n=%d p=%d needs %zu B of shared memory", n, p, smem);
        return GGP_ERR_UNSUPPORTED;
    }
    GGP_CUDA(cudaFuncSetAttribute(sobol_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sobol_stats_kernel<<<R, SOB_NT, smem, (cudaStream_t)stream>>>(fA, fB, fAB, idx, idx_stride, n, N, p, n_dim, clamp,
                                                                  first_out, total_out);
    GGP_CUDA(cudaGetLastError());
    return GGP_OK;
}

}  // extern "C"
