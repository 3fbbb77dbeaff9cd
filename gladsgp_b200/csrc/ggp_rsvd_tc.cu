// Randomized-SVD passes on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Same products as ggp_rsvd.cu (/root/reference/src/svd.py:52-60) -- Y = X Omega and Bt = Y^T X over the (m x n)
// float32 ensemble -- but the multiply-adds run as 3xTF32 split products on the tensor pipe, so that the pass is
// bound by the HBM read of X instead of the FP32 FMA rate:
//     x = x_hi + x_lo (x_hi = tf32(x), x_lo = tf32(x - x_hi)),   x*o ~= x_hi*o_hi + x_hi*o_lo + x_lo*o_hi
// (error ~2^-21 relative per product, i.e. FP32-level; the dropped x_lo*o_lo term is 2^-22 smaller than the product).
// The tensor core adds into its FP32 accumulator with truncation, so a K-chunk (32 columns, 12 accumulator updates)
// is summed in TMEM and then added to round-to-nearest FP32 register accumulators by the CUDA cores -- the long sum
// over n never runs inside the tensor core.
//
// Operands are staged by the producer warps: global -> registers (split hi/lo) -> shared memory in the UMMA
// canonical K-major 128-byte-swizzle layout (the one TMA would write; TMA cannot be used because the row pitch of
// the reference's ensembles, 4*n_y bytes with n_y = 3693*365, is not a multiple of 16 bytes).  One elected thread of
// the MMA warp issues tcgen05.mma and signals completion with tcgen05.commit on an mbarrier.
#include "ggp_common.cuh"
#include "../../include/gladsgp_b200.h"

namespace ggp {

constexpr int TC_ROWS = 256;                 // rows of X per CTA (two M = 128 tiles)
constexpr int TC_K = 32;                     // columns of X per chunk (four K = 8 steps)
constexpr int TC_STAGES = 3;
constexpr int TC_PRODUCERS = 256;            // 8 producer warps
constexpr int TC_THREADS = TC_PRODUCERS + 128;  // + 4 epilogue warps (the first one also issues the MMAs)
constexpr int TC_A_BYTES = TC_ROWS * TC_K * 4;  // 32 KB (one of hi / lo)
constexpr int TC_B_BYTES = 32 * TC_K * 4;       // 4 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;    // 72 KB
constexpr int TC_TMEM_COLS = 128;            // 2 buffers x 2 tiles x 32 columns

#ifndef GGP_TC_LD
#define GGP_TC_LD __ldcs
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol error traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    for (int spin = 0; spin < 4000; ++spin) {               // each try_wait may suspend the thread for up to 1 ms
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(1000000u)
            : "memory");
        if (ok) return;
    }
    __trap();
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, 128-byte swizzle: a row is 128 contiguous bytes (32 tf32 = one K chunk), rows 128 B apart, 8-row groups
// 1024 B apart, and the 16-byte piece kc of row r sits at position kc ^ (r % 8) (Swizzle<3,4,3>; tile 1024-B aligned).
// A K = 8 step advances the start address by 32 bytes.
// The same descriptor as two words, so that the issuing thread only adds a byte offset to the low word per MMA
// (it is a single thread: every instruction it spends on descriptors is on the critical path of the chunk).
constexpr uint32_t UMMA_SW128_HI = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_sw128_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_join(uint32_t lo, uint32_t hi)
{
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool ACC>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// issue all tcgen05.ld of an epilogue first, then wait once
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = hi + lo with hi on the TF32 grid (round to nearest, ties away -- what cvt.rna.tf32.f32 does, but on the integer
// pipe: the conversion instruction runs on the quarter-rate XU pipe and was the producers' bottleneck) and lo = x - hi
// exact in FP32 (|lo| <= 2^-11 |x|); the tensor core reads the upper 19 bits of lo (error <= 2^-21 |x|).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// ---------------------------------------------------------------------------------------------------------
// Y = X * Omega.  grid (gx, row blocks of 256): CTA (bx, by) takes the 32-column chunks bx, bx + gx, ... of its
// rows and writes partial[bx][m][32]; tc_reduce_kernel sums the partials in FP64 in a fixed order.
//
// Warp roles (12 warps, 168 registers each = 3 warps per SM sub-partition):
//   warps 0-7   producers: global -> registers (two chunks ahead, three rotating buffers) -> hi/lo split -> swizzled
//               shared-memory stage -> fence.proxy.async -> arrive on full[stage]
//   warp 8      lane 0 issues the 24 tcgen05.mma of a chunk (2 M-tiles x 4 K-steps x 3 split products) and commits
//               to empty[stage] (stage reusable) and tmem_full[buffer] (chunk sum ready)
//   warps 8-11  epilogue: tcgen05.ld the chunk sum (two 128 x 32 tiles, TMEM double-buffered) and add it to FP32
//               register accumulators with round-to-nearest
template <bool VEC, int NPW>
__global__ void __launch_bounds__(NPW * 32 + 128, 1)
sketch_tc_kernel(const float* __restrict__ X, int m, long long n, const float* __restrict__ OmT, int r, int k0,
                 float* __restrict__ partial)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_base = blockIdx.y * TC_ROWS;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], NPW * 32);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar[0], 1);
        mbar_init(&tmem_full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NPW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    const long long nchunk = (n + TC_K - 1) / TC_K;
#ifdef GGP_TC_INTERLEAVE
    const int n_my = (blockIdx.x < nchunk) ? (int)((nchunk - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    auto chunk_col = [&](int it) { return ((long long)blockIdx.x + (long long)it * gridDim.x) * TC_K; };
#else
    // a contiguous range of chunks per CTA: successive loads of a thread walk along one row of X
    const long long per = (nchunk + gridDim.x - 1) / gridDim.x;
    const long long ch_begin = (long long)blockIdx.x * per;
    const int n_my = (int)(ch_begin >= nchunk ? 0 : (nchunk - ch_begin < per ? nchunk - ch_begin : per));
    auto chunk_col = [&](int it) { return (ch_begin + it) * TC_K; };
#endif

    if (warp < NPW) {
        // ================= producers =================
        // A staging: warp w, pass u, lane l -> row 32 w + 4 u + l/8, 16-byte K piece l%8 (full 128-byte lines from
        // global, conflict-free STS.128 into the swizzled tile).  B staging: thread t -> Omega row t/8, K piece t%8.
        constexpr int UP = 64 / NPW;                                       // 128-byte row pieces per thread and chunk
        const bool b_thread = tid < 256;                                   // threads that also stage Omega
        const int b_row = (tid >> 3) & 31, b_kc = tid & 7;
        const bool b_ok = b_thread && (k0 + b_row < r);
        const int lrow0 = warp * (4 * UP) + (lane >> 3);                         // local row of pass 0 (pass u: + 4 u)
        const float* xp = X + (size_t)(row_base + lrow0) * n + 4 * (lane & 7);      // + 4 u n + c0
        const float* bp = OmT + (size_t)(k0 + (b_ok ? b_row : 0)) * n + 4 * b_kc;
        const size_t pass_stride = 4 * (size_t)n;
        unsigned row_ok = 0;
#pragma unroll
        for (int u = 0; u < UP; ++u) row_ok |= (row_base + lrow0 + 4 * u < m) ? (1u << u) : 0u;

        auto load_chunk = [&](int it, float (&xv)[UP][4], float (&bv)[4]) {
            const long long c0 = chunk_col(it);
            const float* src0 = xp + c0;
            if (c0 + TC_K <= n) {                               // whole chunk inside the matrix (CTA-uniform)
#pragma unroll
                for (int u = 0; u < UP; ++u) {
                    const float* src = src0 + u * pass_stride;
                    if (!((row_ok >> u) & 1)) {
                        xv[u][0] = 0.f; xv[u][1] = 0.f; xv[u][2] = 0.f; xv[u][3] = 0.f;
                    } else if (VEC) {
                        const float4 q = GGP_TC_LD(reinterpret_cast<const float4*>(src));
                        xv[u][0] = q.x; xv[u][1] = q.y; xv[u][2] = q.z; xv[u][3] = q.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) xv[u][j] = __ldcs(src + j);
                    }
                }
                if (!b_ok) {
                    bv[0] = 0.f; bv[1] = 0.f; bv[2] = 0.f; bv[3] = 0.f;
                } else if (VEC) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(bp + c0));
                    bv[0] = q.x; bv[1] = q.y; bv[2] = q.z; bv[3] = q.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = __ldg(bp + c0 + j);
                }
            } else {                                            // ragged last chunk
                const long long col = c0 + 4 * (lane & 7);
#pragma unroll
                for (int u = 0; u < UP; ++u)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        xv[u][j] = (((row_ok >> u) & 1) && col + j < n) ? __ldcs(src0 + u * pass_stride + j) : 0.f;
                const long long colb = c0 + 4 * b_kc;
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = (b_ok && colb + j < n) ? __ldg(bp + c0 + j) : 0.f;
            }
        };
        // one pipeline step: chunk `it` is in (xv, bv); chunk it+2 is loaded into (xl, bl)
        auto step = [&](int it, float (&xv)[UP][4], float (&bv)[4], float (&xl)[UP][4], float (&bl)[4]) {
            if (it >= n_my) return;
            if (it + 2 < n_my) load_chunk(it + 2, xl, bl);
            const int s = it % TC_STAGES;
            if (it >= TC_STAGES) mbar_wait(&empty_bar[s], (uint32_t)((it / TC_STAGES - 1) & 1));   // MMAs of chunk it-3 done
            unsigned char* st = smem_raw + (size_t)s * TC_STAGE_BYTES;
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                const uint32_t lrow = (uint32_t)(lrow0 + 4 * u);
                const uint32_t off = lrow * 128u + (uint32_t)(((lane & 7) ^ (lrow & 7)) * 16);
                uint4 hi, lo;
                split_tf32(xv[u][0], hi.x, lo.x);
                split_tf32(xv[u][1], hi.y, lo.y);
                split_tf32(xv[u][2], hi.z, lo.z);
                split_tf32(xv[u][3], hi.w, lo.w);
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + TC_A_BYTES + off) = lo;
            }
            if (b_thread) {
                const uint32_t off = (uint32_t)b_row * 128u + (uint32_t)((b_kc ^ (b_row & 7)) * 16);
                uint4 hi, lo;
                split_tf32(bv[0], hi.x, lo.x);
                split_tf32(bv[1], hi.y, lo.y);
                split_tf32(bv[2], hi.z, lo.z);
                split_tf32(bv[3], hi.w, lo.w);
                *reinterpret_cast<uint4*>(st + 2 * TC_A_BYTES + off) = hi;
                *reinterpret_cast<uint4*>(st + 2 * TC_A_BYTES + TC_B_BYTES + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(&full_bar[s]);
        };
        float x0[UP][4], x1[UP][4], x2[UP][4], b0[4], b1[4], b2[4];
        if (n_my > 0) load_chunk(0, x0, b0);
        if (n_my > 1) load_chunk(1, x1, b1);
#pragma unroll 1
        for (int it = 0; it < n_my; it += 3) {
            step(it, x0, b0, x2, b2);
            step(it + 1, x1, b1, x0, b0);
            step(it + 2, x2, b2, x1, b1);
        }
    } else {
        // ================= MMA issue (warp 8, lane 0) + epilogue (warps 8-11) =================
        constexpr uint32_t idesc = umma_idesc_tf32(128, 32);
        const int ew = warp - NPW;                                             // TMEM lanes 32 ew .. 32 ew + 31
        const uint32_t t_lane = (uint32_t)(32 * ew) << 16;
        const uint32_t desc_lo0 = umma_sw128_lo(smem_u32(smem_raw));
        float acc[2][32];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[t][c] = 0.f;
#pragma unroll 1
        for (int it = 0; it <= n_my; ++it) {
            if (it < n_my && warp == NPW && lane == 0) {
                const int s = it % TC_STAGES;
                mbar_wait(&full_bar[s], (uint32_t)((it / TC_STAGES) & 1));
                tc_fence_after();
                // descriptor low words: base + (byte offset >> 4); A tile t-step: + 2, next M tile: + 1024, lo: + 2048
                const uint32_t a_lo0 = desc_lo0 + (uint32_t)s * (TC_STAGE_BYTES >> 4);
                const uint32_t b_lo0 = a_lo0 + ((2 * TC_A_BYTES) >> 4);
                const uint32_t d0 = tmem_base + (uint32_t)((it & 1) * 64);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint64_t bh = umma_join(b_lo0 + 2 * t, UMMA_SW128_HI);
                    const uint64_t bl = umma_join(b_lo0 + (TC_B_BYTES >> 4) + 2 * t, UMMA_SW128_HI);
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
                        const uint64_t ah = umma_join(a_lo0 + tile * 1024 + 2 * t, UMMA_SW128_HI);
                        const uint64_t al = umma_join(a_lo0 + (TC_A_BYTES >> 4) + tile * 1024 + 2 * t, UMMA_SW128_HI);
                        const uint32_t d = d0 + tile * 32;
                        if (t == 0) umma_tf32<false>(d, al, bh, idesc); else umma_tf32<true>(d, al, bh, idesc);   // small terms first
                        umma_tf32<true>(d, ah, bl, idesc);
                        umma_tf32<true>(d, ah, bh, idesc);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tmem_full_bar[it & 1]);
            }
            __syncwarp();
            if (it >= 1) {
                // chunk it-1 has been multiplied: add its TMEM tiles into the register accumulators
                const int j = it - 1;
                mbar_wait(&tmem_full_bar[j & 1], (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                uint32_t v0[32], v1[32];
                tmem_ld32(tmem_base + t_lane + (uint32_t)((j & 1) * 64), v0);
                tmem_ld32(tmem_base + t_lane + (uint32_t)((j & 1) * 64 + 32), v1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[0][c] += __uint_as_float(v0[c]);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[1][c] += __uint_as_float(v1[c]);
                tc_fence_before();
            }
            // all four epilogue warps are done with TMEM buffer (it-1)&1 before the MMAs of chunk it+1 overwrite it
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
            const int row = row_base + tile * 128 + 32 * ew + lane;
            if (row < m) {
                float4* out = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * m + row) * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    out[c] = make_float4(acc[tile][4 * c], acc[tile][4 * c + 1], acc[tile][4 * c + 2], acc[tile][4 * c + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// ---- A operand from tensor memory (tcgen05.mma "TS" form): lanes = M rows, one 32-bit column per K element ----
template <bool ACC>
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int TS_STAGES = 3;
constexpr int TS_B_STAGE_BYTES = 2 * TC_B_BYTES;     // Omega / Y chunk, hi + lo (8 KB)
constexpr int TS_TMEM_COLS = 512;                    // 3 stages x 2 tiles x (32 hi + 32 lo) + 2 buffers x 2 tiles x 32
constexpr uint32_t TS_D_COL = 384;

// ---------------------------------------------------------------------------------------------------------
// Bt = Y^T X with the A operand (X^T tile: M = 256 columns of X, K = 32 rows) written by the producers straight
// into tensor memory: lane = column, so every warp load reads 128 contiguous bytes of one row of X and no
// shared-memory staging or transposition of X is needed at all; only the small Y chunk goes through shared memory.
__global__ void __launch_bounds__(TC_THREADS, 1)
xty_ts_kernel(const float* __restrict__ X, int m, long long n, const float* __restrict__ Y, int r, int k0,
              float* __restrict__ Bt)
{
    __shared__ __align__(1024) unsigned char bsm[TS_STAGES * TS_B_STAGE_BYTES];
    __shared__ __align__(8) uint64_t full_bar[TS_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TS_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < TS_STAGES; ++s) {
            mbar_init(&full_bar[s], TC_PRODUCERS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar[0], 1);
        mbar_init(&tmem_full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(TS_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    const int nk = (m + 31) / 32;
    const long long nblk = (n + TC_ROWS - 1) / TC_ROWS;
    const int blk_my = (blockIdx.x < nblk) ? (int)((nblk - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    const int n_my = blk_my * nk;
    auto blk_col = [&](int b) { return ((long long)blockIdx.x + (long long)b * gridDim.x) * TC_ROWS; };

    if (warp < 8) {
        // ================= producers: lane <-> column (warp w: tile w/4, TMEM lanes 32 (w%4) ..) =================
        const int b_n = tid >> 3, b_kc = tid & 7;
        const bool b_ok = (k0 + b_n < r);
        const int col_in_blk = (warp >> 2) * 128 + 32 * (warp & 3) + lane;
        const uint32_t t_lane = (uint32_t)(32 * (warp & 3)) << 16;
        const uint32_t a_col = (uint32_t)((warp >> 2) * 64);                // + stage * 128 (+ 32 for lo)
        auto load_chunk = [&](int it, float (&xv)[32], float (&bv)[4]) {
            const int b = it / nk, kc = it - b * nk;
            const long long col = blk_col(b) + col_in_blk;
            const int i0 = kc * 32;
            const float* src = X + (size_t)i0 * n + col;
            if (col < n && i0 + 32 <= m) {
#pragma unroll
                for (int k = 0; k < 32; ++k) xv[k] = GGP_TC_LD(src + (size_t)k * n);
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) xv[k] = (col < n && i0 + k < m) ? GGP_TC_LD(src + (size_t)k * n) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = i0 + 4 * b_kc + j;
                bv[j] = (b_ok && row < m) ? __ldg(Y + (size_t)row * r + k0 + b_n) : 0.f;
            }
        };
        auto step = [&](int it, float (&xv)[32], float (&bv)[4], float (&xl)[32], float (&bl)[4]) {
            if (it >= n_my) return;
            if (it + 2 < n_my) load_chunk(it + 2, xl, bl);
            const int s = it % TS_STAGES;
            if (it >= TS_STAGES) {
                mbar_wait(&empty_bar[s], (uint32_t)((it / TS_STAGES - 1) & 1));     // MMAs of chunk it-3 have read the stage
                tc_fence_after();
            }
            const uint32_t ta = tmem_base + t_lane + (uint32_t)(s * 128) + a_col;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) split_tf32(xv[16 * h + k], hi[k], lo[k]);
                tmem_st16(ta + 16 * h, hi);
                tmem_st16(ta + 32 + 16 * h, lo);
            }
            {
                unsigned char* st = bsm + (size_t)s * TS_B_STAGE_BYTES;
                const uint32_t off = (uint32_t)b_n * 128u + (uint32_t)((b_kc ^ (b_n & 7)) * 16);
                uint4 hi, lo;
                split_tf32(bv[0], hi.x, lo.x);
                split_tf32(bv[1], hi.y, lo.y);
                split_tf32(bv[2], hi.z, lo.z);
                split_tf32(bv[3], hi.w, lo.w);
                *reinterpret_cast<uint4*>(st + off) = hi;
                *reinterpret_cast<uint4*>(st + TC_B_BYTES + off) = lo;
            }
            tmem_wait_st();
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(&full_bar[s]);
        };
        float x0[32], x1[32], x2[32], b0[4], b1[4], b2[4];
        if (n_my > 0) load_chunk(0, x0, b0);
        if (n_my > 1) load_chunk(1, x1, b1);
#pragma unroll 1
        for (int it = 0; it < n_my; it += 3) {
            step(it, x0, b0, x2, b2);
            step(it + 1, x1, b1, x0, b0);
            step(it + 2, x2, b2, x1, b1);
        }
    } else {
        // ================= MMA issue (warp 8, lane 0) + epilogue (warps 8-11) =================
        constexpr uint32_t idesc = umma_idesc_tf32(128, 32);
        const int ew = warp - 8;
        const uint32_t t_lane = (uint32_t)(32 * ew) << 16;
        const uint32_t desc_lo0 = umma_sw128_lo(smem_u32(bsm));
        float acc[2][32];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[t][c] = 0.f;
#pragma unroll 1
        for (int it = 0; it <= n_my; ++it) {
            if (it < n_my && warp == 8 && lane == 0) {
                const int s = it % TS_STAGES;
                mbar_wait(&full_bar[s], (uint32_t)((it / TS_STAGES) & 1));
                tc_fence_after();
                const uint32_t b_lo0 = desc_lo0 + (uint32_t)s * (TS_B_STAGE_BYTES >> 4);
                const uint32_t d0 = tmem_base + TS_D_COL + (uint32_t)((it & 1) * 64);
                const uint32_t a0 = tmem_base + (uint32_t)(s * 128);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint64_t bh = umma_join(b_lo0 + 2 * t, UMMA_SW128_HI);
                    const uint64_t bl = umma_join(b_lo0 + (TC_B_BYTES >> 4) + 2 * t, UMMA_SW128_HI);
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
                        const uint32_t d = d0 + tile * 32, a_hi = a0 + tile * 64 + 8 * t;
                        if (t == 0) umma_tf32_ts<false>(d, a_hi + 32, bh, idesc); else umma_tf32_ts<true>(d, a_hi + 32, bh, idesc);   // x_lo * y_hi
                        umma_tf32_ts<true>(d, a_hi, bl, idesc);                                    // x_hi * y_lo
                        umma_tf32_ts<true>(d, a_hi, bh, idesc);                                    // x_hi * y_hi
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tmem_full_bar[it & 1]);
            }
            __syncwarp();
            if (it >= 1) {
                const int j = it - 1;
                mbar_wait(&tmem_full_bar[j & 1], (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                uint32_t v0[32], v1[32];
                tmem_ld32(tmem_base + t_lane + TS_D_COL + (uint32_t)((j & 1) * 64), v0);
                tmem_ld32(tmem_base + t_lane + TS_D_COL + (uint32_t)((j & 1) * 64 + 32), v1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[0][c] += __uint_as_float(v0[c]);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[1][c] += __uint_as_float(v1[c]);
                tc_fence_before();
                if ((j + 1) % nk == 0) {
                    const long long c0 = blk_col(j / nk);
#pragma unroll
                    for (int tile = 0; tile < 2; ++tile) {
                        const long long col = c0 + tile * 128 + 32 * ew + lane;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (col < n && k0 + c < r) __stcs(Bt + (size_t)(k0 + c) * n + col, acc[tile][c]);
                            acc[tile][c] = 0.f;
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TS_TMEM_COLS));
    }
}

// fixed-order FP64 sum of the per-CTA partials [nparts][m][32] -> Y[m][r] columns k0 .. k0+31
__global__ void tc_reduce_kernel(const float* __restrict__ partial, int nparts, int m, int r, int k0,
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

static int tc_sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace ggp

using namespace ggp;

extern "C" {

long long ggp_rsvd_tc_workspace_bytes(int m)
{
    if (m <= 0) return -1;
    return (long long)tc_sm_count() * m * 32 * (long long)sizeof(float);
}

int ggp_rsvd_sketch_tc_f32(const float* X, int m, long long n, const float* OmegaT, int r, float* Y_out, void* workspace,
                           long long workspace_bytes, void* stream)
{
    GGP_ARG(X && OmegaT && Y_out && workspace, "null pointer");
    GGP_ARG(m > 0 && n > 0 && r > 0, "m, n, r must be positive");
    if (workspace_bytes < ggp_rsvd_tc_workspace_bytes(m)) {
        set_error("ggp_rsvd_sketch_tc_f32: workspace too small");
        return GGP_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    const int gy = (m + TC_ROWS - 1) / TC_ROWS;
    const long long nchunk = (n + TC_K - 1) / TC_K;
    long long gx = tc_sm_count() / gy;                       // one CTA per SM (216 KB of shared memory each)
    if (gx < 1) gx = 1;
    if (gx > nchunk) gx = nchunk;
    const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    GGP_CUDA(cudaFuncSetAttribute(sketch_tc_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GGP_CUDA(cudaFuncSetAttribute(sketch_tc_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int k0 = 0; k0 < r; k0 += 32) {
        if (vec) sketch_tc_kernel<true, 8><<<dim3((unsigned)gx, gy), 384, smem, st>>>(X, m, n, OmegaT, r, k0, partial);
        else sketch_tc_kernel<false, 8><<<dim3((unsigned)gx, gy), 384, smem, st>>>(X, m, n, OmegaT, r, k0, partial);
        GGP_CUDA(cudaGetLastError());
        tc_reduce_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(partial, (int)gx, m, r, k0, Y_out);
        GGP_CUDA(cudaGetLastError());
    }
    return GGP_OK;
}

int ggp_rsvd_xty_tc_f32(const float* X, int m, long long n, const float* Y, int r, float* Bt_out, void* stream)
{
    GGP_ARG(X && Y && Bt_out, "null pointer");
    GGP_ARG(m > 0 && n > 0 && r > 0, "m, n, r must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nblk = (n + TC_ROWS - 1) / TC_ROWS;
    long long gx = tc_sm_count();
    if (gx > nblk) gx = nblk;
    for (int k0 = 0; k0 < r; k0 += 32) {
        xty_ts_kernel<<<(unsigned)gx, TC_THREADS, 0, st>>>(X, m, n, Y, r, k0, Bt_out);
        GGP_CUDA(cudaGetLastError());
    }
    return GGP_OK;
}

}  // extern "C"
