// Micro-benchmark: how fast can an SM pull "256 rows x 128 bytes" chunks (the access pattern of the rSVD sketch:
// X is (m x n) row-major, a K-chunk is 32 columns of 256 rows) with (0) register-prefetched LDG.128, (1) 1-D bulk
// async copies (cp.async.bulk, the TMA engine) into a shared-memory ring.  nvcc -arch=sm_100a -O3 rowchunk_read.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH, int FENCE>
__global__ void __launch_bounds__(256) ldg_kernel(const float* __restrict__ X, int m, long long n, float* out, int rows)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nchunk = n / 32;
    const long long per = (nchunk + gridDim.x - 1) / gridDim.x;
    const long long b = blockIdx.x * per;
    const int n_my = (int)(b >= nchunk ? 0 : (nchunk - b < per ? nchunk - b : per));
    const int row_base = blockIdx.y * rows;
    const float* xp = X + (size_t)(row_base + warp * (rows / 8) + (lane >> 3)) * n + 4 * (lane & 7);
    const int passes = rows / 32;
    float4 buf[DEPTH][8];
    float acc = 0.f;
    for (int d = 0; d < DEPTH - 1 && d < n_my; ++d)
        for (int u = 0; u < 8; ++u)
            if (u < passes) buf[d][u] = __ldcs(reinterpret_cast<const float4*>(xp + (b + d) * 32 + (size_t)u * 4 * n));
#pragma unroll 1
    for (int it0 = 0; it0 < n_my; it0 += DEPTH) {
#pragma unroll
        for (int ph = 0; ph < DEPTH; ++ph) {
            const int it = it0 + ph;
            if (it < n_my) {
                const int nx = it + DEPTH - 1;
                if (nx < n_my)
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (u < passes) buf[(ph + DEPTH - 1) % DEPTH][u] = __ldcs(reinterpret_cast<const float4*>(xp + (b + nx) * 32 + (size_t)u * 4 * n));
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (u < passes) acc += buf[ph][u].x + buf[ph][u].y + buf[ph][u].z + buf[ph][u].w;
                if (FENCE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (FENCE == 2) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (FENCE == 3) __threadfence_block();
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// one elected thread issues `rows` bulk copies of 128 bytes per chunk into a STAGES-deep ring; all threads consume
template <int STAGES>
__global__ void __launch_bounds__(288) bulk_kernel(const float* __restrict__ X, int m, long long n, float* out, int rows)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long nchunk = n / 32;
    const long long per = (nchunk + gridDim.x - 1) / gridDim.x;
    const long long b = blockIdx.x * per;
    const int n_my = (int)(b >= nchunk ? 0 : (nchunk - b < per ? nchunk - b : per));
    const int row_base = blockIdx.y * rows;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(256));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [&](uint64_t* bar, uint32_t par) {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(s32(bar)), "r"(par), "r"(100000u) : "memory");
    };
    float acc = 0.f;
    if (warp == 8) {
        // producer warp: lanes issue rows/32 copies each per chunk
        for (int it = 0; it < n_my; ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) wait(&empty[s], (uint32_t)((it / STAGES - 1) & 1));
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(rows * 128) : "memory");
            __syncwarp();
            for (int rr = lane; rr < rows; rr += 32) {
                const float* src = X + (size_t)(row_base + rr) * n + (b + it) * 32;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sm + (size_t)s * rows * 128 + rr * 128)), "l"(src), "r"(128), "r"(s32(&full[s])) : "memory");
            }
        }
    } else {
        for (int it = 0; it < n_my; ++it) {
            const int s = it % STAGES;
            wait(&full[s], (uint32_t)((it / STAGES) & 1));
            const float4* p = reinterpret_cast<const float4*>(sm + (size_t)s * rows * 128);
            for (int i = tid; i < rows * 8; i += 256) { const float4 q = p[i]; acc += q.x + q.y + q.z + q.w; }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void fill_kernel(float* X, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        X[i] = (float)(h & 0xFFFFFF) * (1.0f / 16777216.0f) - 0.5f;
    }
}

int main(int argc, char** argv)
{
    const int m = 512;
    const long long n = 1460000;
    float* X; float* out;
    cudaMalloc(&X, (size_t)m * n * 4); cudaMalloc(&out, 4);
    cudaMemset(X, 0, (size_t)m * n * 4);
    if (argc > 1) fill_kernel<<<4096, 256>>>(X, (size_t)m * n);      // pseudo-random contents instead of zeros
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch) {
        launch(); cudaDeviceSynchronize();
        float best = 1e9f;
        for (int i = 0; i < 5; ++i) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        printf("%-40s %.3f ms  %.0f GB/s  (%s)\n", name, best, 4.0 * m * n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    time("ldg depth2 256rows x 74x2", [&] { ldg_kernel<2, 0><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 256rows x 74x2", [&] { ldg_kernel<3, 0><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth4 256rows x 74x2", [&] { ldg_kernel<4, 0><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 256rows x 148x2 (2 CTA/SM)", [&] { ldg_kernel<3, 0><<<dim3(148, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 256rows x 296x2 (4 CTA/SM)", [&] { ldg_kernel<3, 0><<<dim3(296, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 256rows x 592x2 (8 CTA/SM)", [&] { ldg_kernel<3, 0><<<dim3(592, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 + fence.proxy.async", [&] { ldg_kernel<3, 1><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 + tcgen05.fence::before", [&] { ldg_kernel<3, 2><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    time("ldg depth3 + membar.cta", [&] { ldg_kernel<3, 3><<<dim3(74, 2), 256>>>(X, m, n, out, 256); });
    cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 256 * 128);
    cudaFuncSetAttribute(bulk_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 256 * 128);
    time("bulk 4 stages 256rows x 74x2", [&] { bulk_kernel<4><<<dim3(74, 2), 288, 4 * 256 * 128>>>(X, m, n, out, 256); });
    time("bulk 6 stages 256rows x 74x2", [&] { bulk_kernel<6><<<dim3(74, 2), 288, 6 * 256 * 128>>>(X, m, n, out, 256); });
    return 0;
}
