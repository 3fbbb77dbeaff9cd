// Micro-benchmark: L2 -> SM read bandwidth on sm_100a (buffer resident in L2, 16-byte loads, every SM reads the whole buffer).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_read l2_read.cu && ./l2_read
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) rd(const double2* __restrict__ p, size_t n, int reps, double* out)
{
    double s = 0.0;
    for (int r = 0; r < reps; ++r) {
        // each CTA starts at a different offset so that the SMs do not hit the same slice at the same time
        size_t off = ((size_t)blockIdx.x * 7919 + (size_t)r * 104729) % n;
        for (size_t i = threadIdx.x; i < n; i += 4 * blockDim.x) {
            size_t a = (off + i) % n, b = (off + i + blockDim.x) % n, c = (off + i + 2 * blockDim.x) % n, d = (off + i + 3 * blockDim.x) % n;
            double2 v0 = __ldcg(p + a), v1 = __ldcg(p + b), v2 = __ldcg(p + c), v3 = __ldcg(p + d);
            s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
        }
    }
    if (s == 1.2345e300) out[0] = s;
}
int main()
{
    for (size_t mb : {16, 32, 64, 96, 256, 1024}) {
        size_t bytes = mb << 20, n = bytes / sizeof(double2);
        double2* p; double* out;
        cudaMalloc(&p, bytes); cudaMalloc(&out, 8);
        cudaMemset(p, 0, bytes);
        int reps = (int)(4096 / mb) + 1;
        if (reps > 64) reps = 64;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int ctas : {1, 2, 4}) {
            rd<<<148 * ctas, 512>>>(p, n, 1, out);
            cudaEventRecord(e0);
            rd<<<148 * ctas, 512>>>(p, n, reps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double tb = (double)bytes * reps * 148 * ctas / ms / 1e9;
            printf("buffer %5zu MB  %d CTAs/SM x 512 thr: %.2f TB/s delivered to the SMs\n", mb, ctas, tb);
        }
        cudaFree(p); cudaFree(out);
    }
    return 0;
}
