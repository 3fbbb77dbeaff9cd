// Library-level entry points: version, error string, device query.
#include "ggp_common.cuh"
#include "../../include/gladsgp_b200.h"
#include <cstdarg>

namespace ggp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return GGP_ERR_CUDA;
}

}  // namespace ggp

extern "C" {

int ggp_version(void) { return GGP_VERSION; }

const char* ggp_last_error_string(void) { return ggp::g_err; }

int ggp_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* smem_optin)
{
    int dev = 0;
    GGP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GGP_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = (long long)p.sharedMemPerBlockOptin;
    return GGP_OK;
}

}  // extern "C"
