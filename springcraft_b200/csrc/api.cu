// status strings / error bookkeeping of the C ABI
#include <string.h>

#include "common.cuh"

namespace scb {
static thread_local char g_last_error[512] = "";
void set_last_cuda_error(cudaError_t e, const char* file, int line) {
    snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e),
             file, line);
    (void)cudaGetLastError();   // clear the (non-sticky) runtime error so that the next launch check starts clean
}
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
}  // namespace scb

extern "C" uint64_t scb_launch_count(void) { return __atomic_load_n(&scb::g_launches, __ATOMIC_RELAXED); }

extern "C" int scb_version(void) { return SCB_VERSION; }

extern "C" const char* scb_last_cuda_error(void) { return scb::g_last_error; }

extern "C" const char* scb_status_string(int status) {
    switch (status) {
        case SCB_OK: return "ok";
        case SCB_ERR_INVALID: return "invalid argument";
        case SCB_ERR_CUDA: return "CUDA error";
        case SCB_ERR_ABOVE_CUTOFF:
            return "Atom interactions above cutoff distance are not allowed in TabulatedForceField";
        case SCB_ERR_WORKSPACE: return "workspace too small";
        case SCB_ERR_NOT_CONVERGED: return "eigensolver did not converge";
        case SCB_ERR_UNSUPPORTED: return "size or mode not supported";
    }
    return "unknown status";
}
