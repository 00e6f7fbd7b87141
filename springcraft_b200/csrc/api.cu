// status strings / error bookkeeping of the C ABI
#include <string.h>

#include "resident.cuh"

namespace scb {
static thread_local char g_last_error[512] = "";
void set_last_cuda_error(cudaError_t e, const char* file, int line) {
    snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e),
             file, line);
    (void)cudaGetLastError();   // clear the (non-sticky) runtime error so that the next launch check starts clean
}
static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

static int g_profile_on = 0;
static double g_prof_ms = 0.0;
static long long g_prof_launches = 0;
static unsigned long long g_prof_apps = 0;
bool profile_enabled() { return __atomic_load_n(&g_profile_on, __ATOMIC_RELAXED) != 0; }
void profile_add(double filter_ms, long long launches, unsigned long long applications) {
    g_prof_ms += filter_ms;
    g_prof_launches += launches;
    g_prof_apps += applications;
}
}  // namespace scb

extern "C" int scb_profile(int enable, double* filter_ms, int64_t* filter_launches, int64_t* filter_applications) {
    if (filter_ms) *filter_ms = scb::g_prof_ms;
    if (filter_launches) *filter_launches = scb::g_prof_launches;
    if (filter_applications) *filter_applications = (int64_t)scb::g_prof_apps;
    scb::g_prof_ms = 0.0;
    scb::g_prof_launches = 0;
    scb::g_prof_apps = 0;
    if (enable >= 0) __atomic_store_n(&scb::g_profile_on, enable != 0, __ATOMIC_RELAXED);
    return SCB_OK;
}

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
