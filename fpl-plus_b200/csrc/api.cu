// Error reporting and device queries for the C ABI (include/fplplus_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/fplplus_b200.h"

static thread_local char g_err[1024] = "";

void fpl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fpl_last_error(void) { return g_err; }

unsigned long long g_fpl_launches = 0;

// programmatic dependent launch on by default; FPL_PDL=0 falls back to plain stream-ordered launches (A/B measurements)
static int read_pdl_env() {
    const char* e = getenv("FPL_PDL");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
}
int g_fpl_pdl = read_pdl_env();

extern "C" long long fpl_launch_count(int reset) {
    unsigned long long v = __atomic_load_n(&g_fpl_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_fpl_launches, 0ull, __ATOMIC_RELAXED);
    return (long long)v;
}

int g_fpl_num_sms = 148;

extern "C" int fpl_set_sm_budget(int sms) {
    if (sms < 16 || sms > 148) {
        fpl_set_error("fpl_set_sm_budget: %d not in [16, 148]", sms);
        return 2;
    }
    g_fpl_num_sms = sms;
    return 0;
}

extern "C" int fpl_version(void) { return 200; }

extern "C" int fpl_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
