// api.cu -- library-level entry points: version, thread-local error string, device facts.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "../../include/gp_b200.h"

namespace gp {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int g_sm_count = 0, g_smem_optin = 0, g_cc = 0;
static void probe_device() {
    if (g_sm_count) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int maj = 0, min = 0;
    cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
    g_cc = maj * 10 + min;
}
int sm_count() {
    probe_device();
    return g_sm_count;
}
int max_smem_optin() {
    probe_device();
    return g_smem_optin;
}
}  // namespace gp

extern "C" int gp_version(void) { return 100; }
extern "C" const char* gp_last_error(void) { return gp::g_err; }
extern "C" int gp_device_info(int* out3) {
    GP_REQUIRE(out3 != nullptr, "gp_device_info: null output");
    gp::probe_device();
    GP_REQUIRE(gp::g_sm_count > 0, "gp_device_info: no CUDA device visible");
    out3[0] = gp::g_sm_count;
    out3[1] = gp::g_smem_optin;
    out3[2] = gp::g_cc;
    return 0;
}
