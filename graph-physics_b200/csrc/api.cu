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
static bool g_launch_overlap = false;
bool launch_overlap() { return g_launch_overlap; }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    // the driver entry point is fetched through the runtime, so the library links against cudart only
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

bool tma_map_2d(CUtensorMap* map, const void* base, long long rows, int cols, long long ld) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || !base || rows <= 0 || cols <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld * 2) % 16 != 0 || ld < cols) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    *map = m;
    return true;
}
// Row-gather map for cp.async.bulk.tensor.2d ... tile::gather4: box = 64 columns x ONE row (a {64, 4} box faults; probe/
// gather4_probe.cu), each instruction names four row coordinates and lands 4 x 128 bytes in the SW128 tile layout.
bool tma_map_rows(CUtensorMap* map, const void* base, long long rows, int cols, long long ld) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || !base || rows <= 0 || cols <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld * 2) % 16 != 0 || ld < cols) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, 1}, estr[2] = {1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    *map = m;
    return true;
}
}  // namespace gp

extern "C" int gp_version(void) { return 200; }
// sizeof of the argument structs as this library was compiled: a binding checks its own layout against it
extern "C" int gp_sizeof_struct(const char* name) {
    if (!name) return -1;
#define GP_SZ(T) if (!strcmp(name, #T)) return (int)sizeof(T);
    GP_SZ(gp_mlp_fwd_args) GP_SZ(gp_mlp_bwd_args) GP_SZ(gp_linear_bwd_args) GP_SZ(gp_pack_entry) GP_SZ(gp_reduce_seg)
    GP_SZ(gp_attention_args) GP_SZ(gp_gemm_args)
#undef GP_SZ
    return -1;
}
extern "C" int gp_set_launch_overlap(int enabled) {
    const int old = gp::g_launch_overlap ? 1 : 0;
    gp::g_launch_overlap = enabled != 0;
    return old;
}
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
