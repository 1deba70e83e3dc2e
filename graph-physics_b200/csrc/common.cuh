// common.cuh -- host-side error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace gp {

void set_error(const char* fmt, ...);
int sm_count();
int max_smem_optin();

// TMA descriptor of a row-major bf16 matrix [rows x cols] with leading dimension `ld` (elements):
// box = 64 columns x 128 rows, SWIZZLE_128B, i.e. one instruction moves one 64-column block of an
// SW128 row tile; out-of-range rows / columns read as zero and are clipped on store.  Returns
// false (map untouched) when the tensor cannot be described (alignment) -- callers then use
// their per-thread copy path.
bool tma_map_2d(CUtensorMap* map, const void* base, long long rows, int cols, long long ld);
// the same matrix described for row gathers (tile::gather4): box = 64 columns x 1 row
bool tma_map_rows(CUtensorMap* map, const void* base, long long rows, int cols, long long ld);

// Launch-overlap switch (gp_set_launch_overlap): when on, the persistent MLP kernels are launched with
// programmatic stream serialization, so their prologue (weight staging, TMEM allocation) overlaps the tail
// of the previous kernel in the stream.
bool launch_overlap();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = launch_overlap() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

#define GP_CHECK_CUDA(expr)                                                                    \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            gp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

#define GP_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            gp::set_error(__VA_ARGS__); \
            return -1;                 \
        }                              \
    } while (0)

// 1024-byte aligned view of the dynamic shared memory array.  Written as an offset from the
// __shared__ symbol (not an integer round-trip) so the compiler keeps the shared address space and
// emits LDS/STS instead of generic LD/ST for everything carved out of it.
#define GP_SMEM_ALIGNED(raw) ((raw) + ((1024u - (static_cast<uint32_t>(__cvta_generic_to_shared(raw)) & 1023u)) & 1023u))

}  // namespace gp
