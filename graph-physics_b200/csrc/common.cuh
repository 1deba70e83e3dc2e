// common.cuh -- host-side error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace gp {

void set_error(const char* fmt, ...);
int sm_count();
int max_smem_optin();

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
