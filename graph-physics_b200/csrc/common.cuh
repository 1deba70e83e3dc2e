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

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

}  // namespace gp
