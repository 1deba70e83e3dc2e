// cta2_probe.cu -- bring-up test for CTA-pair MMAs (tcgen05 cta_group::2, M = 256), the building block of
// the next kernel generation: each CTA of a 2-CTA cluster stages its own 128 rows of A
// and ITS HALF of the weight rows (N/2), the leader issues one M=256 instruction per k-step for both,
// the commit is multicast to both CTAs' mbarriers, and each CTA reads its 128 x N block out of its own
// TMEM.  Checks the result exactly (integer-valued bf16 inputs) and reports which half of W each CTA
// has to hold.   Run: cta2_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../tc5.cuh"
using namespace tc5;

constexpr int K = 64, N = 128;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_kernel(const __nv_bfloat16* A /*[256][K]*/, const __nv_bfloat16* W /*[N][K]*/, float* out /*[256][N]*/, int swap_halves) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = cluster_rank();
    const int tid = threadIdx.x;
    uint8_t* a_t = smem;              // [128 rows][64] bf16, SW128 K-major: 16 KB
    uint8_t* b_t = smem + 16384;      // [64 rows][64]  bf16: 8 KB
    const uint32_t half = swap_halves ? (rank ^ 1u) : rank;
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(a_t + sw128_chunk_off(r, ch)) =
            *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + r) * K + ch * 8);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(b_t + sw128_chunk_off(r, ch)) =
            *reinterpret_cast<const uint4*>(W + (size_t)(half * 64 + r) * K + ch * 8);
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync();                   // both CTAs' tiles and barriers are ready
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (rank == 0 && tid < 32 && elect_one()) {
        const uint32_t idesc = idesc_bf16(N, false, false, 256);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = desc_kmajor(smem_u32(a_t), 128, ks), bd = desc_kmajor(smem_u32(b_t), 64, ks);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&mma_bar)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    const uint32_t tl = tmem_addr(tmem, (tid >> 5) * 32, 0);
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tl + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(size_t)(rank * 128 + tid) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    std::vector<__nv_bfloat16> A(256 * K), W(N * K);
    std::vector<float> Af(256 * K), Wf(N * K);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((int)(i * 7 % 13) - 6); A[i] = __float2bfloat16(Af[i]); }
    for (size_t i = 0; i < W.size(); ++i) { Wf[i] = (float)((int)(i * 5 % 11) - 5); W[i] = __float2bfloat16(Wf[i]); }
    __nv_bfloat16 *dA, *dW; float* dO;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dO, 256 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    int fails = 0;
    for (int swap = 0; swap < 2; ++swap) {
        CK(cudaMemset(dO, 0xFF, 256 * N * 4));
        cta2_kernel<<<2, 128, 32 * 1024>>>(dA, dW, dO, swap);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        std::vector<float> O(256 * N);
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0, bad_lo = 0, bad_hi = 0;
        for (int r = 0; r < 256; ++r)
            for (int n = 0; n < N; ++n) {
                float ref = 0.f;
                for (int k = 0; k < K; ++k) ref += Af[r * K + k] * Wf[n * K + k];
                if (O[r * N + n] != ref) { ++bad; (n < 64 ? bad_lo : bad_hi)++; }
            }
        printf("cta_group::2 M=256 N=%d K=%d, CTA r holds W rows [%s*64, +64): %s (%d wrong: %d in cols 0..63, %d in cols 64..127)\n", N, K,
               swap ? "(1-r)" : "r", bad ? "FAIL" : "PASS", bad, bad_lo, bad_hi);
        if (swap == 0) fails = bad;
    }
    return fails ? 3 : 0;
}
