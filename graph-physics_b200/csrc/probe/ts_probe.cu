// ts_probe.cu -- bring-up test for tcgen05.mma with the A operand in tensor memory ("TS" form): the hidden
// activations of an MLP can stay in TMEM between layers (bf16 pairs packed into 32-bit columns, lane = row)
// instead of going through shared memory.  Checks D = A . B^T exactly (integer-valued bf16 inputs) against the
// SS form and times a chain of 8 (K = 128) instructions in both forms.   Run: ts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../tc5.cuh"
using namespace tc5;

constexpr int K = 128, N = 128;

__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128, 1) ts_kernel(const __nv_bfloat16* A /*[128][K]*/, const __nv_bfloat16* W /*[N][K]*/,
                                                    float* out_ss, float* out_ts, long long* cycles /*[2]*/) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    uint8_t* a_t = smem;               // [128][128] bf16 SW128 K-major: 32 KB
    uint8_t* b_t = smem + 32768;       // [128][128]: 32 KB
    for (int i = tid; i < 128 * 16; i += 128) {
        const int r = i >> 4, ch = i & 15;
        *reinterpret_cast<uint4*>(a_t + sw128_off(128, r, ch * 8)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + ch * 8);
        *reinterpret_cast<uint4*>(b_t + sw128_off(128, r, ch * 8)) = *reinterpret_cast<const uint4*>(W + (size_t)r * K + ch * 8);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (tid < 32) tmem_alloc(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tl = tmem_addr(tmem, (tid >> 5) * 32, 0);
    // A -> TMEM columns 256..319: thread `tid` owns lane `tid`; column j holds elements (2j, 2j+1) of its row
    for (int c = 0; c < K / 2; c += 16) {
        uint32_t v[16];
        for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + 2 * (c + j));
        tmem_st16(tl + 256 + c, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t idesc = idesc_bf16(N, false, false);
    for (int form = 0; form < 2; ++form) {            // 0: SS into columns 0..127, 1: TS into 128..255
        long long t0 = 0;
        if (tid == 0) {
            tc_fence_after();
            t0 = clock64();
            for (int rep = 0; rep < 16; ++rep)
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint32_t acc = (rep > 0 || ks > 0) ? 1u : 0u;
                    if (form == 0) mma_ss(tmem, desc_kmajor(smem_u32(a_t), 128, ks), desc_kmajor(smem_u32(b_t), 128, ks), idesc, acc);
                    else mma_ts(tmem + 128, tmem + 256 + ks * 8, desc_kmajor(smem_u32(b_t), 128, ks), idesc, acc);
                }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (tid == 0) cycles[form] = clock64() - t0;
        float* o = form == 0 ? out_ss : out_ts;
        for (int c = 0; c < N; c += 16) {
            uint32_t v[16];
            tmem_ld16(tl + form * 128 + c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) o[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
    }
    if (tid < 32) tmem_dealloc(tmem, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    std::vector<__nv_bfloat16> A(128 * K), W(N * K);
    std::vector<float> Af(128 * K), Wf(N * K);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((int)(i * 7 % 5) - 2); A[i] = __float2bfloat16(Af[i]); }
    for (size_t i = 0; i < W.size(); ++i) { Wf[i] = (float)((int)(i * 5 % 3) - 1); W[i] = __float2bfloat16(Wf[i]); }
    __nv_bfloat16 *dA, *dW; float *dS, *dT; long long* dC;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dS, 128 * N * 4)); CK(cudaMalloc(&dT, 128 * N * 4));
    CK(cudaMalloc(&dC, 16));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    ts_kernel<<<1, 128, 72 * 1024>>>(dA, dW, dS, dT, dC);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    std::vector<float> S(128 * N), T(128 * N); long long cyc[2];
    CK(cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
    int bad_ss = 0, bad_ts = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += Af[r * K + k] * Wf[n * K + k];
            ref *= 16.f;                                  // 16 accumulated repetitions
            bad_ss += S[r * N + n] != ref;
            bad_ts += T[r * N + n] != ref;
        }
    printf("SS (A, B in shared memory): %s (%d wrong), %lld cycles for 128 instructions = %.1f per instruction\n", bad_ss ? "FAIL" : "PASS", bad_ss, cyc[0], cyc[0] / 128.0);
    printf("TS (A in tensor memory)   : %s (%d wrong), %lld cycles for 128 instructions = %.1f per instruction\n", bad_ts ? "FAIL" : "PASS", bad_ts, cyc[1], cyc[1] / 128.0);
    return (bad_ss || bad_ts) ? 3 : 0;
}
