// tc5_probe.cu -- standalone bring-up test for the tcgen05 layer in tc5.cuh.
// One CTA computes D[128 x N] = (C +) A * B with every operand-major combination the
// library uses and checks the result exactly against a host reference (inputs are small
// dyadic rationals, so fp32 accumulation is exact).  Run:  tc5_probe <case-index>
//   (each case in its own process so a device trap cannot mask later cases).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../tc5.cuh"

using namespace tc5;

struct Case {
    const char* name;
    int N, K;       // MMA N, reduction length
    int a_mn;       // 0: A given as [128][K] (K-major); 1: A given as [K][am_valid] (MN-major)
    int am_valid;   // valid M columns when a_mn (128 / 64 / 32); rows >= am_valid of D are don't-care
    int b_mn;       // 0: B given as [N][K]; 1: B given as [K][N]
    int use_init;   // 1: accumulator pre-loaded from C via tcgen05.st
    int lbo_zero;   // a_mn && am_valid<128: 1 = alias the missing MN block onto block 0 (LBO = 0)
};

static const Case kCases[] = {
    {"tmem st/ld roundtrip (K=0)", 128, 0, 0, 128, 0, 1, 0},
    {"K-major A,B  N=128 K=64", 128, 64, 0, 128, 0, 0, 0},
    {"K-major A,B  N=128 K=128", 128, 128, 0, 128, 0, 0, 0},
    {"K-major A,B  N=64  K=64", 64, 64, 0, 128, 0, 0, 0},
    {"K-major A,B  N=32  K=32", 32, 32, 0, 128, 0, 0, 0},
    {"K-major A,B  N=16  K=16", 16, 16, 0, 128, 0, 0, 0},
    {"K-major A,B  N=16  K=128", 16, 128, 0, 128, 0, 0, 0},
    {"K-major A,B  N=128 K=384", 128, 384, 0, 128, 0, 0, 0},
    {"K-major A,B  N=256 K=128", 256, 128, 0, 128, 0, 0, 0},
    {"K-major A,B  N=192 K=64", 192, 64, 0, 128, 0, 0, 0},
    {"init + K-major N=128 K=128", 128, 128, 0, 128, 0, 1, 0},
    {"dgrad: A K-major, B MN-major N=128 K=128", 128, 128, 0, 128, 1, 0, 0},
    {"dgrad: A K-major, B MN-major N=64 K=64", 64, 64, 0, 128, 1, 0, 0},
    {"dgrad: A K-major, B MN-major N=32 K=32", 32, 32, 0, 128, 1, 0, 0},
    {"dgrad: A K-major, B MN-major N=64 K=192", 64, 192, 0, 128, 1, 0, 0},
    {"dgrad: A K-major, B MN-major N=16 K=128", 16, 128, 0, 128, 1, 0, 0},
    {"wgrad: A,B MN-major M=128 N=128 K=128", 128, 128, 1, 128, 1, 0, 0},
    {"wgrad: A,B MN-major M=128 N=256 K=128", 256, 128, 1, 128, 1, 0, 0},
    {"wgrad: A,B MN-major M=128 N=16 K=64", 16, 64, 1, 128, 1, 0, 0},
    {"wgrad: A MN-major Mvalid=64 (zero pad block) N=64 K=128", 64, 128, 1, 64, 1, 0, 0},
    {"wgrad: A MN-major Mvalid=64 (LBO=0 alias) N=64 K=128", 64, 128, 1, 64, 1, 0, 1},
    {"wgrad: A MN-major Mvalid=32 (LBO=0 alias) N=32 K=128", 32, 128, 1, 32, 1, 0, 1},
    {"wgrad: A MN-major Mvalid=16 (LBO=0 alias) N=128 K=128", 128, 128, 1, 16, 1, 0, 1},
};
static const int kNumCases = sizeof(kCases) / sizeof(kCases[0]);

// Stage a row-major [rows][cols] bf16 global matrix into an SW128 row tile (cols % 8 == 0).
__device__ void stage_tile(uint8_t* tile, const __nv_bfloat16* g, int rows, int cols, int tile_rows) {
    int chunks_per_row = cols / 8;
    for (int i = threadIdx.x; i < rows * chunks_per_row; i += blockDim.x) {
        int r = i / chunks_per_row, ch = i % chunks_per_row;
        uint4 v = *reinterpret_cast<const uint4*>(g + (size_t)r * cols + ch * 8);
        *reinterpret_cast<uint4*>(tile + sw128_off(tile_rows, r, ch * 8)) = v;
    }
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* A_g, const __nv_bfloat16* B_g, const float* C_g, float* D_g, Case cs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;

    const int N = cs.N, K = cs.K;
    const int rows_a = cs.a_mn ? K : 128;
    const int cols_a = cs.a_mn ? cs.am_valid : K;
    const int rows_b = cs.b_mn ? K : N;
    const int cols_b = cs.b_mn ? N : K;
    // A tile always reserves 2 MN blocks when MN-major so M=128 has somewhere to read.
    const int blocks_a = cs.a_mn ? 2 : (cols_a + 63) / 64;
    uint8_t* a_tile = smem;
    uint8_t* b_tile = smem + (size_t)blocks_a * rows_a * 128;

    // zero everything we might read (keeps don't-care rows finite)
    {
        size_t total = (size_t)blocks_a * rows_a * 128 + (size_t)((cols_b + 63) / 64) * rows_b * 128;
        for (size_t i = threadIdx.x; i < total / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (K > 0) {
        stage_tile(a_tile, A_g, rows_a, cols_a, rows_a);
        stage_tile(b_tile, B_g, rows_b, cols_b, rows_b);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t row = threadIdx.x;   // == TMEM lane
    const uint32_t tlane = tmem_addr(tmem, warp * 32, 0);

    if (cs.use_init) {
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(C_g[(size_t)row * N + c + j]);
            tmem_st8(tlane + c, v);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_bf16(N, cs.a_mn != 0, cs.b_mn != 0);
        const uint32_t a_s = smem_u32(a_tile), b_s = smem_u32(b_tile);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t ad = cs.a_mn ? desc_mnmajor(a_s, rows_a, ks, 0, cs.lbo_zero ? 0u : 0xFFFFFFFFu)
                                  : desc_kmajor(a_s, 128, ks);
            uint64_t bd = cs.b_mn ? desc_mnmajor(b_s, rows_b, ks) : desc_kmajor(b_s, N, ks);
            mma_ss(tmem, ad, bd, idesc, (ks > 0 || cs.use_init) ? 1u : 0u);
        }
        mma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    tc_fence_after();

    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tmem_ld8(tlane + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) D_g[(size_t)row * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
    (void)lane;
}

static float rnd_val(uint32_t& s) {   // dyadic values in [-1, 1], exactly representable in bf16
    s = s * 1664525u + 1013904223u;
    return (float)((int)((s >> 16) % 9) - 4) / 4.0f;
}

#define CK(x)                                                                           \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) {
        printf("%d\n", kNumCases);
        return 0;
    }
    int idx = atoi(argv[1]);
    if (idx < 0 || idx >= kNumCases) return 1;
    Case cs = kCases[idx];
    const int N = cs.N, K = cs.K;
    const int rows_a = cs.a_mn ? K : 128, cols_a = cs.a_mn ? cs.am_valid : K;
    const int rows_b = cs.b_mn ? K : N, cols_b = cs.b_mn ? N : K;
    std::vector<float> A((size_t)std::max(1, rows_a * cols_a)), B((size_t)std::max(1, rows_b * cols_b)), C(128 * N);
    uint32_t seed = 1234u + idx;
    for (auto& v : A) v = rnd_val(seed);
    for (auto& v : B) v = rnd_val(seed);
    for (auto& v : C) v = rnd_val(seed) * 8.0f;
    std::vector<__nv_bfloat16> Ab(A.size()), Bb(B.size());
    for (size_t i = 0; i < A.size(); ++i) Ab[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); ++i) Bb[i] = __float2bfloat16(B[i]);

    __nv_bfloat16 *dA, *dB;
    float *dC, *dD;
    CK(cudaMalloc(&dA, Ab.size() * 2));
    CK(cudaMalloc(&dB, Bb.size() * 2));
    CK(cudaMalloc(&dC, C.size() * 4));
    CK(cudaMalloc(&dD, C.size() * 4));
    CK(cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, C.size() * 4));
    const int smem_bytes = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dC, dD, cs);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(C.size());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));

    const int m_valid = cs.a_mn ? cs.am_valid : 128;
    double max_err = 0;
    int bad = 0;
    for (int m = 0; m < m_valid; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = cs.use_init ? C[(size_t)m * N + n] : 0.0;
            for (int k = 0; k < K; ++k) {
                double a = cs.a_mn ? A[(size_t)k * cols_a + m] : A[(size_t)m * K + k];
                double b = cs.b_mn ? B[(size_t)k * N + n] : B[(size_t)n * K + k];
                ref += a * b;
            }
            double err = fabs(ref - (double)D[(size_t)m * N + n]);
            if (!(err <= 1e-6)) {
                if (bad < 4) printf("   mismatch m=%d n=%d ref=%g got=%g\n", m, n, ref, D[(size_t)m * N + n]);
                ++bad;
            }
            if (err > max_err || err != err) max_err = err;
        }
    printf("[%2d] %-58s %s  (max_err=%g, bad=%d)\n", idx, cs.name, bad ? "FAIL" : "PASS", max_err, bad);
    return bad ? 3 : 0;
}
