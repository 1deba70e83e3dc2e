// gemm.cu -- strided row-tile GEMM on tcgen05 (gp_gemm), the dense building block of the graph-Transformer path and of
// the "tight" arithmetic mode.
//
//   C(m, n) = [C(m, n) +] [resid(m, n) +] bias[n] + sum_k A(m, k) * B(n, k),  then max(., 0) if relu
//   A(m, k) = a[m*a_sm + k*a_sk],  B(n, k) = b[n*b_sn + k*b_sk],  C(m, n) = c[m*c_sm + n*c_sn]   (element strides)
//
// By choice of strides this is a Linear forward (A = activations, B = weight [out][in]), its dgrad (B = weight read
// transposed) or its wgrad (A = dY read transposed, B = X read transposed, the long contraction cut over CTAs with
// split_k and reduced in fixed order).  Operands are fp32 or bf16 in global memory and are staged into SW128 K-major
// tiles by the threads (16-byte vector path when K is the contiguous axis), one 128 x 128 output tile per CTA,
// 64-wide K chunks, fp32 accumulation in TMEM.
//
// terms = 1: operands rounded to bf16 (one MMA per k-step) -- the Transformer block's projections and gated MLP
//            (graphphysics/models/layers.py:213-278, 637-697, 766-819).
// terms = 3: every fp32 operand split on the fly into THREE bf16 terms, v = hi + mid + lo (24 mantissa bits, i.e. the
//            fp32 value up to its last bit), one product = six MMAs  hi.hi + hi.mid + mid.hi + mid.mid + hi.lo + lo.hi
//            accumulated smallest first -- precision="tight" (graphphysics_b200/tight.py; SURVEY §7 iii).  Measured on
//            the 15-layer, 128-wide model: a two-term split (3 MMAs, 16 bits) reproduces the fp32 reference's OUTPUT to
//            7e-5 but flips ~1e-5 of the ReLU gates, which moves the GRADIENTS by 3e-3..7e-3; the reference's own fp32
//            gradients are only defined to 1e-3..4e-3 (fp32 vs fp64 evaluation of the same modules), so gradient
//            parity needs fp32-grade pre-activations: the three-term split.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tile_util.cuh"

namespace {
using namespace gp;

constexpr int kChunk = 64;                  // K elements staged per step (one SW128 column block)
constexpr int kTileBytes = 128 * 128;       // [128 rows][64 bf16]

// element (r, k) of a [128 x 64] SW128 K-major tile
__device__ __forceinline__ uint32_t tile_off(int r, int k) { return sw128_chunk_off(r, k >> 3) + (k & 7) * 2; }

__device__ __forceinline__ float load_elem(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_elem(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <int TERMS>
__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* mid, uint8_t* lo, uint32_t o, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(hi + o) = h;
    if (TERMS == 3) {
        const float r1 = v - __bfloat162float(h);                      // exact in fp32
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        *reinterpret_cast<__nv_bfloat16*>(mid + o) = m;
        *reinterpret_cast<__nv_bfloat16*>(lo + o) = __float2bfloat16_rn(r1 - __bfloat162float(m));
    }
}

// Stage rows [r0, r0+128) x k [k0, k0+64) of X(r, k) = x[r*sr + k*sk] as bf16 tile(s); out-of-range -> 0.
// vec: K is the contiguous axis and every row start is 16-byte aligned (checked on the host).
template <int TERMS, typename T>
__device__ __forceinline__ void stage_tile(uint8_t* hi, uint8_t* mid, uint8_t* lo, const T* __restrict__ x, long long sr, long long sk,
                                           int r0, int nrows, int k0, int K, int tid, bool vec) {
    if (vec) {
        for (int i = tid; i < 128 * 8; i += 128) {          // one 8-element (16-byte bf16) chunk per step
            const int r = i >> 3, ch = i & 7;
            const int k = k0 + ch * 8;
            float f[8];
            const bool in = (r0 + r < nrows) && (k < K);     // K % 8 == 0 on this path
            if (in) {
                const T* src = x + (long long)(r0 + r) * sr + k;
                if (sizeof(T) == 2) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
                    if (TERMS == 1) {
                        *reinterpret_cast<uint4*>(hi + sw128_chunk_off(r, ch)) = q;
                        continue;
                    }
                    unpack8(q, f);
                } else {
                    const float4 u0 = __ldg(reinterpret_cast<const float4*>(src)), u1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.f;
            }
            const uint32_t o = sw128_chunk_off(r, ch);
            if (TERMS == 1) {
                *reinterpret_cast<uint4*>(hi + o) = pack8(f);
            } else {
                float m[8], l[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = __bfloat162float(__float2bfloat16_rn(f[j]));
                    const float r1 = f[j] - h;
                    m[j] = __bfloat162float(__float2bfloat16_rn(r1));
                    l[j] = r1 - m[j];
                }
                *reinterpret_cast<uint4*>(hi + o) = pack8(f);
                *reinterpret_cast<uint4*>(mid + o) = pack8(m);
                *reinterpret_cast<uint4*>(lo + o) = pack8(l);
            }
        }
        return;
    }
    const bool r_fast = (sr == 1);          // rows contiguous in memory: let consecutive threads walk rows
    for (int i = tid; i < 128 * kChunk; i += 128) {
        const int r = r_fast ? (i & 127) : (i >> 6);
        const int k = r_fast ? (i >> 7) : (i & 63);
        float v = 0.f;
        if (r0 + r < nrows && k0 + k < K) v = load_elem(x + (long long)(r0 + r) * sr + (long long)(k0 + k) * sk);
        split_store<TERMS>(hi, mid, lo, tile_off(r, k), v);
    }
}

template <int TERMS>
__global__ void __launch_bounds__(128, 1) gemm_kernel(const gp_gemm_args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* a_t[3] = {smem, smem + kTileBytes, smem + 2 * kTileBytes};
    uint8_t* b_t[3] = {smem + TERMS * kTileBytes, smem + (TERMS + 1) * kTileBytes, smem + (TERMS + 2) * kTileBytes};
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
    const int n_chunks = (p.K + kChunk - 1) / kChunk;
    const int per_z = (n_chunks + gridDim.z - 1) / gridDim.z;
    const int c_begin = blockIdx.z * per_z, c_end = min(n_chunks, c_begin + per_z);

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = idesc_bf16(128, false, false);
    uint32_t phase = 0;
    const bool a_vec = p.flags & 1, b_vec = p.flags & 2;
    for (int c = c_begin; c < c_end; ++c) {
        if (p.a_bf16) stage_tile<TERMS>(a_t[0], a_t[1], a_t[2], reinterpret_cast<const __nv_bfloat16*>(p.a), p.a_sm, p.a_sk, m0, p.M, c * kChunk, p.K, tid, a_vec);
        else stage_tile<TERMS>(a_t[0], a_t[1], a_t[2], reinterpret_cast<const float*>(p.a), p.a_sm, p.a_sk, m0, p.M, c * kChunk, p.K, tid, a_vec);
        if (p.b_bf16) stage_tile<TERMS>(b_t[0], b_t[1], b_t[2], reinterpret_cast<const __nv_bfloat16*>(p.b), p.b_sn, p.b_sk, n0, p.N, c * kChunk, p.K, tid, b_vec);
        else stage_tile<TERMS>(b_t[0], b_t[1], b_t[2], reinterpret_cast<const float*>(p.b), p.b_sn, p.b_sk, n0, p.N, c * kChunk, p.K, tid, b_vec);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t ah = smem_u32(a_t[0]), am = smem_u32(a_t[1]), al = smem_u32(a_t[2]);
            const uint32_t bh = smem_u32(b_t[0]), bm = smem_u32(b_t[1]), bl = smem_u32(b_t[2]);
            for (int ks = 0; ks < kChunk / 16; ++ks) {
                const uint32_t first = (c > c_begin || ks > 0) ? 1u : 0u;
                if (TERMS == 3) {       // smallest terms first, the hi.hi term last
                    mma_ss(tmem, desc_kmajor(al, 128, ks), desc_kmajor(bh, 128, ks), idesc, first);
                    mma_ss(tmem, desc_kmajor(ah, 128, ks), desc_kmajor(bl, 128, ks), idesc, 1u);
                    mma_ss(tmem, desc_kmajor(am, 128, ks), desc_kmajor(bm, 128, ks), idesc, 1u);
                    mma_ss(tmem, desc_kmajor(am, 128, ks), desc_kmajor(bh, 128, ks), idesc, 1u);
                    mma_ss(tmem, desc_kmajor(ah, 128, ks), desc_kmajor(bm, 128, ks), idesc, 1u);
                    mma_ss(tmem, desc_kmajor(ah, 128, ks), desc_kmajor(bh, 128, ks), idesc, 1u);
                } else {
                    mma_ss(tmem, desc_kmajor(ah, 128, ks), desc_kmajor(bh, 128, ks), idesc, first);
                }
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);        // the staged tiles have been read: the next chunk may overwrite them
        phase ^= 1;
        tc_fence_after();
    }
    // epilogue: lane == row of the tile
    const int m = m0 + tid;
    const uint32_t tl = tmem_addr(tmem, (tid >> 5) * 32, 0);
    const bool direct = (gridDim.z == 1);
    float* const cf = reinterpret_cast<float*>(p.c);
    __nv_bfloat16* const cb = reinterpret_cast<__nv_bfloat16*>(p.c);
    for (int cc = 0; cc < 128 && n0 + cc < p.N; cc += 16) {
        uint32_t v[16];
        if (c_end > c_begin) {
            tmem_ld16(tl + cc, v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (m >= p.M) continue;
        if (!direct) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + cc + j < p.N) p.partials[((size_t)blockIdx.z * p.M + m) * p.N + n0 + cc + j] = __uint_as_float(v[j]);
            continue;
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = n0 + cc + j;
            f[j] = __uint_as_float(v[j]);
            if (n < p.N) {
                if (p.bias) f[j] += p.bias[n];
                if (p.resid) f[j] += p.resid[(long long)m * p.c_sm + (long long)n * p.c_sn];
                if (p.accumulate) f[j] += p.c_bf16 ? __bfloat162float(cb[(long long)m * p.c_sm + (long long)n * p.c_sn]) : cf[(long long)m * p.c_sm + (long long)n * p.c_sn];
                if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
        }
        const bool full = (n0 + cc + 16 <= p.N) && (p.flags & 4);      // contiguous, 16-byte aligned output rows
        if (full && p.c_bf16) {
            uint4* d = reinterpret_cast<uint4*>(cb + (long long)m * p.c_sm + n0 + cc);
            d[0] = pack8(f);
            d[1] = pack8(f + 8);
        } else if (full) {
            float4* d = reinterpret_cast<float4*>(cf + (long long)m * p.c_sm + n0 + cc);
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + cc + j;
                if (n < p.N) {
                    if (p.c_bf16) cb[(long long)m * p.c_sm + (long long)n * p.c_sn] = __float2bfloat16_rn(f[j]);
                    else cf[(long long)m * p.c_sm + (long long)n * p.c_sn] = f[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 128);
}

// C(m, n) = [C(m, n)] + [resid] + bias[n] + sum_z partials[z][m][n], z ascending (bit-reproducible)
__global__ void gemm_reduce_kernel(const gp_gemm_args p, int nz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.M * p.N) return;
    const int m = (int)(i / p.N), n = (int)(i - (long long)m * p.N);
    float f = 0.f;
    for (int z = 0; z < nz; ++z) f += p.partials[((size_t)z * p.M + m) * p.N + n];
    const long long o = (long long)m * p.c_sm + (long long)n * p.c_sn;
    if (p.bias) f += p.bias[n];
    if (p.resid) f += p.resid[o];
    float* cf = reinterpret_cast<float*>(p.c);
    __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.c);
    if (p.accumulate) f += p.c_bf16 ? __bfloat162float(cb[o]) : cf[o];
    if (p.relu) f = fmaxf(f, 0.f);
    if (p.c_bf16) cb[o] = __float2bfloat16_rn(f);
    else cf[o] = f;
}
}  // namespace

extern "C" int gp_gemm(const gp_gemm_args* args, void* stream) {
    GP_REQUIRE(args != nullptr, "gp_gemm: null args");
    gp_gemm_args a = *args;
    GP_REQUIRE(a.M > 0 && a.N > 0 && a.K >= 0, "gp_gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
    GP_REQUIRE(a.a && a.b && a.c, "gp_gemm: null operand");
    GP_REQUIRE(a.terms == 1 || a.terms == 3, "gp_gemm: terms must be 1 (bf16 operands) or 3 (three-term split)");
    GP_REQUIRE(a.split_k >= 1 && a.split_k <= 1024, "gp_gemm: split_k must be in [1, 1024]");
    GP_REQUIRE(a.split_k == 1 || a.partials != nullptr, "gp_gemm: split_k > 1 needs a partials buffer of split_k*M*N floats");
    GP_REQUIRE(!(a.resid && a.c_bf16), "gp_gemm: resid needs an fp32 output");
    // vector paths: K contiguous + 16-byte aligned rows (operands), N contiguous + aligned rows (output)
    auto aligned = [](const void* base, long long stride_elems, int elem) {
        return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (stride_elems * elem) % 16 == 0;
    };
    a.flags = 0;
    if (a.a_sk == 1 && a.K % 8 == 0 && aligned(a.a, a.a_sm, a.a_bf16 ? 2 : 4)) a.flags |= 1;
    if (a.b_sk == 1 && a.K % 8 == 0 && aligned(a.b, a.b_sn, a.b_bf16 ? 2 : 4)) a.flags |= 2;
    if (a.c_sn == 1 && aligned(a.c, a.c_sm, a.c_bf16 ? 2 : 4)) a.flags |= 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * a.terms * kTileBytes + 1024;
    static bool attr_set[2] = {false, false};
    if (!attr_set[a.terms == 3]) {
        if (a.terms == 3) GP_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else GP_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[a.terms == 3] = true;
    }
    const dim3 grid((a.M + 127) / 128, (a.N + 127) / 128, a.split_k);
    if (a.terms == 3) gemm_kernel<3><<<grid, 128, smem, st>>>(a);
    else gemm_kernel<1><<<grid, 128, smem, st>>>(a);
    GP_CHECK_CUDA(cudaGetLastError());
    if (a.split_k > 1) {
        const long long total = (long long)a.M * a.N;
        gemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, a.split_k);
        GP_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}
