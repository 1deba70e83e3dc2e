// gemm.cu -- strided row-tile GEMM on tcgen05 (gp_gemm), the dense building block of the graph-Transformer path and of
// the "tight" arithmetic mode.
//
//   C(m, n) = [C(m, n) +] [resid(m, n) +] bias[n] + sum_k A(m, k) * B(n, k),  then max(., 0) if relu
//   A(m, k) = a[m*a_sm + k*a_sk],  B(n, k) = b[n*b_sn + k*b_sk],  C(m, n) = c[m*c_sm + n*c_sn]   (element strides)
//
// By choice of strides this is a Linear forward (A = activations, B = weight [out][in]), its dgrad (B = weight read
// transposed) or its wgrad (A = dY read transposed, B = X read transposed, the long contraction cut over CTAs with
// split_k and reduced in fixed order).  Operands are fp32 or bf16 in global memory and are staged into SW128 K-major
// tiles by the threads -- 16-byte vector loads into a K-major tile when K is the contiguous axis, into an MN-major tile
// (same SW128 bytes, transposed view in the descriptors) when M / N is: no operand is ever transposed in memory --
// one 128 x 128 output tile per CTA, 64-wide K chunks, fp32 accumulation in TMEM.
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

#include <type_traits>

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

// Vector path: stage the [TR x TC] block  X(r0 + r, c0 + c) = x[(r0 + r) * sr + (c0 + c)]  (c contiguous in memory, rows
// 16-byte aligned -- checked on the host) as an SW128 row tile with TR rows; out-of-range -> 0.  TR = 128, TC = 64: a
// K-major operand (rows = M or N index, columns = K).  TR = 64, TC = 128: an MN-major operand (rows = K index, columns =
// M or N index) -- the transposed reads of wgrad / dgrad without a transposing copy (tc5.cuh: the same bytes serve both
// views).  ncols % 8 == 0 on this path.
// ones_row / ones_col >= 0: the operand has one more row / column than the memory behind it, holding 1.0 (the bias
// gradient as one extra column of a weight gradient).
template <int TERMS, int TR, int TC, typename T, int UNROLL>
__device__ __forceinline__ void stage_vec(uint8_t* hi, uint8_t* mid, uint8_t* lo, const T* __restrict__ x, long long sr, int r0, int nrows,
                                          int c0, int ncols, int tid, int ones_row = -1, int ones_col = -1) {
    constexpr int CPR = TC / 8;                              // 16-byte bf16 chunks per row
    static_assert(TR * CPR == 1024, "eight chunks per thread");
#pragma unroll(UNROLL)
    for (int it = 0; it < 8; ++it) {                         // UNROLL = 8: all eight loads of a thread are in flight together
        const int i = tid + it * 128;
        const int r = i / CPR, ch = i - r * CPR;
        const int c = c0 + ch * 8;
        const uint32_t o = sw128_off(TR, r, ch * 8);
        float f[8];
        if (r0 + r == ones_row && c < ncols) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 1.f;
        } else if (c == ones_col && r0 + r < nrows) {
            f[0] = 1.f;
#pragma unroll
            for (int j = 1; j < 8; ++j) f[j] = 0.f;
        } else if ((r0 + r < nrows) && (c < ncols)) {
            const T* src = x + (long long)(r0 + r) * sr + c;
            if (sizeof(T) == 2) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
                if (TERMS == 1) {
                    *reinterpret_cast<uint4*>(hi + o) = q;
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
}

// General path: any strides, element by element, into the K-major tile [128 rows x 64 k].
template <int TERMS, typename T>
__device__ __forceinline__ void stage_scalar(uint8_t* hi, uint8_t* mid, uint8_t* lo, const T* __restrict__ x, long long sr, long long sk,
                                             int r0, int nrows, int k0, int K, int tid, int ones_row = -1) {
    const bool r_fast = (sr == 1);          // rows contiguous in memory: let consecutive threads walk rows
    for (int i = tid; i < 128 * kChunk; i += 128) {
        const int r = r_fast ? (i & 127) : (i >> 6);
        const int k = r_fast ? (i >> 7) : (i & 63);
        float v = 0.f;
        if (r0 + r == ones_row && k0 + k < K) v = 1.f;
        else if (r0 + r < nrows && k0 + k < K) v = load_elem(x + (long long)(r0 + r) * sr + (long long)(k0 + k) * sk);
        split_store<TERMS>(hi, mid, lo, tile_off(r, k), v);
    }
}

// mode: 0 = general (K-major tile, scalar), 1 = K contiguous (K-major tile, vector), 2 = M / N contiguous (MN-major tile, vector)
// `ones`: row index `nrows` of the operand (one past the memory rows) reads as 1.0.
// MODE / BF known at compile time (the specialised instantiations: only that path is compiled, staging loops unrolled) or
// MODE = -1: run-time mode and element type (the general kernel; loops rolled to keep its code small -- a one-tile CTA
// runs every instruction exactly once, so the time of a small GEMM is instruction fetch as much as anything else).
template <int TERMS, int MODE, bool BF>
__device__ __forceinline__ void stage_operand(int mode_rt, bool bf_rt, uint8_t* hi, uint8_t* mid, uint8_t* lo, const void* x, long long s_row,
                                              long long s_k, int r0, int nrows, int k0, int K, int tid, bool ones) {
    const int o = ones ? nrows : -1;
    if constexpr (MODE >= 0) {
        using T = typename std::conditional<BF, __nv_bfloat16, float>::type;
        const T* xt = reinterpret_cast<const T*>(x);
        if constexpr (MODE == 1) stage_vec<TERMS, 128, 64, T, 8>(hi, mid, lo, xt, s_row, r0, nrows, k0, K, tid, o, -1);
        else if constexpr (MODE == 2) stage_vec<TERMS, 64, 128, T, 8>(hi, mid, lo, xt, s_k, k0, K, r0, nrows, tid, -1, o);
        else stage_scalar<TERMS, T>(hi, mid, lo, xt, s_row, s_k, r0, nrows, k0, K, tid, o);
    } else if (bf_rt) {
        const __nv_bfloat16* xt = reinterpret_cast<const __nv_bfloat16*>(x);
        if (mode_rt == 1) stage_vec<TERMS, 128, 64, __nv_bfloat16, 1>(hi, mid, lo, xt, s_row, r0, nrows, k0, K, tid, o, -1);
        else if (mode_rt == 2) stage_vec<TERMS, 64, 128, __nv_bfloat16, 1>(hi, mid, lo, xt, s_k, k0, K, r0, nrows, tid, -1, o);
        else stage_scalar<TERMS, __nv_bfloat16>(hi, mid, lo, xt, s_row, s_k, r0, nrows, k0, K, tid, o);
    } else {
        const float* xt = reinterpret_cast<const float*>(x);
        if (mode_rt == 1) stage_vec<TERMS, 128, 64, float, 1>(hi, mid, lo, xt, s_row, r0, nrows, k0, K, tid, o, -1);
        else if (mode_rt == 2) stage_vec<TERMS, 64, 128, float, 1>(hi, mid, lo, xt, s_k, k0, K, r0, nrows, tid, -1, o);
        else stage_scalar<TERMS, float>(hi, mid, lo, xt, s_row, s_k, r0, nrows, k0, K, tid, o);
    }
}

template <int TERMS, int AM, bool ABF, int BM, bool BBF>
__global__ void __launch_bounds__(128, 1) gemm_kernel(const gp_gemm_args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[128];
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
    sbias[tid] = (p.bias && n0 + tid < p.N) ? p.bias[n0 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;
    const int a_mode = AM >= 0 ? AM : (p.flags & 3), b_mode = BM >= 0 ? BM : ((p.flags >> 2) & 3);       // operand staging modes
    const uint32_t idesc = idesc_bf16(128, a_mode == 2, b_mode == 2);
    auto adesc = [&](uint32_t base, int ks) { return a_mode == 2 ? desc_mnmajor(base, 64, ks) : desc_kmajor(base, 128, ks); };
    auto bdesc = [&](uint32_t base, int ks) { return b_mode == 2 ? desc_mnmajor(base, 64, ks) : desc_kmajor(base, 128, ks); };
    for (int c = c_begin; c < c_end; ++c) {
        stage_operand<TERMS, AM, ABF>(a_mode, p.a_bf16 != 0, a_t[0], a_t[1], a_t[2], p.a, p.a_sm, p.a_sk, m0, p.M, c * kChunk, p.K, tid, false);
        const int nb_rows = p.b_ones ? p.N - 1 : p.N;       // rows of B that exist in memory
        stage_operand<TERMS, BM, BBF>(b_mode, p.b_bf16 != 0, b_t[0], b_t[1], b_t[2], p.b, p.b_sn, p.b_sk, n0, nb_rows, c * kChunk, p.K, tid, p.b_ones != 0);
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
                    mma_ss(tmem, adesc(al, ks), bdesc(bh, ks), idesc, first);
                    mma_ss(tmem, adesc(ah, ks), bdesc(bl, ks), idesc, 1u);
                    mma_ss(tmem, adesc(am, ks), bdesc(bm, ks), idesc, 1u);
                    mma_ss(tmem, adesc(am, ks), bdesc(bh, ks), idesc, 1u);
                    mma_ss(tmem, adesc(ah, ks), bdesc(bm, ks), idesc, 1u);
                    mma_ss(tmem, adesc(ah, ks), bdesc(bh, ks), idesc, 1u);
                } else {
                    mma_ss(tmem, adesc(ah, ks), bdesc(bh, ks), idesc, first);
                }
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);        // the staged tiles have been read: the next chunk may overwrite them
        phase ^= 1;
        tc_fence_after();
    }
    // epilogue: lane == row of the tile; 16 columns at a time.  Whole 16-column groups of contiguous, aligned output rows
    // take the vector path (64 contiguous bytes per thread and tensor); ragged edges and strided outputs the scalar one.
    const int m = m0 + tid;
    const uint32_t tl = tmem_addr(tmem, (tid >> 5) * 32, 0);
    const bool direct = (gridDim.z == 1);
    float* const cf = reinterpret_cast<float*>(p.c);
    __nv_bfloat16* const cb = reinterpret_cast<__nv_bfloat16*>(p.c);
    const bool has_work = c_end > c_begin;
#pragma unroll 1
    for (int cc = 0; cc < 128 && n0 + cc < p.N; cc += 16) {
        uint32_t v[16];
        if (has_work) {
            tmem_ld16(tl + cc, v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (m >= p.M) continue;
        if (!direct) {
            // partial rows are padded to a multiple of four floats (gp_gemm_partial_ld): always 16-byte stores; the pad
            // columns receive the zero products of the zero-filled operand rows and are never read
            const int ldp = (p.N + 3) & ~3;
            float* pp = p.partials + ((size_t)blockIdx.z * p.M + m) * ldp + n0 + cc;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + cc + 4 * j < p.N)
                    reinterpret_cast<float4*>(pp)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            continue;
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {          // bias of this column tile from shared memory (zeros without one / past N)
            const float4 q = *reinterpret_cast<const float4*>(sbias + cc + 4 * j);
            f[4 * j] = __uint_as_float(v[4 * j]) + q.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + q.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + q.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + q.w;
        }
        const long long row_off = (long long)m * p.c_sm + n0 + cc;
        if ((n0 + cc + 16 <= p.N) && (p.flags & 16)) {
            if (p.resid) {
                const float4* r4 = reinterpret_cast<const float4*>(p.resid + row_off);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float4 q = __ldg(r4 + j); f[4 * j] += q.x; f[4 * j + 1] += q.y; f[4 * j + 2] += q.z; f[4 * j + 3] += q.w; }
            }
            if (p.accumulate) {
                if (p.c_bf16) {
                    acc8(*reinterpret_cast<const uint4*>(cb + row_off), f);
                    acc8(*reinterpret_cast<const uint4*>(cb + row_off + 8), f + 8);
                } else {
                    const float4* c4 = reinterpret_cast<const float4*>(cf + row_off);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float4 q = c4[j]; f[4 * j] += q.x; f[4 * j + 1] += q.y; f[4 * j + 2] += q.z; f[4 * j + 3] += q.w; }
                }
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.c_bf16) {
                uint4* d = reinterpret_cast<uint4*>(cb + row_off);
                d[0] = pack8(f);
                d[1] = pack8(f + 8);
            } else {
                float4* d = reinterpret_cast<float4*>(cf + row_off);
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + cc + j;
                if (n >= p.N) break;
                const long long o = (long long)m * p.c_sm + (long long)n * p.c_sn;
                float x = f[j];
                if (p.resid) x += p.resid[o];
                if (p.accumulate) x += p.c_bf16 ? __bfloat162float(cb[o]) : cf[o];
                if (p.relu) x = fmaxf(x, 0.f);
                if (p.c_bf16) cb[o] = __float2bfloat16_rn(x);
                else cf[o] = x;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 128);
}

// C(m, n) = [C(m, n)] + [resid] + bias[n] + sum_z partials[z][m][n].  Block = 32 outputs x 8 z-groups: group g adds
// z = g, g + 8, ... in ascending order, then the eight group sums are added in fixed order (bit-reproducible).
__global__ void __launch_bounds__(256) gemm_reduce_kernel(const gp_gemm_args p, int nz) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long total = (long long)p.M * p.N;
    const long long i = (long long)blockIdx.x * 32 + tx;
    const int m = (int)(i / p.N), n = (int)(i - (long long)m * p.N);
    const int ldp = (p.N + 3) & ~3;
    float f = 0.f;
    if (i < total)
        for (int z = ty; z < nz; z += 8) f += p.partials[((size_t)z * p.M + m) * ldp + n];
    sh[ty][tx] = f;
    __syncthreads();
    if (ty != 0 || i >= total) return;
    f = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) f += sh[g][tx];
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
    GP_REQUIRE(a.M >= 0 && a.N >= 0 && a.K >= 0, "gp_gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
    if (a.M == 0 || a.N == 0) return 0;                      // an empty output (e.g. the edge MLP of an edge-less graph)
    // K == 0 (a weight gradient over zero rows) is legal: the product is zero, the epilogue still runs; A / B may be NULL then
    GP_REQUIRE(a.c && (a.K == 0 || (a.a && a.b)), "gp_gemm: null operand");
    GP_REQUIRE(a.terms == 1 || a.terms == 3, "gp_gemm: terms must be 1 (bf16 operands) or 3 (three-term split)");
    GP_REQUIRE(a.split_k >= 1 && a.split_k <= 1024, "gp_gemm: split_k must be in [1, 1024]");
    GP_REQUIRE(a.split_k == 1 || a.partials != nullptr, "gp_gemm: split_k > 1 needs a partials buffer of split_k * M * ((N + 3) & ~3) floats");
    GP_REQUIRE(!(a.resid && a.c_bf16), "gp_gemm: resid needs an fp32 output");
    // vector paths: K contiguous + 16-byte aligned rows (operands), N contiguous + aligned rows (output)
    auto aligned = [](const void* base, long long stride_elems, int elem) {
        return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (stride_elems * elem) % 16 == 0;
    };
    // flags: bits 0-1 staging mode of A, bits 2-3 of B (0 general, 1 K contiguous, 2 M / N contiguous), bit 4 vector output
    a.flags = 0;
    if (a.a_sk == 1 && a.K % 8 == 0 && aligned(a.a, a.a_sm, a.a_bf16 ? 2 : 4)) a.flags |= 1;
    else if (a.a_sm == 1 && a.M % 8 == 0 && aligned(a.a, a.a_sk, a.a_bf16 ? 2 : 4)) a.flags |= 2;
    if (a.b_sk == 1 && a.K % 8 == 0 && aligned(a.b, a.b_sn, a.b_bf16 ? 2 : 4)) a.flags |= 1 << 2;
    else if (a.b_sn == 1 && (a.N - (a.b_ones ? 1 : 0)) % 8 == 0 && aligned(a.b, a.b_sk, a.b_bf16 ? 2 : 4)) a.flags |= 2 << 2;
    if (a.c_sn == 1 && aligned(a.c, a.c_sm, a.c_bf16 ? 2 : 4)) a.flags |= 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * a.terms * kTileBytes + 1024;
    const dim3 grid((a.M + 127) / 128, (a.N + 127) / 128, a.split_k);
    // specialised instantiations for the operand patterns of a Linear layer (forward: A, B K-contiguous; dgrad: B read
    // transposed; wgrad: both read transposed); anything else runs the general kernel
    const int am = a.flags & 3, bm = (a.flags >> 2) & 3;
    const bool abf = a.a_bf16 != 0, bbf = a.b_bf16 != 0;
#define GP_GEMM_CASE(T_, AM_, ABF_, BM_, BBF_)                                                                                  \
    if (a.terms == T_ && ((AM_) < 0 || (am == (AM_) && abf == (ABF_) && bm == (BM_) && bbf == (BBF_)))) {                      \
        static bool attr = false;                                                                                              \
        auto kern = gemm_kernel<T_, AM_, ABF_, BM_, BBF_>;                                                                      \
        if (!attr) {                                                                                                           \
            GP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
            attr = true;                                                                                                       \
        }                                                                                                                      \
        kern<<<grid, 128, smem, st>>>(a);                                                                                      \
        launched = true;                                                                                                       \
    } else
    bool launched = false;
    GP_GEMM_CASE(1, 1, false, 1, false)      // forward, fp32 activations
    GP_GEMM_CASE(1, 1, true, 1, false)       // forward, bf16 activations
    GP_GEMM_CASE(1, 1, false, 2, false)      // dgrad
    GP_GEMM_CASE(1, 2, false, 2, false)      // wgrad, fp32 activations
    GP_GEMM_CASE(1, 2, false, 2, true)       // wgrad, bf16 activations
    GP_GEMM_CASE(3, 1, false, 1, false)
    GP_GEMM_CASE(3, 1, false, 2, false)
    GP_GEMM_CASE(3, 2, false, 2, false)
    GP_GEMM_CASE(1, -1, false, -1, false)
    GP_GEMM_CASE(3, -1, false, -1, false)
    {}
#undef GP_GEMM_CASE
    GP_REQUIRE(launched, "gp_gemm: no kernel for terms=%d", a.terms);
    GP_CHECK_CUDA(cudaGetLastError());
    if (a.split_k > 1) {
        const long long total = (long long)a.M * a.N;
        gemm_reduce_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(a, a.split_k);
        GP_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}
