// linear_bwd.cu -- backward of the bias-free per-node projection  P = x . Wp^T  (gp_linear_bwd).
//
// P holds the node-side halves of the first edge-MLP layer and of the first node-MLP layer
// (W1 = [W1e | W1d | W1s] of layers.py:1058 splits by linearity; see gp_b200.h).  Its gradient
// arrives in up to three `hidden`-wide pieces (receiver sums, sender sums, node-MLP delta).
// Per 128-row tile, for each piece s:   dX += dP_s . Wp_s      (dgrad, accumulated in TMEM)
//                                       dWp_s += dP_s^T . x    (wgrad, TMEM-resident per CTA)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tile_util.cuh"

namespace {
using namespace gp;

struct LinMaps {
    CUtensorMap x, src[3];
    uint32_t use;                                   // bit 0: x, bits 1..3: bf16 sources
};

template <int H>
__global__ void __launch_bounds__(256, 1) linear_bwd_kernel(const gp_linear_bwd_args p, const __grid_constant__ LinMaps maps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t mma_bar, tma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, row = tid & 127, half = tid >> 7;
    const int warp = warp_uniform(tid >> 5);
    const int S = p.n_src;
    const int wrows = S * H;
    constexpr int NSTG = H >= 128 ? 2 : 1;          // tile buffers covered by the fp32 staging tile
    const int nbuf = S > NSTG ? S : NSTG;
    uint32_t off = 0;
    uint8_t* w_t = smem + off;  off += ((H + 63) >> 6) * wrows * 128;
    uint8_t* abuf = smem + off; off += nbuf * kBufBytes;          // one gradient tile per source
    uint8_t* xbuf = smem + off;
    float* stg = reinterpret_cast<float*>(abuf);                 // reused after the MMAs have read abuf

    // the first tile of this CTA: x is forward data (cold), the incoming dX rows are older than the previous kernels --
    // both go to L2 while the weights are staged (a CTA runs ~4 tiles)
    if ((int)blockIdx.x < ((p.rows + 127) >> 7)) {
        const int R0_ = (int)blockIdx.x << 7;
        if ((maps.use & 1u) && warp == 0 && elect_one())
            for (int b = 0; b < (H + 63) >> 6; ++b) tma_prefetch_2d(&maps.x, b * 64, R0_);
        if (p.dx_in) {
            constexpr int LPR = (H * 4 / 128) > 0 ? H * 4 / 128 : 1;
            for (int i = tid; i < 128 * LPR; i += 256)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dx_in + (size_t)min(R0_ + i / LPR, p.rows - 1) * H + (i % LPR) * 32));
        }
    }
    stage_weight(w_t, p.w, wrows, H);
    cp_async_commit();
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        mbar_init(&tma_bar, 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&tmem_slot, 512);
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // everything above is independent of the previous kernel in the stream (parameters only); from here on
    // its outputs are read, and the next kernel may start its own prologue
    pdl_wait();
    pdl_launch_dependents();

    const uint32_t tmem = tmem_slot;
    const uint32_t tlane = tmem_addr(tmem, (row >> 5) * 32, 0);
    const uint32_t a_s = smem_u32(abuf), x_s = smem_u32(xbuf), w_s = smem_u32(w_t);
    const uint32_t lbo_h = (H >= 128) ? 16384u : 0u;
    uint32_t phase = 0, tphase = 0;
    const int n_tiles = (p.rows + 127) >> 7;
    constexpr int CH = H / 2;
    constexpr int KC = H / 8;                        // 16-byte chunks per bf16 row
    constexpr int CPT = 128 * KC / 256;              // bf16 chunks per thread per tile
    constexpr int CPR = H / 4;                       // 16-byte chunks per fp32 row
    constexpr int FPT = 128 * CPR / 256;             // fp32 chunks per thread per tile
    bool first = true;
    uint32_t tma_blocks = (maps.use & 1u) ? (H + 63) >> 6 : 0;
    for (int s = 0; s < S; ++s) tma_blocks += (maps.use & (2u << s)) ? (H + 63) >> 6 : 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        const int R0 = tile << 7;
        // ---- loads: x and bf16 pieces by TMA (rows past the end = 0), fp32 pieces converted on the fly
        if (tma_blocks && warp == 0 && elect_one()) {
            mbar_arrive_expect_tx(&tma_bar, tma_blocks * 16384u);
            if (maps.use & 1u)
                for (int b = 0; b < (H + 63) >> 6; ++b) tma_load_2d(x_s + b * 16384, &maps.x, b * 64, R0, &tma_bar);
#pragma unroll
            for (int s = 0; s < 3; ++s)          // constant indices: the descriptors must stay in parameter space
                if (s < S && (maps.use & (2u << s)))
                    for (int b = 0; b < (H + 63) >> 6; ++b)
                        tma_load_2d(a_s + s * kBufBytes + b * 16384, &maps.src[s], b * 64, R0, &tma_bar);
        }
        if (!(maps.use & 1u)) stage_rows(xbuf, p.x, nullptr, H, p.ldx, R0, p.rows, tid, 256);
        cp_async_commit();
        for (int s = 0; s < S; ++s) {
            uint8_t* ab = abuf + s * kBufBytes;
            if (p.src_f32[s]) {
                float4 u[2 * CPT];                   // the whole piece is requested before the first conversion
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int i = tid + j * 256;
                    const int r = i / KC, ch = i % KC;
                    const float4* sp = reinterpret_cast<const float4*>(p.src_f32[s] + (size_t)min(R0 + r, p.rows - 1) * p.ld_src[s] + ch * 8);
                    u[2 * j] = __ldg(sp);
                    u[2 * j + 1] = __ldg(sp + 1);
                }
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int i = tid + j * 256;
                    const int r = i / KC, ch = i % KC;
                    const bool ok = R0 + r < p.rows;      // rows past the end add nothing to the weight gradients
                    *reinterpret_cast<uint4*>(ab + sw128_off(128, r, ch * 8)) =
                        ok ? make_uint4(pack_bf16(u[2 * j].x, u[2 * j].y), pack_bf16(u[2 * j].z, u[2 * j].w),
                                        pack_bf16(u[2 * j + 1].x, u[2 * j + 1].y), pack_bf16(u[2 * j + 1].z, u[2 * j + 1].w))
                           : make_uint4(0, 0, 0, 0);
                }
            } else if (!(maps.use & (2u << s))) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int i = tid + j * 256;
                    const int r = i / KC, ch = i % KC;
                    *reinterpret_cast<uint4*>(ab + sw128_off(128, r, ch * 8)) =
                        (R0 + r < p.rows) ? ldg16(p.src_bf16[s] + (size_t)(R0 + r) * p.ld_src[s] + ch * 8) : make_uint4(0, 0, 0, 0);
                }
            }
        }
        cp_async_wait<0>();
        if (tma_blocks) {
            mbar_wait(&tma_bar, tphase);
            tphase ^= 1;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- all MMAs of the tile in one batch:  dWp_s += dP_s^T . x   and   dX = sum_s dP_s . Wp_s
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t id_w = idesc_bf16(H, true, true), id_d = idesc_bf16(H, false, true);
            for (int s = 0; s < S; ++s) {
                const uint32_t as = a_s + s * kBufBytes;
                for (int ks = 0; ks < 8; ++ks)
                    mma_ss(tmem + 128 * (1 + s), desc_mnmajor(as, 128, ks, 0, lbo_h), desc_mnmajor(x_s, 128, ks), id_w,
                           (ks > 0) ? 1u : (first ? 0u : 1u));
                for (int ks = 0; ks < (H >> 4); ++ks)
                    mma_ss(tmem, desc_kmajor(as, 128, ks), desc_mnmajor(w_s + s * H * 128, wrows, ks), id_d,
                           (s > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(&mma_bar);
        }
        // the next tile of this CTA: everything its loads will ask for goes to L2 now (bulk tiles by the copy engine, the
        // fp32 rows one prefetch per 128-byte line), so that they wait on L2 instead of HBM
        if (tile + (int)gridDim.x < n_tiles) {
            const int Rn = (tile + (int)gridDim.x) << 7;
            if (warp == 0 && elect_one()) {
                if (maps.use & 1u)
                    for (int b = 0; b < (H + 63) >> 6; ++b) tma_prefetch_2d(&maps.x, b * 64, Rn);
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    if (s < S && (maps.use & (2u << s)))
                        for (int b = 0; b < (H + 63) >> 6; ++b) tma_prefetch_2d(&maps.src[s], b * 64, Rn);
            }
            if (p.dx_in) {
                constexpr int LPR = (H * 4 / 128) > 0 ? H * 4 / 128 : 1;       // 128-byte lines per fp32 row
                for (int i = tid; i < 128 * LPR; i += 256)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dx_in + (size_t)min(Rn + i / LPR, p.rows - 1) * H + (i % LPR) * 32));
            }
        }
        // while the tensor core runs: the incoming dX rows of this tile (row-major chunks)
        float4 din[FPT];
        if (p.dx_in) {
#pragma unroll
            for (int j = 0; j < FPT; ++j) {
                const int i = tid + j * 256;
                const int r = i / CPR, ch = i % CPR;
                din[j] = __ldg(reinterpret_cast<const float4*>(p.dx_in + (size_t)min(R0 + r, p.rows - 1) * H) + ch);
            }
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- dX: TMEM (row per lane) -> fp32 staging tile -> coalesced rows
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tlane + half * CH + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(stg) + stage_f32_off(row, (half * CH + c0) / 4 + q, H)) =
                    make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        tc_fence_before();
        __syncthreads();
#pragma unroll
        for (int j = 0; j < FPT; ++j) {
            const int i = tid + j * 256;
            const int r = i / CPR, ch = i % CPR;
            if (R0 + r < p.rows) {
                float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stg) + stage_f32_off(r, ch, H));
                if (p.dx_in) { o.x += din[j].x; o.y += din[j].y; o.z += din[j].z; o.w += din[j].w; }
                reinterpret_cast<float4*>(p.dx_out + (size_t)(R0 + r) * H)[ch] = o;
            }
        }
        fence_async_smem();      // the staging tile is refilled by TMA / per-thread copies next
        __syncthreads();
    }

    // ---- dump the weight-gradient accumulators (lane r <-> row r of dWp_s) through the staging tile
    tc_fence_after();
    float* P = p.partials + (size_t)blockIdx.x * S * H * H;
    for (int s = 0; s < S; ++s)
        tmem_rows_to_global<256>(tlane, 128 * (1 + s), H, H, P + (size_t)s * H * H, H, reinterpret_cast<uint8_t*>(stg), tid);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

template <int H>
int launch(const gp_linear_bwd_args& a, int32_t* grid_out, cudaStream_t st) {
    const int nbuf = a.n_src > (H >= 128 ? 2 : 1) ? a.n_src : (H >= 128 ? 2 : 1);
    const size_t smem = 1024 + (size_t)((H + 63) / 64) * a.n_src * H * 128 + (size_t)(nbuf + 1) * kBufBytes;
    GP_REQUIRE((int)smem <= gp::max_smem_optin(), "gp_linear_bwd: needs %zu B of shared memory (> %d)", smem, gp::max_smem_optin());
    static int smem_set = 0;      // raised once per instantiation (and never inside a stream capture twice)
    if ((int)smem > smem_set) {
        GP_CHECK_CUDA(cudaFuncSetAttribute(linear_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = (int)smem;
    }
    LinMaps maps;
    memset(&maps, 0, sizeof(maps));
    static const bool no_tma = getenv("GP_NO_TMA") != nullptr;
    if (!no_tma) {
        if (gp::tma_map_2d(&maps.x, a.x, a.rows, H, a.ldx)) maps.use |= 1u;
        for (int s = 0; s < a.n_src; ++s)
            if (a.src_bf16[s] && gp::tma_map_2d(&maps.src[s], a.src_bf16[s], a.rows, H, a.ld_src[s])) maps.use |= 2u << s;
    }
    const int n_tiles = (a.rows + 127) / 128;
    int grid = n_tiles < gp::sm_count() ? n_tiles : gp::sm_count();
    if (grid > 0) {           // minimal number of tile rounds with the fewest CTAs (each dumps a block of partials)
        const int rounds = (n_tiles + grid - 1) / grid;
        grid = (n_tiles + rounds - 1) / rounds;
    }
    GP_CHECK_CUDA(gp::launch_kernel(linear_bwd_kernel<H>, dim3(grid), dim3(256), smem, st, a, maps));
    if (grid_out) *grid_out = grid;
    return 0;
}

// out[n][:] = sum over j in [rowptr[n], rowptr[n+1]) of src[perm[j]][:]   (bf16 rows, fp32 sum, fixed order)
template <int VPT, typename OutT>
__global__ void segsum_gather_kernel(const gp_bf16* __restrict__ src, int ld, const int32_t* __restrict__ perm,
                                     const int32_t* __restrict__ rowptr, int num_segments, OutT* __restrict__ out) {
    const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (seg >= num_segments) return;
    const int b = rowptr[seg], e = rowptr[seg + 1];
    pdl_wait();                 // (the layout is older than the previous kernel; its rows are read from here on)
    pdl_launch_dependents();
    float acc[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc[i] = 0.f;
    auto add_row = [&](int r) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(src) + (size_t)r * ld + lane * VPT;
        if constexpr (VPT == 4) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(rp));
            acc[0] = add_bf16_lo(q.x, acc[0]); acc[1] = add_bf16_hi(q.x, acc[1]);
            acc[2] = add_bf16_lo(q.y, acc[2]); acc[3] = add_bf16_hi(q.y, acc[3]);
        } else if constexpr (VPT == 2) {
            const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(rp));
            acc[0] = add_bf16_lo(q, acc[0]); acc[1] = add_bf16_hi(q, acc[1]);
        } else {
            acc[0] += __bfloat162float(rp[0]);
        }
    };
    // the segment's row ids are fetched by the lanes in one coalesced read and broadcast, so the row
    // loads of a segment do not wait on one another (same summation order as a serial walk)
    for (int base = b; base < e; base += 32) {
        const int mine = (base + lane < e) ? (perm ? __ldg(perm + base + lane) : base + lane) : 0;
        const int n = min(32, e - base);
        int j = 0;
        for (; j + 4 <= n; j += 4) {
            const int r0 = __shfl_sync(0xffffffffu, mine, j), r1 = __shfl_sync(0xffffffffu, mine, j + 1);
            const int r2 = __shfl_sync(0xffffffffu, mine, j + 2), r3 = __shfl_sync(0xffffffffu, mine, j + 3);
            if constexpr (VPT == 4) {
                const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(src) + lane * VPT;
                const uint2 q0 = __ldg(reinterpret_cast<const uint2*>(bp + (size_t)r0 * ld));
                const uint2 q1 = __ldg(reinterpret_cast<const uint2*>(bp + (size_t)r1 * ld));
                const uint2 q2 = __ldg(reinterpret_cast<const uint2*>(bp + (size_t)r2 * ld));
                const uint2 q3 = __ldg(reinterpret_cast<const uint2*>(bp + (size_t)r3 * ld));
                const uint2 qs[4] = {q0, q1, q2, q3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[0] = add_bf16_lo(qs[k].x, acc[0]); acc[1] = add_bf16_hi(qs[k].x, acc[1]);
                    acc[2] = add_bf16_lo(qs[k].y, acc[2]); acc[3] = add_bf16_hi(qs[k].y, acc[3]);
                }
            } else {
                add_row(r0); add_row(r1); add_row(r2); add_row(r3);
            }
        }
        for (; j < n; ++j) add_row(__shfl_sync(0xffffffffu, mine, j));
    }
    OutT* o = out + (size_t)seg * (32 * VPT) + lane * VPT;
    if constexpr (sizeof(OutT) == 4) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) o[i] = acc[i];
    } else if constexpr (VPT == 4) {
        *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]));
    } else if constexpr (VPT == 2) {
        *reinterpret_cast<uint32_t*>(o) = pack_bf16(acc[0], acc[1]);
    } else {
        *reinterpret_cast<__nv_bfloat16*>(o) = __float2bfloat16(acc[0]);
    }
}
}  // namespace

extern "C" int gp_linear_bwd(const gp_linear_bwd_args* args, int hidden, int32_t* grid_out, void* stream) {
    GP_REQUIRE(args != nullptr, "gp_linear_bwd: null args");
    const gp_linear_bwd_args& a = *args;
    GP_REQUIRE(a.rows > 0 && a.n_src >= 1 && a.n_src <= 3, "gp_linear_bwd: bad rows / n_src");
    for (int s = 0; s < a.n_src; ++s)
        GP_REQUIRE((a.src_f32[s] != nullptr) != (a.src_bf16[s] != nullptr) && a.ld_src[s] % 8 == 0,
                   "gp_linear_bwd: source %d needs exactly one pointer and an aligned stride", s);
    GP_REQUIRE(a.w && a.x && a.dx_out && a.partials, "gp_linear_bwd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (hidden) {
        case 128: return launch<128>(a, grid_out, st);
        case 64: return launch<64>(a, grid_out, st);
        case 32: return launch<32>(a, grid_out, st);
        default: gp::set_error("gp_linear_bwd: unsupported hidden size %d", hidden); return -1;
    }
}

template <typename OutT>
static int segsum_gather_launch(const gp_bf16* src, int32_t ld, const int32_t* perm, const int32_t* rowptr, int32_t num_segments,
                                int32_t hidden, OutT* out, void* stream) {
    if (num_segments <= 0) return 0;
    const int threads = 256;
    const int blocks = (int)(((size_t)num_segments * 32 + threads - 1) / threads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (hidden) {
        case 128: GP_CHECK_CUDA(gp::launch_kernel(segsum_gather_kernel<4, OutT>, dim3(blocks), dim3(threads), 0, st, src, (int)ld, perm, rowptr, (int)num_segments, out)); break;
        case 64: GP_CHECK_CUDA(gp::launch_kernel(segsum_gather_kernel<2, OutT>, dim3(blocks), dim3(threads), 0, st, src, (int)ld, perm, rowptr, (int)num_segments, out)); break;
        case 32: GP_CHECK_CUDA(gp::launch_kernel(segsum_gather_kernel<1, OutT>, dim3(blocks), dim3(threads), 0, st, src, (int)ld, perm, rowptr, (int)num_segments, out)); break;
        default: gp::set_error("gp_segsum_gather: unsupported hidden size %d", hidden); return -1;
    }
    return 0;
}

extern "C" int gp_segsum_gather(const gp_bf16* src, int32_t ld, const int32_t* perm, const int32_t* rowptr,
                                int32_t num_segments, int32_t hidden, float* out, void* stream) {
    return segsum_gather_launch<float>(src, ld, perm, rowptr, num_segments, hidden, out, stream);
}

extern "C" int gp_segsum_gather_bf16(const gp_bf16* src, int32_t ld, const int32_t* perm, const int32_t* rowptr,
                                     int32_t num_segments, int32_t hidden, gp_bf16* out, void* stream) {
    return segsum_gather_launch<gp_bf16>(src, ld, perm, rowptr, num_segments, hidden, out, stream);
}
