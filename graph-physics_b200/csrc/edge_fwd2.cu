// edge_fwd2.cu -- the processor's edge update (gp_mlp_fwd, edge form, H = 128) as a warp-specialised CTA-pair
// kernel: gather -> 4-layer MLP on tcgen05 (cta_group::2, M = 256) -> RMSNorm -> residual -> receiver-sorted
// segment sum (graphphysics/models/layers.py:989-1072 of the reference, in the re-associated form of DESIGN.md §2).
//
// Why pairs: the four 128x128 bf16 weight matrices fill 128 KB of one SM's shared memory and leave room for two
// tile buffers -- the round-1 kernel therefore ran load, compute and write-back of a tile back to back in one
// buffer.  Here a cluster of two CTAs shares the weights: each CTA keeps rows [rank*64, +64) of every matrix
// (64 KB) and one tcgen05.mma.cta_group::2 instruction multiplies BOTH CTAs' 128-row tiles by the full matrix.
// The freed 64 KB become three more tile buffers per CTA, and the tile loop is split into roles that overlap:
//
//   warps  0-15  epilogue, two tile slots of 256 threads (thread = (row of the tile = TMEM lane, column half): the four
//                layers of a tile are one dependent chain per slot, so its length is what bounds the kernel -- two
//                threads per row halve every step of it): accumulator pre-load from the gathered rows, ReLU / bf16
//                conversion between layers, RMSNorm
//   warps 16-23  drain, 128 threads per slot: segment walk over the normalised update u, e' = e + u, stores
//   warps 24-25  MMA issue (leader CTA): one thread per slot waits for BOTH CTAs' operands and issues the pair MMAs
//   warps 26-27  gather producers: the sender rows P[src] of the NEXT tile into a staging buffer (TMA gather4 / cp.async)
//
// Buffers per CTA: ACT[slot] (layer operand, rewritten in place by the epilogues; the next tile's e arrives here by
// TMA while the RMSNorm epilogue of the current one runs), U[slot] (normalised update, handed to the drain warps so
// the slot starts its next tile at once), G (staged sender rows, filled one tile ahead, shared by the slots in tile
// order).  All hand-offs are mbarriers; the only block-wide barrier is in the prologue.  An odd tile at the end of
// the edge list is processed as an all-padding tile (loads clamp / zero-fill, stores are clipped), so both CTAs of a
// pair always run the same protocol.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tile_util.cuh"

struct gp_mlp_fwd_args;
namespace gp {
int try_edge_fwd2(const gp_mlp_fwd_args& a, cudaStream_t st);
}

// registers per thread of the three roles after setmaxnreg (16 epilogue, 8 drain, 4 control warps; 28 x 72 in total)
#ifndef GP_FWD2_REGS_EPI
#define GP_FWD2_REGS_EPI 80
#define GP_FWD2_REGS_DRAIN 72
#define GP_FWD2_REGS_CTRL 40
#endif
static_assert(16 * GP_FWD2_REGS_EPI + 8 * GP_FWD2_REGS_DRAIN + 4 * GP_FWD2_REGS_CTRL <= 28 * 72, "register budget of the CTA");

namespace {
using namespace gp;

constexpr int H = 128;
constexpr int kEpi = 256;                    // epilogue threads per slot: (row of the tile = TMEM lane) x (column half)
constexpr int kDrain = 128;                  // drain threads per slot
constexpr int kNT = 2 * kEpi + 2 * kDrain + 128;      // 896: 16 epilogue + 8 drain + 4 control warps
constexpr int kWDrain = 2 * kEpi / 32, kWMma = kWDrain + 2 * kDrain / 32, kWProd = kWMma + 2;     // first warp of each role
constexpr int kTile = 128 * H * 2;           // 32 KB
constexpr int kWHalf = 64 * H * 2;           // 16 KB: rows [rank*64, +64) of one weight matrix

// shared-memory map (dynamic, base 1024-aligned)
constexpr uint32_t kOffW = 0;
constexpr uint32_t kOffAct = 4 * kWHalf;                 // 65536
constexpr uint32_t kOffU = kOffAct + 2 * kTile;
constexpr uint32_t kOffG = kOffU + 2 * kTile;
// With ~226 KB of shared memory per CTA the L1 is a few KB, so per-tile parameter reads must not go to global memory
// (and nothing may spill): the four biases and the RMSNorm scale live in shared memory.
// TMEM per slot: a 128-column fp32 accumulator and 64 columns holding the bf16 activation of the hidden layers as the
// A operand of the next MMA (tcgen05.mma with A in tensor memory), so h1 / h3 never touch shared memory: the L1/shared
// data pipe is the busiest unit of this kernel (operand fetch + epilogue stores + staging, ~75 % in the ncu capture).
constexpr uint32_t kOffBias = kOffG + kTile;             // float bias[4][128], scale[128]
constexpr uint32_t kOffBar = kOffBias + 5 * H * 4;
constexpr uint32_t kSmemBytes = kOffBar + 14 * 8 + 8;
constexpr uint32_t kSlotCols = 192, kColAct = 128;      // TMEM columns of a slot: [0, 128) accumulator, [128, 192) activation
static_assert(kSmemBytes <= 232448, "shared memory budget of one CTA");

struct Fwd2Maps {
    CUtensorMap e, h2, psrc;
    uint32_t save_h2;
    uint32_t gather4;      // sender rows P[src] by TMA tile::gather4 (psrc valid) instead of cp.async
};

struct Bars {
    uint64_t a_ready[2], mma_done[2], tma_e[2], g_full[2], g_empty[2], u_full[2], u_empty[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Arrive on a barrier of any CTA of the cluster (address from mapa).  Default (CTA-scope release) semantics, as in
// the 2-SM pipelines of CUTLASS: the operands this signals were already handed to the async proxy / tensor memory
// by fence.proxy.async and tcgen05.fence::before_thread_sync, and they are read by the tensor cores, not by the
// waiting thread.  A cluster-scope release here measured 1.2-1.7 k cycles per arrive (as much as a whole epilogue).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (t0 == 0) t0 = clock64();
        if (clock64() - t0 > (1ll << 32)) {
            printf("edge_fwd2: pair barrier timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (each CTA reads its own 128 lanes at `tmem_a`)
__device__ __forceinline__ void mma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all pair MMAs issued so far arrive on `bar` (same offset) in BOTH CTAs when complete
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// four indexed rows (64 columns each) of the gather map -> 4 x 128 bytes of an SW128 tile, completion on `bar`
__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const void* tmap, int32_t col, int32_t r0, int32_t r1, int32_t r2, int32_t r3,
                                            uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Segment sum of one (32-row sub-tile, column pair) unit of the bf16 tile `buf` -- the partition, the summation order
// and the output convention of tile_segment_sum<128, 256> (tile_util.cuh; complete segments -> seg_out_bf16 rounded
// once, pieces cut by the sub-tile -> seg_bnd for gp_seg_fixup_bf16), but straight-line: the row loop is unrolled
// with predicated stores instead of a branch per segment change.  The drain warps share the SM's instruction cache
// with four other roles, and every taken branch of the branchy walk cost an instruction-fetch stall
// (stall_no_inst at each reconvergence point in the ncu source view).  The segment ids come straight from global
// memory (coalesced, 34 per unit; rows past the end read as -1).
struct WalkIds {
    int4 v[8];
    int prev, next;
};
__device__ __forceinline__ WalkIds walk_load_ids(const int32_t* __restrict__ seg_id, int rows, int R0, int unit) {
    const int rb = R0 + (unit / (H / 2)) * 32;
    WalkIds w;
    if (rb + 32 <= rows) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w.v[i] = __ldg(reinterpret_cast<const int4*>(seg_id + rb) + i);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rb + 4 * i;
            w.v[i] = make_int4(r < rows ? __ldg(seg_id + r) : -1, r + 1 < rows ? __ldg(seg_id + r + 1) : -1,
                               r + 2 < rows ? __ldg(seg_id + r + 2) : -1, r + 3 < rows ? __ldg(seg_id + r + 3) : -1);
        }
    }
    // context rows of the TILE only (a sub-tile inside the tile sees its neighbours through the same array)
    w.prev = (rb > 0 && rb - 1 < rows) ? __ldg(seg_id + rb - 1) : -1;
    w.next = (rb + 32 < rows) ? __ldg(seg_id + rb + 32) : -1;
    return w;
}
__device__ __forceinline__ void segment_walk_unit(const uint8_t* buf, const WalkIds& ids, int R0, int unit, float* seg_bnd,
                                                  gp_bf16* seg_out_bf16) {
    constexpr int NP = H / 2, SUB = 32;
    const int part = unit / NP, cp = unit - part * NP;
    const int rb = part * SUB, c = cp * 2;
    const size_t sub_index = (size_t)(R0 + rb) / SUB;
    const uint32_t chunk = (c & 63) >> 3;
    const uint8_t* colbase = buf + (c >> 6) * 16384 + (c & 7) * 2;
    const int seg_prev = ids.prev, seg_next = ids.next;
    float* const bnd0 = seg_bnd + (sub_index * 2) * H + c;           // piece continuing from the previous sub-tile
    int cur = ids.v[0].x;
    bool first = true;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r8 = 0; r8 < SUB; r8 += 8) {
        const int4 sa = ids.v[r8 / 4], sb = ids.v[r8 / 4 + 1];
        const int sid[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const uint8_t* rowbase = colbase + (rb + r8) * 128;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const uint32_t*>(rowbase + j * 128 + ((chunk ^ j) << 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool brk = sid[j] != cur;
            const bool before = first && (seg_prev == cur);
            st_global_u32_if(seg_out_bf16 + (size_t)cur * H + c, pack_bf16(s0, s1), brk && !before && cur >= 0);
            st_global_f2_if(bnd0, s0, s1, brk && before && cur >= 0);
            first = first && !brk;
            s0 = brk ? 0.f : s0;
            s1 = brk ? 0.f : s1;
            cur = sid[j];
            s0 = add_bf16_lo(w[j], s0);
            s1 = add_bf16_hi(w[j], s1);
        }
    }
    const bool before = first && (seg_prev == cur), after = (seg_next == cur);
    st_global_u32_if(seg_out_bf16 + (size_t)cur * H + c, pack_bf16(s0, s1), !before && !after && cur >= 0);
    st_global_f2_if(before ? bnd0 : bnd0 + H, s0, s1, (before || after) && cur >= 0);
}

// Phase timing for tuning (scratch/phase2.py): compiled in only with -DGP_FWD2_PROF; the product build has none of it.
#ifdef GP_FWD2_PROF
#define PROF_DECL(cond) const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && (cond); long long prof_t = prof_on ? clock64() : 0
#define PROF_TICK(slot_)                                                          \
    do {                                                                          \
        if (prof_on) {                                                            \
            const long long now_ = clock64();                                     \
            atomicAdd(p.prof + (slot_), (unsigned long long)(now_ - prof_t));     \
            prof_t = now_;                                                        \
        }                                                                         \
    } while (0)
#define PROF_COUNT(slot_) do { if (prof_on) atomicAdd(p.prof + (slot_), 1ull); } while (0)
#else
#define PROF_DECL(cond) do { } while (0)
#define PROF_TICK(slot_) do { } while (0)
#define PROF_COUNT(slot_) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNT, 1)
edge_fwd2_kernel(const gp_mlp_fwd_args p, const __grid_constant__ Fwd2Maps maps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars& B = *reinterpret_cast<Bars*>(smem + kOffBar);
    const int tid = threadIdx.x;
    const int warp = warp_uniform(tid >> 5);
    const int lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_tiles = (p.rows + 127) >> 7;
    const int n_pairs = (n_tiles + 1) >> 1;
    const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int q_stride = 2 * n_clusters;

    // ---- prologue (parameters only): weights, barriers, TMEM
    if ((smem_u32(smem) & 1023u) != 0) __trap();
#ifdef GP_FWD2_PROF
    long long t_entry = 0;
    if (p.prof && blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
#define PROF_STAMP(slot_) do { if (p.prof && blockIdx.x == 0 && tid == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.prof[slot_] = (unsigned long long)(t_ - t_entry); } } while (0)
#else
#define PROF_STAMP(slot_) do { } while (0)
#endif
    for (int l = 0; l < 4; ++l) stage_weight(smem + kOffW + l * kWHalf, p.w[l] + (size_t)rank * 64 * H, 64, H);
    cp_async_commit();
    float* sbias = reinterpret_cast<float*>(smem + kOffBias);
    for (int i = tid; i < 5 * H; i += kNT) sbias[i] = i < 4 * H ? p.bias[i >> 7][i & 127] : p.norm_scale[i - 4 * H];
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&B.a_ready[s], 16);     // 8 epilogue warps x 2 CTAs
            mbar_init(&B.mma_done[s], 1);
            mbar_init(&B.tma_e[s], 1);
            mbar_init(&B.u_full[s], 8);       // epilogue warps of the slot
            mbar_init(&B.u_empty[s], 4);      // drain warps of the slot
            mbar_init(&B.g_full[s], maps.gather4 ? 2 : 64);      // cp.async: two producer warps x 32 lanes; TMA: one expect_tx per warp
            mbar_init(&B.g_empty[s], 8);      // consuming slot: each slot sees its own phases in order
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = B.tmem_slot;
    cluster_sync_all();               // both CTAs' weights, barriers and TMEM exist before any cross-CTA traffic
    tc_fence_after();
    // the first e tile of each slot is older than the previous kernel (it wrote P): into L2 before the wait
    if (tid < 2) {
        const int q0 = cid * 2 + tid;
        if (q0 < n_pairs) {
            const int R0_ = (2 * q0 + (int)rank) << 7;
            tma_prefetch_2d(&maps.e, 0, R0_);
            tma_prefetch_2d(&maps.e, 64, R0_);
        }
    }
    pdl_wait();
    pdl_launch_dependents();

    // Register budget per role: the launch gives every warp 72 registers per thread and setmaxnreg only moves
    // registers inside the CTA's own allocation (896 x 72): 16 x EPI + 8 x DRAIN + 4 x CTRL = 28 x 72 exactly.
    if (warp < kWDrain) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GP_FWD2_REGS_EPI));
    } else if (warp < kWMma) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GP_FWD2_REGS_DRAIN));
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GP_FWD2_REGS_CTRL));
    }

    if (warp < kWDrain) {
        // =========================================================================== epilogue slots
        const int s = warp >> 3;                 // slot
        const int hf = (warp >> 2) & 1;          // column half of the row this thread works on
        const int cb = hf * 64;
        const int row = (warp & 3) * 32 + lane;  // row of the tile == TMEM lane (warp w may touch lanes [32 (w % 4), +32))
        uint8_t* act = smem + kOffAct + s * kTile;
        uint8_t* ubuf = smem + kOffU + s * kTile;
        const uint8_t* gbuf = smem + kOffG;
        const uint32_t act_u = smem_u32(smem + kOffAct) + (uint32_t)s * kTile;
        const uint32_t tacc = tmem_addr(tmem_base, (warp & 3) * 32, s * kSlotCols);
        const uint32_t tact = tacc + kColAct;
        const uint32_t a_ready_leader = mapa_u32(smem_u32(&B.a_ready[s]), 0);
        uint32_t ph_mma = 0;
        const bool issuer = ((warp & 7) == 0) && lane == 0;      // the thread that issues this slot's bulk copies
        auto dst_of = [&](int qq) { return __ldg(p.idx0 + min(((2 * qq + (int)rank) << 7) + row, p.rows - 1)); };

        PROF_DECL(tid == 0);
        PROF_STAMP(24);      // ns from kernel entry to the start of the tile loop (prologue)
        int q = cid * 2 + s;
        if (q < n_pairs && issuer) {
            const int R0 = (2 * q + (int)rank) << 7;
            mbar_arrive_expect_tx(&B.tma_e[s], kTile);
            tma_load_2d(act_u, &maps.e, 0, R0, &B.tma_e[s]);
            tma_load_2d(act_u + 16384, &maps.e, 64, R0, &B.tma_e[s]);
        }
        int i_dst_next = q < n_pairs ? dst_of(q) : 0;
        for (int k = 0; q < n_pairs; ++k, q += q_stride) {
            const int R0 = (2 * q + (int)rank) << 7;
            const bool has_next = q + q_stride < n_pairs;
            // ---- accumulator pre-load: b1 + P[src][H:2H] (staged rows; the buffer goes back to the producers right
            //      after this pass) + P[dst][0:H] (sorted: read directly, requested before the wait)
            const gp_bf16* pd = p.init + (size_t)i_dst_next * p.ld_init + p.init_off0 + cb;
            uint4 dq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dq[i] = ldg16(pd + i * 8);
            PROF_COUNT(15);
            PROF_TICK(0);        // tile bookkeeping + requests
            mbar_wait(&B.g_full[s], k & 1);
            PROF_TICK(1);        // waiting for the staged sender rows
#pragma unroll
            for (int c2 = cb; c2 < cb + 64; c2 += 32) {
                float f[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(f + 4 * i) = *reinterpret_cast<const float4*>(sbias + c2 + 4 * i);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc8(*reinterpret_cast<const uint4*>(gbuf + sw128_off(128, row, c2 + i * 8)), f + 8 * i);
                    acc8(dq[(c2 - cb) / 8 + i], f + 8 * i);
                }
                tmem_st16(tacc + c2, *reinterpret_cast<const uint32_t(*)[16]>(f));
                tmem_st16(tacc + c2 + 16, *reinterpret_cast<const uint32_t(*)[16]>(f + 16));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&B.g_empty[s]);
            tmem_st_wait();
            tc_fence_before();
            PROF_TICK(2);        // accumulator pre-load
            mbar_wait(&B.tma_e[s], k & 1);       // this tile's e is in ACT (bulk copy issued a tile ago)
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_ready_leader);
            // the next tile's receiver index: requested here, so that its latency (and the store, should the value be
            // spilled) falls into the wait for the first MMA instead of into the chain
            if (has_next) i_dst_next = dst_of(q + q_stride);
            PROF_TICK(3);        // waiting for the e tile

            // ---- layers
#pragma unroll 1
            for (int l = 0; l < 4; ++l) {
                mbar_wait(&B.mma_done[s], ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
                PROF_TICK(4);    // waiting for the pair MMA (both CTAs' operands + issue + tensor time), x4
                if (l < 3) {
                    // hidden layer: relu(acc) -> bf16 -> TMEM (A operand of the next MMA); the accumulator restarts from
                    // the next layer's bias.  h2 (l == 1) additionally goes to ACT -- free since layer 0 -- and from
                    // there to global memory by one bulk store (it is the input of backward stage B).
                    const float* bn = sbias + (l + 1) * H;
                    const bool to_smem = (l == 1) && maps.save_h2;
#pragma unroll
                    for (int c2 = cb; c2 < cb + 64; c2 += 32) {
                        uint32_t v[32];
                        tmem_ld16(tacc + c2, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                        tmem_ld16(tacc + c2 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; c += 16) {
                            uint32_t b16[16];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 bq = *reinterpret_cast<const uint4*>(bn + c2 + c + 4 * i);
                                b16[4 * i] = bq.x; b16[4 * i + 1] = bq.y; b16[4 * i + 2] = bq.z; b16[4 * i + 3] = bq.w;
                            }
                            tmem_st16(tacc + c2 + c, b16);
                        }
                        uint32_t hp[16];          // 32 bf16 values: column j of the operand holds elements (2j, 2j+1)
#pragma unroll
                        for (int c = 0; c < 32; c += 2) hp[c >> 1] = pack_bf16_relu(__uint_as_float(v[c]), __uint_as_float(v[c + 1]));
                        tmem_st16(tact + (c2 >> 1), *reinterpret_cast<const uint32_t(*)[16]>(&hp[0]));
                        if (to_smem) {
#pragma unroll
                            for (int c = 0; c < 32; c += 8)
                                *reinterpret_cast<uint4*>(act + sw128_off(128, row, c2 + c)) =
                                    make_uint4(hp[c >> 1], hp[(c >> 1) + 1], hp[(c >> 1) + 2], hp[(c >> 1) + 3]);
                        }
                    }
                    PROF_TICK(17);   // TMEM -> ReLU -> TMEM, bias pre-store
                    tmem_st_wait();
                    tc_fence_before();
                    PROF_TICK(18);   // TMEM store wait
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(a_ready_leader);
                    PROF_TICK(5);    // arrive
                    if (to_smem) {
                        fence_async_smem();
                        named_sync(1 + s, kEpi);             // the whole h2 tile is in ACT
                        if (issuer) {                        // bulk-copy groups are per thread: always the same thread
                            tma_store_2d(&maps.h2, 0, R0, act_u);
                            tma_store_2d(&maps.h2, 64, R0, act_u + 16384);
                            tma_store_commit();
                        }
                        PROF_TICK(19);   // (l = 1) slot barrier + bulk store issue (off the MMA's critical path)
                    }
                } else {
                    // ---- RMSNorm (layers.py:104-129): u = scale * m / (||m||/sqrt(H) + 1e-8), rounded to bf16 once.
                    // ACT is free (layer 0 read e long ago; the bulk store of h2, if any, must have read it too): the next
                    // tile's e starts arriving now.
                    if (has_next && issuer) {
                        if (maps.save_h2) tma_store_wait_read<0>();
                        const int Rn = (2 * (q + q_stride) + (int)rank) << 7;
                        mbar_arrive_expect_tx(&B.tma_e[s], kTile);
                        tma_load_2d(act_u, &maps.e, 0, Rn, &B.tma_e[s]);
                        tma_load_2d(act_u + 16384, &maps.e, 64, Rn, &B.tma_e[s]);
                    }
                    // Both threads of a row read the WHOLE row for the sum of squares (tensor-memory reads are cheap; an
                    // exchange would cost a slot barrier and shared memory that is not there) and then scale their own half.
                    // (two 64-column partial sums added at the end: the summation order of the single-CTA kernel, whose
                    //  two column halves belong to different threads -- the kernels stay bit-identical)
                    float ssh[2] = {0.f, 0.f};
#pragma unroll
                    for (int c2 = 0; c2 < H; c2 += 32) {
                        uint32_t v[32];
                        tmem_ld16(tacc + c2, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                        tmem_ld16(tacc + c2 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float m = __uint_as_float(v[c]);
                            ssh[c2 >> 6] = fmaf(m, m, ssh[c2 >> 6]);
                        }
                    }
                    const float ss = ssh[0] + ssh[1];
                    const float rinv = 1.f / (sqrtf(ss * (1.f / H)) + 1e-8f);
                    PROF_TICK(6);    // norm: sum of squares
                    if (k > 0) mbar_wait(&B.u_empty[s], (k - 1) & 1);     // the drain warps are done with the previous u
                    PROF_TICK(7);    // waiting for the drain warps
                    const float* sc = sbias + 4 * H;
#pragma unroll
                    for (int c2 = cb; c2 < cb + 64; c2 += 32) {
                        uint32_t v[32];
                        tmem_ld16(tacc + c2, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                        tmem_ld16(tacc + c2 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            const float4 g0 = *reinterpret_cast<const float4*>(sc + c2 + c), g1 = *reinterpret_cast<const float4*>(sc + c2 + c + 4);
                            float u[8];
                            u[0] = g0.x * (__uint_as_float(v[c]) * rinv);     u[1] = g0.y * (__uint_as_float(v[c + 1]) * rinv);
                            u[2] = g0.z * (__uint_as_float(v[c + 2]) * rinv); u[3] = g0.w * (__uint_as_float(v[c + 3]) * rinv);
                            u[4] = g1.x * (__uint_as_float(v[c + 4]) * rinv); u[5] = g1.y * (__uint_as_float(v[c + 5]) * rinv);
                            u[6] = g1.z * (__uint_as_float(v[c + 6]) * rinv); u[7] = g1.w * (__uint_as_float(v[c + 7]) * rinv);
                            *reinterpret_cast<uint4*>(ubuf + sw128_off(128, row, c2 + c)) = pack8(u);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&B.u_full[s]);
                    PROF_TICK(8);    // norm: scale + write u
                }
            }
        }
        if (maps.save_h2 && issuer) tma_store_wait_all();
        PROF_STAMP(25);      // ... to the end of this thread's tile loop
    } else if (warp < kWMma) {
        // =========================================================================== drain: segment sum + e' = e + u
        const int s = (warp - kWDrain) >> 2;
        const int d = (tid - 2 * kEpi) & (kDrain - 1);
        const uint8_t* ubuf = smem + kOffU + s * kTile;
        constexpr int KC = H / 8;
        PROF_DECL(tid == 2 * kEpi);
        int q = cid * 2 + s;
        for (int k = 0; q < n_pairs; ++k, q += q_stride) {
            const int R0 = (2 * q + (int)rank) << 7;
            // the residual rows (the tile this kernel read as its layer-0 operand: L2 hits) are requested before the wait
            uint4 rq[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int i = d + jj * kDrain;
                rq[jj] = ldg16(p.resid + (size_t)min(R0 + i / KC, p.rows - 1) * p.ld_out + (i % KC) * 8);
            }
            PROF_TICK(9);        // drain: residual requests
            mbar_wait(&B.u_full[s], k & 1);
            PROF_TICK(10);       // drain: waiting for u
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                uint4 y[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int i = d + (hb * 8 + jj) * kDrain;
                    y[jj] = add8_bf16(*reinterpret_cast<const uint4*>(ubuf + sw128_off(128, i / KC, (i % KC) * 8)), rq[jj]);
                }
                if (hb == 0) {       // second half of the residual rows: requested before the stores of the first
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int i = d + (8 + jj) * kDrain;
                        rq[jj] = ldg16(p.resid + (size_t)min(R0 + i / KC, p.rows - 1) * p.ld_out + (i % KC) * 8);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int i = d + (hb * 8 + jj) * kDrain;
                    const int r = i / KC, ch = i % KC;
                    if (R0 + r < p.rows) *reinterpret_cast<uint4*>(p.y_bf16 + (size_t)(R0 + r) * p.ld_out + ch * 8) = y[jj];
                }
            }
            PROF_TICK(11);       // drain: e' = e + u
            // receiver-sorted segment sum of bf16(u): same partition as the single-CTA kernel (sub-tiles of 32 rows x
            // column pairs), two units per thread
#pragma unroll 1
            for (int un = 0; un < 2; ++un) {
                const WalkIds ids = walk_load_ids(p.seg_id, p.rows, R0, d + un * kDrain);
                segment_walk_unit(ubuf, ids, R0, d + un * kDrain, p.seg_bnd, p.seg_out_bf16);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&B.u_empty[s]);
            PROF_TICK(12);       // drain: segment walk
        }
    } else if (warp < kWProd) {
        // =========================================================================== MMA issue (leader CTA)
        if (rank == 0) {
            const int s = warp - kWMma;
            const uint32_t act_u = smem_u32(smem + kOffAct) + (uint32_t)s * kTile;
            const uint32_t w_u = smem_u32(smem + kOffW);
            const uint32_t idesc = idesc_bf16(H, false, false, 256);
            const uint32_t tacc = tmem_base + s * kSlotCols, tact = tacc + kColAct;
            uint32_t ph = 0;
            for (int q = cid * 2 + s; q < n_pairs; q += q_stride) {
#pragma unroll 1
                for (int l = 0; l < 4; ++l) {
                    mbar_wait_cluster(&B.a_ready[s], ph);
                    ph ^= 1;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bd = desc_kmajor(w_u + l * kWHalf, 64, 0);
                        // the accumulator already holds the bias (and, for layer 0, the gathered rows)
                        if (l == 0) {        // A = the e tile in shared memory
                            const uint64_t ad = desc_kmajor(act_u, 128, 0);
#pragma unroll
                            for (int ks = 0; ks < H / 16; ++ks) {
                                const uint32_t ko = (ks & 3) * 2;
                                mma2_ss(tacc, ad + (uint64_t)((ks >> 2) * 1024u + ko), bd + (uint64_t)((ks >> 2) * 512u + ko), idesc, 1u);
                            }
                        } else {             // A = the previous layer's bf16 activation in tensor memory (8 columns per k-step)
#pragma unroll
                            for (int ks = 0; ks < H / 16; ++ks) {
                                const uint32_t ko = (ks & 3) * 2;
                                mma2_ts(tacc, tact + ks * 8, bd + (uint64_t)((ks >> 2) * 512u + ko), idesc, 1u);
                            }
                        }
                        mma2_commit_both(&B.mma_done[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =========================================================================== gather producers: P[src] rows -> G
        // two warps, 64 rows each; the rows are pulled into L2 before the staging buffer is free, so the copies that
        // follow the hand-over are L2 hits
        const int pw = warp - kWProd;
        const uint32_t g_u = smem_u32(smem + kOffG);
        PROF_DECL(lane == 0 && pw == 0);
        for (int j = 0;; ++j) {
            const int q = cid * 2 + (j & 1) + (j >> 1) * q_stride;
            if (q >= n_pairs) break;
            const int R0 = ((2 * q + (int)rank) << 7) + pw * 64;
            int my[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) my[i] = __ldg(p.idx1 + min(R0 + lane + 32 * i, p.rows - 1));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const gp_bf16* rp = p.init + (size_t)my[i] * p.ld_init + p.init_off1;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 64));
            }
            PROF_TICK(13);       // producer: issue (previous tile) + index loads
            if (j > 0) mbar_wait(&B.g_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);      // the previous tile's rows have been consumed
            PROF_TICK(14);       // producer: waiting for the staging buffer
            if (maps.gather4) {
                // TMA row gather: lane = (column block, group of four rows); one instruction lands 4 x 128 bytes in the
                // tile layout, the copy engine does the address arithmetic and the swizzle
                if (lane == 0) mbar_arrive_expect_tx(&B.g_full[j & 1], 64 * H * 2);
                const int gq = lane & 15, blk = lane >> 4;
                const int rl = 4 * gq;                                     // first of this lane's four rows within the warp's 64
                int ix[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {        // (both halves are shuffled: the selection belongs to the receiving lane)
                    const int lo = __shfl_sync(0xffffffffu, my[0], (rl + t) & 31), hi = __shfl_sync(0xffffffffu, my[1], (rl + t) & 31);
                    ix[t] = rl < 32 ? lo : hi;
                }
                const int i0 = ix[0], i1 = ix[1], i2 = ix[2], i3 = ix[3];
                tma_gather4(g_u + blk * 16384 + (pw * 64 + rl) * 128, &maps.psrc, blk * 64, i0, i1, i2, i3, &B.g_full[j & 1]);
                continue;
            }
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj) {
                const int rl = 2 * jj + (lane >> 4), ch = lane & 15;       // row within this warp's 64
                const int idx = __shfl_sync(0xffffffffu, jj < 16 ? my[0] : my[1], rl & 31);
                cp_async16(g_u + sw128_off(128, pw * 64 + rl, ch * 8), p.init + (size_t)idx * p.ld_init + p.init_off1 + ch * 8);
            }
            cp_async_arrive_noinc(&B.g_full[j & 1]);
        }
        cp_async_wait<0>();
    }

    // ---- teardown: nobody leaves while the peer may still address this CTA
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    PROF_STAMP(26);          // ... to the exit of CTA 0
}
}  // namespace

namespace gp {
// 1 = launched, 0 = not this kernel's case, < 0 = error
int try_edge_fwd2(const gp_mlp_fwd_args& a, cudaStream_t st) {
    static const bool off = getenv("GP_EDGE_FWD_V1") != nullptr;
    if (off) return 0;
    if (a.n_layers != 4 || a.ka != H || !a.init || !a.two_inits || !a.idx0 || !a.idx1 || !a.a_bf16 || !a.norm_scale || !a.resid ||
        !a.y_bf16 || !a.seg_id || !a.seg_out_bf16 || !a.seg_bnd || a.seg_id != a.idx0 || a.save_h1 || a.save_h3 || a.lda != H ||
        a.ld_out != H || a.resid != a.a_bf16)
        return 0;
    for (int l = 0; l < 4; ++l)
        if (a.k[l] != H || a.n[l] != H || !a.bias[l]) return 0;
    Fwd2Maps maps;
    memset(&maps, 0, sizeof(maps));
    if (!gp::tma_map_2d(&maps.e, a.a_bf16, a.rows, H, a.lda)) return 0;
    if (a.save_h2) {
        if (!gp::tma_map_2d(&maps.h2, a.save_h2, a.rows, H, H)) return 0;
        maps.save_h2 = 1;
    }
    // sender rows by TMA tile::gather4 (GP_EDGE_FWD_GATHER4=0 keeps the cp.async producers).  The row bound of the map is
    // an upper limit only (the argument block does not carry the node count; every index is valid by contract).
    static const bool g4_off = getenv("GP_EDGE_FWD_GATHER4") != nullptr && getenv("GP_EDGE_FWD_GATHER4")[0] == '0';
    if (!g4_off && gp::tma_map_rows(&maps.psrc, a.init + a.init_off1, 1ll << 30, H, a.ld_init)) maps.gather4 = 1;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(edge_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr_set = true;
    }
    const int n_tiles = (a.rows + 127) / 128, n_pairs = (n_tiles + 1) / 2;
    int clusters = (n_pairs + 1) / 2;
    if (clusters > gp::sm_count() / 2) clusters = gp::sm_count() / 2;
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kNT);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gp::launch_overlap() ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, edge_fwd2_kernel, a, maps);
    if (e != cudaSuccess) {
        gp::set_error("edge_fwd2 launch failed: %s", cudaGetErrorString(e));
        return -2;
    }
    return 1;
}
}  // namespace gp
