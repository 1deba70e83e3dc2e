// mlp_bwd.cu -- two-layer backward stage of the fused MLP on tcgen05 (gp_mlp_bwd_stage).
//
// One persistent CTA per SM, 512 threads at H = 128 (256 below): one 128-row tile at a time.
// Resident in shared memory: the two packed weights (each used K-major for the recompute and
// MN-major for dgrad), a tile of ones (bias / scale gradients as delta^T . 1 on the tensor core)
// and the activation buffers Ain / Ha / Db / Q of the tile.  Contiguous tiles arrive and leave by
// TMA (one elected thread, mbarrier), gathered rows by 16-byte cp.async; the next tile's row ids are
// loaded and its data prefetched into L2 while this tile computes.  TMEM holds the working
// accumulator plus the weight-gradient accumulators, which persist across all tiles of the CTA
// (column map below); each CTA dumps them once, as coalesced rows, into its partial block.
// Thread (row = tid & 127, part = tid >> 7) owns row `row` of the tile and one part (1/2 or 1/4) of its
// columns, so ReLU masks, the RMSNorm backward and residuals are thread-local apart from one
// two-float exchange between the parts.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tile_util.cuh"

namespace {
using namespace gp;

// TMEM columns.  General kernel: ACC 0, dWb 128, dWa 256, dbb 384, dba 400, dscale 416.  Specialised
// instantiations (H = ka = nb = 128) compute every bias gradient as 16 extra columns of its weight
// gradient (the B operand is [activation tile | ones tile], N = 144), one MMA chain instead of two:
// ACC 0, [dWb | dbb] 128..271, [dWa | dba] 272..415, dscale 416.
constexpr uint32_t kColAcc = 0, kColDWB = 128, kColDBA = 400, kColDSC = 416;
template <bool F> struct BwdCols { static constexpr uint32_t DWA = F ? 272 : 256, DBB = F ? 256 : 384; };

struct BwdLayout {
    int off_dwb, off_dwa, off_dbb, off_dba, off_dsc, stride;
};
__host__ __device__ inline BwdLayout bwd_layout(int H, int ka, int nb) {
    BwdLayout L;
    L.off_dwb = 0;
    L.off_dwa = nb * H;
    L.off_dbb = L.off_dwa + H * ka;
    L.off_dba = L.off_dbb + nb;
    L.off_dsc = L.off_dba + H;
    L.stride = L.off_dsc + H;
    return L;
}

// Tile-shaped global tensors that meet the TMA alignment rules are moved by the copy engine (one
// elected thread, 16 KB per instruction, straight into / out of the SW128 layout); the flags say
// which descriptors are valid, everything else falls back to per-thread 16-byte copies.
struct BwdMaps {
    CUtensorMap ain, db, gy, resid, da_out, out, ha;
    uint32_t use;
};
enum : uint32_t { kMapAin = 1, kMapDb = 2, kMapGy = 4, kMapResid = 8, kMapDaOut = 16, kMapOut = 32, kMapHa = 64 };

// Threads per CTA: 128 rows x NPART column parts (4 parts = 16 warps at H = 128, where the epilogues are
// latency-bound and need the extra warps; 2 parts for narrower layers).
template <int H>
struct BwdCfg {
    static constexpr int NPART = H >= 128 ? 4 : 2;
    static constexpr int NT = 128 * NPART;
};

// MODE 0 is the general kernel.  MODE 1 / 2 are the instantiations for the two stages of the processor's
// edge MLP at H = 128 (1: layers 3-4 with the RMSNorm backward and the gathered receiver gradient;
// 2: layers 1-2 with the two gathered pre-activation sources, the stored delta_1 and its segment
// sum), MODE 3 / 4 those of the node MLP (3: fp32 upstream gradient read per row; 4: bf16 aggregate as
// operand, bf16 d_agg output), with every option -- including which tensors move by TMA -- fixed
// at compile time.  The
// general kernel is 100 KB of code that each tile streams through the instruction cache; the
// specialised ones keep only their own path.
template <int H, int MODE>
__global__ void __launch_bounds__(BwdCfg<H>::NT, 1) mlp_bwd_kernel(const gp_mlp_bwd_args p,
                                                                     const __grid_constant__ BwdMaps maps) {
    constexpr int NPART = BwdCfg<H>::NPART;
    constexpr int NT = BwdCfg<H>::NT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t mma_bar, tma_bar, wg_bar;     // wg_bar: completion of the weight-gradient MMA chains
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int row = tid & 127, part = tid >> 7;
#ifdef GP_MLP_PROF
    long long t_entry = 0;
    if (p.prof && blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry));
#endif
    // option tables of the specialised instantiations:    general, edge B, edge A, node B, node A
    constexpr bool F = MODE != 0, F1 = MODE == 1;
    constexpr bool kNorm[5]   = {false, true,  false, true,  false};
    constexpr bool kTAin[5]   = {false, true,  true,  true,  true};    // a_in tile by TMA
    constexpr bool kTDb[5]    = {false, false, true,  false, true};
    constexpr bool kTGy[5]    = {false, true,  false, false, false};
    constexpr bool kTRes[5]   = {false, false, true,  false, false};
    constexpr bool kTDa[5]    = {false, false, true,  false, true};
    constexpr bool kTOut[5]   = {false, true,  true,  true,  true};    // bf16 output tile through shared memory + TMA
    constexpr bool kGather[5] = {false, true,  false, false, false};
    constexpr bool kSeg[5]    = {false, false, true,  false, false};
    constexpr bool kMask[5]   = {false, true,  false, true,  false};
    constexpr bool kInit[5]   = {false, false, true,  false, true};
    constexpr bool kTwo[5]    = {false, false, true,  false, false};
    const bool t_ain = F ? kTAin[MODE] : bool(maps.use & kMapAin), t_db = F ? kTDb[MODE] : bool(maps.use & kMapDb);
    const bool t_gy = F ? kTGy[MODE] : bool(maps.use & kMapGy), t_res = F ? kTRes[MODE] : bool(maps.use & kMapResid);
    const bool t_da = F ? kTDa[MODE] : bool(maps.use & kMapDaOut), t_out = F ? kTOut[MODE] : bool(maps.use & kMapOut);
    const bool t_ha = !F && (maps.use & kMapHa);
    const bool ha_given = !F && p.ha_saved != nullptr;       // h_a comes from memory: no gather, no recompute
    // receiver-indexed rows are added to gy: bf16 rows or none in the specialised stage B (the edge encoder's
    // stage B is the same kernel without them), any form in the general kernel
    const bool f_gather = F ? (kGather[MODE] && p.gy_gather_bf16 != nullptr) : (p.gy_gather != nullptr || p.gy_gather_bf16 != nullptr);
    const bool f_gbf = F ? kGather[MODE] : p.gy_gather_bf16 != nullptr;
    const bool f_seg = F ? kSeg[MODE] : p.seg_id != nullptr, f_da = F ? kTDa[MODE] : p.delta_a_out != nullptr;
    const bool f_din = F || p.need_din, f_mask = F ? kMask[MODE] : bool(p.mask_by_ain);
    const bool f_resid = F ? kTRes[MODE] : p.out_resid != nullptr;
    const bool f_obf = F ? kTOut[MODE] : p.out_bf16 != nullptr;          // bf16 (else fp32) d_in output
    const bool f_gyf32 = F ? MODE == 3 : p.gy_f32 != nullptr;
    const int warp = warp_uniform(tid >> 5);
    const int ka = F ? H : p.ka, nb = F ? H : p.nb;
    const bool norm = F ? kNorm[MODE] : p.mode == 1;

    // ---- carve
    uint32_t off = 0;
    uint8_t* wa_t = smem + off;  off += ((ka + 63) >> 6) * H * 128;          // [H rows][ka]
    uint8_t* wb_t = smem + off;  off += ((H + 63) >> 6) * nb * 128;          // [nb rows][H]
    // Tile buffers.  General kernel: ain, ha, db, qb, ones back to back.  Specialised instantiations: the two
    // 64-column blocks of ha are 48 KB apart and the ones tile sits where "block 2" of both ha and ain
    // falls (ha0 | db | ha1 | ain | ones | qb), so [ha | ones] and [ain | ones] are valid 144-column MN-major
    // operands with block strides of 48 KB and 16 KB.
    constexpr uint32_t kHaStride = F ? 49152u : 16384u;     // bytes between the 64-column blocks of ha
    constexpr uint32_t HR = kHaStride / 128u;               // the same as a "rows" argument of the tile helpers
    uint8_t* const tiles = smem + off;
    uint8_t* ain = tiles + (F ? 65536 : 0);
    uint8_t* ha = tiles + (F ? 0 : kBufBytes);
    uint8_t* db = tiles + (F ? 16384 : 2 * kBufBytes);
    uint8_t* ones = tiles + (F ? 98304 : 3 * kBufBytes + ((norm || t_res) ? kBufBytes : 0));
    uint8_t* qb = tiles + (F ? 114688 : 3 * kBufBytes);     // NORM: du / q tile; else residual tile (absent in node stage A)
    off += 3 * kBufBytes + ((norm || t_res) ? kBufBytes : 0) + 128 * 128;
    constexpr uint32_t kColDWA = BwdCols<F>::DWA, kColDBB = BwdCols<F>::DBB;
    float* s_ba = reinterpret_cast<float*>(smem + off);  off += 128 * 4;
    float* s_bb = reinterpret_cast<float*>(smem + off);  off += 128 * 4;
    float* s_g = reinterpret_cast<float*>(smem + off);   off += 128 * 4;
    float* s_red = reinterpret_cast<float*>(smem + off); off += 2 * NPART * 128 * 4;   // [2 quantities][NPART][128]
    int* sseg = reinterpret_cast<int*>(smem + off);

    // ---- one-time staging
    stage_weight(wa_t, p.wa, H, ka);
    stage_weight(wb_t, p.wb, nb, H);
    cp_async_commit();
    for (int i = tid; i < 128 * 8; i += NT) {       // ones tile: first 16 columns of every row = 1.0
        const int r = i >> 3, ch = i & 7;
        const uint32_t one2 = 0x3F803F80u;
        *reinterpret_cast<uint4*>(ones + sw128_chunk_off(r, ch)) =
            ch < 2 ? make_uint4(one2, one2, one2, one2) : make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < 128; i += NT) {
        s_ba[i] = (i < H && p.ba) ? p.ba[i] : 0.f;
        s_bb[i] = (i < nb && p.bb) ? p.bb[i] : 0.f;
        s_g[i] = (i < H && p.norm_scale) ? p.norm_scale[i] : 1.f;
    }
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        mbar_init(&wg_bar, 1);
        mbar_init(&tma_bar, 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&tmem_slot, 512);
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tlane = tmem_addr(tmem, (row >> 5) * 32, 0);
    const uint32_t ain_s = smem_u32(ain), ha_s = smem_u32(ha), db_s = smem_u32(db), qb_s = smem_u32(qb);
    const uint32_t wa_s = smem_u32(wa_t), wb_s = smem_u32(wb_t), ones_s = smem_u32(ones);
    const uint32_t lbo_h = (H >= 128) ? 16384u : 0u;      // M=128 MN-major A with < 128 valid columns: alias block
    const uint32_t lbo_ha = F ? kHaStride : lbo_h;         // the same for an A operand that lives in ha
    const uint32_t lbo_nb = (nb >= 128) ? 16384u : 0u;
    uint32_t phase = 0, tphase = 0;
    const bool has_init = F ? kInit[MODE] : (p.init != nullptr && !ha_given);
    const int n_tiles = (p.rows + 127) >> 7;
    constexpr int CH = H / NPART;                          // columns per thread
    const int cb = part * CH;

    uint32_t wphase = 0;
    auto wait_wg = [&]() {
        mbar_wait(&wg_bar, wphase);
        wphase ^= 1;
        tc_fence_after();
    };
    auto wait_mma = [&]() {
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
    };
    auto publish = [&]() {   // smem tiles written by threads -> visible to the MMA issuer
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };

    constexpr int KC = H / 8;                 // 16-byte chunks per H-wide row
    constexpr int CPT = 128 * KC / NT;       // chunks per thread in a row-major tile copy
    const bool stage1 = F ? kTwo[MODE] : p.two_inits != 0;     // which pre-activation source is gathered through shared memory
    const int32_t* sidx = stage1 ? p.idx1 : p.idx0;
    const int soff = stage1 ? p.init_off1 : p.init_off0;
    const bool du_smem = F ? kTGy[MODE] : (norm && p.gy_bf16 != nullptr);     // upstream gradient tile staged in qb (+ gathered rows in db)

#ifdef GP_MLP_PROF
    const bool prof = p.prof != nullptr && tid == 0;      // phase timing (scratch/phase*.py): lib/libgp_b200_prof.so only
#define BWD_STAMP(slot_) do { if (prof && blockIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.prof[slot_] = (unsigned long long)(t_ - t_entry); } } while (0)
#else
#define BWD_STAMP(slot_) do { } while (0)
#endif
#ifndef GP_MLP_PROF
    constexpr bool prof = false;                        // the product build carries no profiling code
#endif
    long long tk = 0;
    auto tick = [&](int slot) {
        if (prof) {
            const long long now = clock64();
            atomicAdd(p.prof + slot, (unsigned long long)(now - tk));
            tk = now;
        }
    };
    // indices of a tile, prefetched one tile ahead so P0 starts with the rows, not with their ids
    int ridx[CPT], gidx[CPT], i0n = 0;
    auto load_idx = [&](int tile_) {
        const int R0_ = tile_ << 7;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int gr = min(R0_ + (tid + j * NT) / KC, p.rows - 1);
            ridx[j] = (has_init && sidx) ? __ldg(sidx + gr) : gr;
            gidx[j] = (du_smem && f_gather && (F1 || p.gy_idx)) ? __ldg(p.gy_idx + gr) : gr;
        }
        if (has_init && stage1) {
            const int r = min(R0_ + row, p.rows - 1);
            i0n = (MODE == 2 || p.idx0) ? __ldg(p.idx0 + r) : r;
        }
    };
    // everything the P0 of the tile starting at row Rn will ask for goes to L2 (bulk tiles by the copy engine, gathered
    // rows one prefetch per 128-byte line; row ids from the last load_idx), so that P0 waits on L2 only
    auto prefetch_inputs = [&](int Rn) {
        if (warp == 0 && elect_one()) {
            if (t_ain) for (int b = 0; b < (ka + 63) >> 6; ++b) tma_prefetch_2d(&maps.ain, b * 64, Rn);
            if (t_db) for (int b = 0; b < (nb + 63) >> 6; ++b) tma_prefetch_2d(&maps.db, b * 64, Rn);
            if (t_gy) for (int b = 0; b < (H + 63) >> 6; ++b) tma_prefetch_2d(&maps.gy, b * 64, Rn);
            if (t_res) for (int b = 0; b < (ka + 63) >> 6; ++b) tma_prefetch_2d(&maps.resid, b * 64, Rn);
            if (t_ha) for (int b = 0; b < (H + 63) >> 6; ++b) tma_prefetch_2d(&maps.ha, b * 64, Rn);
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = tid + j * NT;
            if ((i & 7) == 0) {         // chunk 0 of a 128-byte line
                if (has_init)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.init + (size_t)ridx[j] * p.ld_init + soff + (i % KC) * 8));
                if (du_smem && f_gather && f_gbf) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gy_gather_bf16 + (size_t)gidx[j] * H + (i % KC) * 8));
                } else if (du_smem && f_gather) {       // fp32 rows: two lines per 8-chunk group
                    const float* gp_ = p.gy_gather + (size_t)gidx[j] * H + (i % KC) * 8;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gp_));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(gp_ + 32));
                }
            }
        }
        if (MODE == 3)      // the tile's fp32 gradient rows: 512 lines, one per thread
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gy_f32 + (size_t)min(Rn + (tid >> 2), p.rows - 1) * p.ld_gy + (tid & 3) * 32));
        if (has_init && stage1 && tid < 128)
            for (int c = 0; c < H; c += 64)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.init + (size_t)i0n * p.ld_init + p.init_off0 + c));
    };
    if ((int)blockIdx.x < n_tiles) {
        // the first tile of this CTA: ids and L2 prefetch before the wait below -- these requests overlap the tail of the
        // previous kernel (what it is still writing reaches L2 anyway: a prefetched line is never stale there)
        load_idx(blockIdx.x);
        prefetch_inputs((int)blockIdx.x << 7);
    }
    // everything above is independent of the previous kernel in the stream (parameters, graph layout, prefetches); from here
    // on its outputs are read, and the next kernel may start its own prologue
    pdl_wait();
    pdl_launch_dependents();
    BWD_STAMP(16);        // ns from kernel entry: prologue done (weights staged, previous kernel finished)
    bool first = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        if (prof) { tk = clock64(); atomicAdd(p.prof + 15, 1ull); }
        const int R0 = tile << 7;
        const int grow = R0 + row;
        const bool valid = grow < p.rows;
        const int crow = valid ? grow : p.rows - 1;
        const uint32_t acc_flag = first ? 0u : 1u;

        // ---- P0: stage inputs.  Every tile-shaped transfer uses row-major 16-byte chunks (8 lanes per
        //      cache line); tiles are transposed to the row-per-thread mapping through shared memory.
        const uint32_t tma_blocks = (t_ain ? (ka + 63) >> 6 : 0) + (t_db ? (nb + 63) >> 6 : 0) + (t_gy ? (H + 63) >> 6 : 0) +
                                    (t_res ? (ka + 63) >> 6 : 0) + (t_ha ? (H + 63) >> 6 : 0);
        if (tma_blocks && warp == 0 && elect_one()) {
            // the bulk stores of the previous tile have read their shared-memory sources (the output
            // staging buffer is only ever refilled by this thread's own loads below)
            tma_store_wait_read<0>();
            mbar_arrive_expect_tx(&tma_bar, tma_blocks * 16384u);
            if (t_ain)
                for (int b = 0; b < (ka + 63) >> 6; ++b) tma_load_2d(ain_s + b * 16384, &maps.ain, b * 64, R0, &tma_bar);
            if (t_db)    // rows past the end arrive as zeros, so they add nothing to the weight gradients
                for (int b = 0; b < (nb + 63) >> 6; ++b) tma_load_2d(db_s + b * 16384, &maps.db, b * 64, R0, &tma_bar);
            if (t_gy)
                for (int b = 0; b < (H + 63) >> 6; ++b) tma_load_2d(qb_s + b * 16384, &maps.gy, b * 64, R0, &tma_bar);
            if (t_res)
                for (int b = 0; b < (ka + 63) >> 6; ++b) tma_load_2d(qb_s + b * 16384, &maps.resid, b * 64, R0, &tma_bar);
            if (t_ha)
                for (int b = 0; b < (H + 63) >> 6; ++b) tma_load_2d(ha_s + b * kHaStride, &maps.ha, b * 64, R0, &tma_bar);
        }
        if (ha_given && !t_ha) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = tid + j * NT;
                const int r = i / KC, ch = i % KC;
                if (R0 + r < p.rows)
                    cp_async16(ha_s + sw128_off(HR, r, ch * 8), p.ha_saved + (size_t)(R0 + r) * H + ch * 8);
                else
                    *reinterpret_cast<uint4*>(ha + sw128_off(HR, r, ch * 8)) = make_uint4(0, 0, 0, 0);
            }
        }
        if (!t_ain) stage_rows(ain, p.a_bf16, F ? nullptr : p.a_f32, ka, p.lda, R0, p.rows, tid, NT);
        if (!norm && !t_db) {   // delta_b given: zero rows past the end so they add nothing to the weight gradients
            const int kc = nb >> 3;
            for (int i = tid; i < 128 * kc; i += NT) {
                const int r = i / kc, ch = i - r * kc;
                if (R0 + r < p.rows)
                    cp_async16(db_s + sw128_off(128, r, ch * 8), p.delta_b + (size_t)(R0 + r) * p.ld_db + ch * 8);
                else
                    *reinterpret_cast<uint4*>(db + sw128_off(128, r, ch * 8)) = make_uint4(0, 0, 0, 0);
            }
        }
        if (du_smem && !t_gy) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = tid + j * NT;
                const int r = i / KC, ch = i % KC;
                cp_async16(qb_s + sw128_off(128, r, ch * 8), p.gy_bf16 + (size_t)min(R0 + r, p.rows - 1) * p.ld_gy + ch * 8);
            }
        }
        if (MODE == 3) {
            // node stage B: the fp32 upstream gradient tile, coalesced, into the two buffers that are idle until E2 writes
            // delta_b / q into them: columns 0-63 -> qb, 64-127 -> db, rows of 256 bytes with their sixteen 16-byte chunks
            // swizzled by the row (thread = row reads of E2 are then conflict-free).  Read row-per-thread straight from
            // global memory, these 64 KB cost 11 k of the tile's 22 k cycles (32 lines per load instruction).
#pragma unroll
            for (int j = 0; j < 4096 / NT; ++j) {
                const int i = tid + j * NT;
                const int r = i >> 5, ch = i & 31;
                const uint32_t dst = ((ch & 16) ? db_s : qb_s) + r * 256 + (((ch & 15) ^ (r & 15)) << 4);
                cp_async16(dst, p.gy_f32 + (size_t)min(R0 + r, p.rows - 1) * p.ld_gy + ch * 4);
            }
        }
        if (has_init) {   // gathered pre-activation rows of the randomly indexed source -> ha
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = tid + j * NT;
                const int r = i / KC, ch = i % KC;
                cp_async16(ha_s + sw128_off(HR, r, ch * 8), p.init + (size_t)ridx[j] * p.ld_init + soff + ch * 8);
            }
        }
        if (du_smem && f_gather && f_gbf) {   // receiver-indexed bf16 rows -> db
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = tid + j * NT;
                const int r = i / KC, ch = i % KC;
                cp_async16(db_s + sw128_off(128, r, ch * 8), p.gy_gather_bf16 + (size_t)gidx[j] * H + ch * 8);
            }
        } else if (du_smem && f_gather) {   // receiver-indexed fp32 rows, rounded to bf16 into db; loads batched by 4 chunks
#pragma unroll
            for (int j0 = 0; j0 < CPT; j0 += 4) {
                float4 u[8];
#pragma unroll
                for (int j = 0; j < 4 && j0 + j < CPT; ++j) {
                    const int ch = (tid + (j0 + j) * NT) % KC;
                    const float4* sp = reinterpret_cast<const float4*>(p.gy_gather + (size_t)gidx[j0 + j] * H + ch * 8);
                    u[2 * j] = __ldg(sp);
                    u[2 * j + 1] = __ldg(sp + 1);
                }
#pragma unroll
                for (int j = 0; j < 4 && j0 + j < CPT; ++j) {
                    const int i = tid + (j0 + j) * NT;
                    const int r = i / KC, ch = i % KC;
                    *reinterpret_cast<uint4*>(db + sw128_off(128, r, ch * 8)) =
                        make_uint4(pack_bf16(u[2 * j].x, u[2 * j].y), pack_bf16(u[2 * j].z, u[2 * j].w),
                                   pack_bf16(u[2 * j + 1].x, u[2 * j + 1].y), pack_bf16(u[2 * j + 1].z, u[2 * j + 1].w));
                }
            }
        }
        cp_async_commit();
        int sid_me = -1, sid_prev = -1, sid_next = -1;
        if (f_seg && tid < 128) {
            if (valid) sid_me = __ldg(p.seg_id + grow);
            if (row == 0) {
                if (R0 > 0) sid_prev = __ldg(p.seg_id + R0 - 1);
                if (R0 + 128 < p.rows) sid_next = __ldg(p.seg_id + R0 + 128);
            }
        }
        uint4 dq[CH / 8];                                   // this thread's part of the receiver-indexed row
        if (has_init && stage1) {
            const gp_bf16* dp = p.init + (size_t)i0n * p.ld_init + p.init_off0 + cb;
#pragma unroll
            for (int i = 0; i < CH / 8; ++i) dq[i] = ldg16(dp + i * 8);
        }
        tick(0);      // P0 issue
        cp_async_wait<0>();
        if (tma_blocks) {
            mbar_wait(&tma_bar, tphase);
            tphase ^= 1;
        }
        __syncthreads();
        tick(1);      // P0 wait
        // accumulator pre-load: bias ba (+ the gathered pre-activation rows), so E1 adds nothing;
        // with h_a given the first MMA is P2 (NORM), which accumulates onto bb
#pragma unroll
        for (int c = 0; c < CH; c += 16) {
            if (ha_given && !norm) break;
            float f[16];
            const float* bsrc = ha_given ? s_bb : s_ba;
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(f + j) = *reinterpret_cast<const float4*>(bsrc + cb + c + j);
            if (has_init) {
                acc8(*reinterpret_cast<const uint4*>(ha + sw128_off(HR, row, cb + c)), f);
                acc8(*reinterpret_cast<const uint4*>(ha + sw128_off(HR, row, cb + c + 8)), f + 8);
                if (stage1) {
                    acc8(dq[c / 8], f);
                    acc8(dq[c / 8 + 1], f + 8);
                }
            }
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(f[j]);
            tmem_st16(tlane + kColAcc + cb + c, v);
        }
        tmem_st_wait();
        if (f_seg && tid < 128) {
            sseg[4 + row] = sid_me;
            if (row == 0) {
                sseg[3] = sid_prev;
                sseg[132] = sid_next;
            }
        }
        publish();
        tick(2);      // gather combine + publish

        // ---- P1: recompute h_a = relu(a_in . Wa^T + init + ba)
        if (!ha_given && warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t id = idesc_bf16(H, false, false);
            for (int ks = 0; ks < (ka >> 4); ++ks)
                mma_ss(tmem + kColAcc, desc_kmajor(ain_s, 128, ks), desc_kmajor(wa_s, H, ks), id, 1u);
            mma_commit(&mma_bar);
        }
        if (tile + (int)gridDim.x < n_tiles) {
            // next tile: its row ids now, and everything P0 will ask for into L2 (bulk tiles by the
            // copy engine, gathered rows one prefetch per 128-byte line), so that P0 waits on L2 only
            const int Rn = (tile + (int)gridDim.x) << 7;
            load_idx(tile + gridDim.x);
            prefetch_inputs(Rn);
        }
        if (!ha_given) wait_mma();
        tick(3);      // P1 MMA
        if (!ha_given) {
            uint32_t v[CH];
#pragma unroll
            for (int c = 0; c < CH; c += 16) tmem_ld16(tlane + kColAcc + cb + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
            tmem_ld_wait();
            if (norm) {      // P2 accumulates onto bb
#pragma unroll
                for (int c = 0; c < CH; c += 16) {
                    uint32_t b16[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const uint4 q = *reinterpret_cast<const uint4*>(s_bb + cb + c + j);
                        b16[j] = q.x; b16[j + 1] = q.y; b16[j + 2] = q.z; b16[j + 3] = q.w;
                    }
                    tmem_st16(tlane + kColAcc + cb + c, b16);
                }
            }
#pragma unroll
            for (int c = 0; c < CH; c += 8)
                *reinterpret_cast<uint4*>(ha + sw128_off(HR, row, cb + c)) = pack8_relu(reinterpret_cast<const float*>(&v[c]));
            tmem_st_wait();
            publish();
        }
        tick(4);      // E1

        // ---- P2 (NORM): m = h_a . Wb^T + bb ; delta_b = dRMSNorm(m) . du ; q = du * m/(rms+eps)
        if (norm) {
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint32_t id = idesc_bf16(H, false, false);
                for (int ks = 0; ks < (H >> 4); ++ks)
                    mma_ss(tmem + kColAcc, desc_kmajor(ha_s, HR, ks), desc_kmajor(wb_s, nb, ks), id, 1u);
                mma_commit(&mma_bar);
            }
            wait_mma();
            tick(5);  // P2 MMA
            // du for 8 columns of this thread's row: staged bf16 tiles, or fp32 rows read directly
            auto load_du = [&](int c0, float* du) {
                if (du_smem) {
                    unpack8(*reinterpret_cast<const uint4*>(qb + sw128_off(128, row, c0)), du);
                    if (f_gather) acc8(*reinterpret_cast<const uint4*>(db + sw128_off(128, row, c0)), du);
                } else if (MODE == 3) {      // staged in P0 (see there)
                    const uint8_t* base = (c0 >= 64 ? db : qb) + row * 256;
                    const int cc = (c0 & 63) >> 2;
                    const float4 t0 = *reinterpret_cast<const float4*>(base + ((cc ^ (row & 15)) << 4));
                    const float4 t1 = *reinterpret_cast<const float4*>(base + (((cc + 1) ^ (row & 15)) << 4));
                    du[0] = t0.x; du[1] = t0.y; du[2] = t0.z; du[3] = t0.w;
                    du[4] = t1.x; du[5] = t1.y; du[6] = t1.z; du[7] = t1.w;
                } else {
                    if (f_gyf32) {
                        const float4* gp_ = reinterpret_cast<const float4*>(p.gy_f32 + (size_t)crow * p.ld_gy + c0);
                        const float4 t0 = __ldg(gp_), t1 = __ldg(gp_ + 1);
                        du[0] = t0.x; du[1] = t0.y; du[2] = t0.z; du[3] = t0.w;
                        du[4] = t1.x; du[5] = t1.y; du[6] = t1.z; du[7] = t1.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) du[j] = 0.f;
                    }
                    if (f_gather) {
                        const int gi = p.gy_idx ? __ldg(p.gy_idx + crow) : crow;
                        const float4* ap = reinterpret_cast<const float4*>(p.gy_gather + (size_t)gi * H + c0);
                        const float4 t0 = __ldg(ap), t1 = __ldg(ap + 1);
                        du[0] += t0.x; du[1] += t0.y; du[2] += t0.z; du[3] += t0.w;
                        du[4] += t1.x; du[5] += t1.y; du[6] += t1.z; du[7] += t1.w;
                    }
                }
            };
            uint32_t v[CH];
#pragma unroll
            for (int c = 0; c < CH; c += 16) tmem_ld16(tlane + kColAcc + cb + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
            tmem_ld_wait();
            // du is read once and kept in registers for both passes (the direct fp32 path issues all
            // of its row loads here, back to back)
            float duv[CH];
#pragma unroll
            for (int c = 0; c < CH; c += 8) load_du(cb + c, duv + c);
            float ss = 0.f, dot = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float m = __uint_as_float(v[c]);
                ss = fmaf(m, m, ss);
                dot = fmaf(duv[c] * s_g[cb + c], m, dot);
            }
            s_red[(0 * NPART + part) * 128 + row] = ss;
            s_red[(1 * NPART + part) * 128 + row] = dot;
            __syncthreads();
            ss = 0.f;
            dot = 0.f;
#pragma unroll
            for (int q = 0; q < NPART; ++q) {       // fixed order: every part of the row gets the same sums
                ss += s_red[(0 * NPART + q) * 128 + row];
                dot += s_red[(1 * NPART + q) * 128 + row];
            }
            const float rms = sqrtf(ss * (1.f / H));
            const float s1 = 1.f / (rms + 1e-8f);
            // rows past the end get s = coef = 0, i.e. delta_b = q = 0 (their du / m are finite)
            const float s = valid ? s1 : 0.f;
            const float coef = (valid && rms > 0.f) ? dot * s1 * s1 / (rms * H) : 0.f;
#pragma unroll
            for (int c = 0; c < CH; c += 8) {
                float dm[8], q[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float m = __uint_as_float(v[c + j]);
                    dm[j] = fmaf(s_g[cb + c + j] * duv[c + j], s, -(coef * m));
                    q[j] = (duv[c + j] * m) * s;
                }
                *reinterpret_cast<uint4*>(db + sw128_off(128, row, cb + c)) = pack8(dm);
                *reinterpret_cast<uint4*>(qb + sw128_off(128, row, cb + c)) = pack8(q);
            }
            publish();
            tick(6);  // E2 (norm backward)
        }

        // ---- P3: dWb += delta_b^T h_a ; dbb += delta_b^T 1 ; dscale += q^T 1 ; acc = delta_b . Wb
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            // the dgrad chain first, on its own barrier: the epilogue that needs it starts while the
            // weight-gradient chains (which nobody waits for until their operands are rewritten) still run
            const uint32_t id_d = idesc_bf16(H, false, true);
            for (int ks = 0; ks < (nb >> 4); ++ks)
                mma_ss(tmem + kColAcc, desc_kmajor(db_s, 128, ks), desc_mnmajor(wb_s, nb, ks), id_d, ks > 0 ? 1u : 0u);
            mma_commit(&mma_bar);
            // [dWb | dbb] in one chain when the ones tile is "block 2" of ha (specialised instantiations)
            const uint32_t id_w = idesc_bf16(F ? H + 16 : H, true, true), id_1 = idesc_bf16(16, true, true);
            for (int ks = 0; ks < 8; ++ks)
                mma_ss(tmem + kColDWB, desc_mnmajor(db_s, 128, ks, 0, lbo_nb), desc_mnmajor(ha_s, HR, ks), id_w,
                       (ks > 0) ? 1u : acc_flag);
            if (!F)
                for (int ks = 0; ks < 8; ++ks)
                    mma_ss(tmem + kColDBB, desc_mnmajor(db_s, 128, ks, 0, lbo_nb), desc_mnmajor(ones_s, 128, ks), id_1,
                           (ks > 0) ? 1u : acc_flag);
            if (norm)
                for (int ks = 0; ks < 8; ++ks)
                    mma_ss(tmem + kColDSC, desc_mnmajor(qb_s, 128, ks, 0, lbo_h), desc_mnmajor(ones_s, 128, ks), id_1,
                           (ks > 0) ? 1u : acc_flag);
            mma_commit(&wg_bar);
        }
        wait_mma();
        tick(7);      // P3 dgrad MMAs
        // delta_a = acc * (h_a > 0), written in place over h_a once its reader (the dWb chain) has completed
        {
            uint32_t v[CH];
#pragma unroll
            for (int c = 0; c < CH; c += 16) tmem_ld16(tlane + kColAcc + cb + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
            tmem_ld_wait();
            uint4 o[CH / 8];
#pragma unroll
            for (int c = 0; c < CH; c += 8)
                o[c / 8] = mask8_pos(pack8(reinterpret_cast<const float*>(&v[c])),
                                     *reinterpret_cast<const uint4*>(ha + sw128_off(HR, row, cb + c)));
            wait_wg();
#pragma unroll
            for (int c = 0; c < CH; c += 8) *reinterpret_cast<uint4*>(ha + sw128_off(HR, row, cb + c)) = o[c / 8];
        }
        publish();
        tick(8);      // E3

        // ---- P4: dWa += delta_a^T a_in ; dba += delta_a^T 1 ; d_in = delta_a . Wa
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            // d_in first (its epilogue overlaps the weight-gradient chain); [dWa | dba] likewise in one chain:
            // the ones tile follows ain
            if (f_din) {
                const uint32_t id_d = idesc_bf16(ka, false, true);
                for (int ks = 0; ks < (H >> 4); ++ks)
                    mma_ss(tmem + kColAcc, desc_kmajor(ha_s, HR, ks), desc_mnmajor(wa_s, H, ks), id_d, ks > 0 ? 1u : 0u);
                mma_commit(&mma_bar);
            }
            const uint32_t id_w = idesc_bf16(F ? ka + 16 : ka, true, true), id_1 = idesc_bf16(16, true, true);
            for (int ks = 0; ks < 8; ++ks)
                mma_ss(tmem + kColDWA, desc_mnmajor(ha_s, HR, ks, 0, lbo_ha), desc_mnmajor(ain_s, 128, ks), id_w,
                       (ks > 0) ? 1u : acc_flag);
            if (!F)
                for (int ks = 0; ks < 8; ++ks)
                    mma_ss(tmem + kColDBA, desc_mnmajor(ha_s, HR, ks, 0, lbo_ha), desc_mnmajor(ones_s, 128, ks), id_1,
                           (ks > 0) ? 1u : acc_flag);
            mma_commit(&wg_bar);
        }
        // while the tensor core runs: delta_a tile -> global (row-major chunks), and its segment sum
        if (f_da && t_da) {
            if (warp == 0 && elect_one()) {
                for (int b = 0; b < (H + 63) >> 6; ++b) tma_store_2d(&maps.da_out, b * 64, R0, ha_s + b * kHaStride);
                tma_store_commit();
            }
        } else if (f_da) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = tid + j * NT;
                const int r = i / KC, ch = i % KC;
                if (R0 + r < p.rows)
                    *reinterpret_cast<uint4*>(p.delta_a_out + (size_t)(R0 + r) * H + ch * 8) =
                        *reinterpret_cast<const uint4*>(ha + sw128_off(HR, r, ch * 8));
            }
        }
        if (f_seg) {
            if (MODE == 2) tile_segment_sum_bf16_flat<H, NT>(ha, sseg, R0, tid, p.seg_bnd, p.seg_out_bf16, kHaStride);     // straight-line walk
            else tile_segment_sum<H, NT>(ha, sseg, R0, tid, p.seg_out, p.seg_bnd, p.seg_out_bf16, kHaStride);
        }
        tick(9);      // P4 issue + copy-out + segment walk
        if (f_din) wait_mma();
        tick(10);     // P4 d_in MMA wait
        if (f_din) {
            const bool via_smem = F ? kTOut[MODE] : (p.out_bf16 != nullptr && ka == H);     // bf16 tile output: transpose through shared memory
            uint8_t* ob = (t_out && norm) ? qb : db;                    // staging tile (free since P3)
            const int nsplit = (ka >= 16 * NPART) ? NPART : (ka >= 32 ? 2 : 1);   // column parts that take part
            const int kh = ka / nsplit;
            if (part < nsplit) {
                for (int c0 = part * kh; c0 < part * kh + kh; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tlane + kColAcc + c0, v);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                    if (via_smem) {
                        uint4 q0 = pack8(f), q1 = pack8(f + 8);
                        if (f_mask) {
                            q0 = mask8_pos(q0, *reinterpret_cast<const uint4*>(ain + sw128_off(128, row, c0)));
                            q1 = mask8_pos(q1, *reinterpret_cast<const uint4*>(ain + sw128_off(128, row, c0 + 8)));
                        }
                        if (t_out && t_res) {     // residual tile already in shared memory (bulk-loaded in P0)
                            q0 = add8_bf16(q0, *reinterpret_cast<const uint4*>(qb + sw128_off(128, row, c0)));
                            q1 = add8_bf16(q1, *reinterpret_cast<const uint4*>(qb + sw128_off(128, row, c0 + 8)));
                        }
                        *reinterpret_cast<uint4*>(ob + sw128_off(128, row, c0)) = q0;
                        *reinterpret_cast<uint4*>(ob + sw128_off(128, row, c0 + 8)) = q1;
                        continue;
                    }
                    if (f_mask) {
                        float av[16];
                        unpack8(*reinterpret_cast<const uint4*>(ain + sw128_off(128, row, c0)), av);
                        unpack8(*reinterpret_cast<const uint4*>(ain + sw128_off(128, row, c0 + 8)), av + 8);
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = av[j] > 0.f ? f[j] : 0.f;
                    }
                    if (valid) {
                        if (f_resid) {
                            float rv[16];
                            const gp_bf16* rp = p.out_resid + (size_t)grow * p.ld_out + c0;
                            unpack8(ldg16(rp), rv);
                            unpack8(ldg16(rp + 8), rv + 8);
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] += rv[j];
                        }
                        if (f_obf) {
                            uint4* d = reinterpret_cast<uint4*>(p.out_bf16 + (size_t)grow * p.ld_out + c0);
                            d[0] = pack8(f);
                            d[1] = pack8(f + 8);
                        } else {
                            float4* d = reinterpret_cast<float4*>(p.out_f32 + (size_t)grow * p.ld_out + c0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) d[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        }
                    }
                }
            }
            if (via_smem && t_out) {
                fence_async_smem();
                __syncthreads();
                if (warp == 0 && elect_one()) {
                    for (int b = 0; b < (ka + 63) >> 6; ++b)
                        tma_store_2d(&maps.out, b * 64, R0, smem_u32(ob) + b * 16384);
                    tma_store_commit();
                }
            } else if (via_smem) {
                uint4 rq[CPT];
                if (f_resid) {       // residual chunks requested together, before the barrier
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int i = tid + j * NT;
                        rq[j] = ldg16(p.out_resid + (size_t)min(R0 + i / KC, p.rows - 1) * p.ld_out + (i % KC) * 8);
                    }
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int i = tid + j * NT;
                    const int r = i / KC, ch = i % KC;
                    if (R0 + r < p.rows) {
                        uint4 q = *reinterpret_cast<const uint4*>(db + sw128_off(128, r, ch * 8));
                        if (f_resid) {
                            q = add8_bf16(q, rq[j]);
                        }
                        *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)(R0 + r) * p.ld_out + ch * 8) = q;
                    }
                }
            }
        }
        // bulk stores still reading a buffer that the next tile refills with per-thread copies must
        // finish their reads first; the output staging tile is refilled by the elected thread's own
        // bulk loads (after its wait in P0), so the newest store group may stay in flight then
        if ((t_da || t_out) && warp == 0 && elect_one()) {
            const bool ob_refilled_by_tma = norm ? t_gy : t_db;
            if (t_out && ob_refilled_by_tma) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        wait_wg();         // the weight-gradient chain of P4 has read ha / ain
        fence_async_smem();
        tc_fence_before();
        __syncthreads();   // buffers and ACC are free for the next tile
        tick(11);     // E4 + output
    }

    BWD_STAMP(17);        // ... tile loop done
    if ((t_da || t_out) && warp == 0 && elect_one()) tma_store_wait_all();
    __syncthreads();      // the staging tile below reuses buffers the last bulk stores were reading
    // ---- dump the weight-gradient accumulators of this CTA (lane r <-> output row r)
    tc_fence_after();
    const BwdLayout L = bwd_layout(H, ka, nb);
    float* P = p.partials + (size_t)blockIdx.x * L.stride;
    // the two weight-gradient matrices leave through a 64 KB staging tile (two adjacent dead tile buffers) as
    // coalesced rows
    uint8_t* const stg = tiles + (F ? 16384 : 0);
    tmem_rows_to_global<NT>(tlane, kColDWB, nb, H, P + L.off_dwb, H, stg, tid);
    tmem_rows_to_global<NT>(tlane, kColDWA, H, ka, P + L.off_dwa, ka, stg, tid);
    if (tid < 128) {
        uint32_t v8[8];
        tmem_ld8(tlane + kColDBB, v8);
        tmem_ld_wait();
        if (row < nb) P[L.off_dbb + row] = __uint_as_float(v8[0]);
        tmem_ld8(tlane + kColDBA, v8);
        tmem_ld_wait();
        if (row < H) P[L.off_dba + row] = __uint_as_float(v8[0]);
        if (norm) {
            tmem_ld8(tlane + kColDSC, v8);
            tmem_ld_wait();
            if (row < H) P[L.off_dsc + row] = __uint_as_float(v8[0]);
        } else if (row < H) {
            P[L.off_dsc + row] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    BWD_STAMP(18);        // ... partial blocks written
    if (tid < 32) tmem_dealloc(tmem, 512);
}

// Block = 32 outputs x 8 groups: group g adds the partial blocks p = g, g + 8, ... in ascending order, then the eight
// group sums are added in fixed order (bit-reproducible; an eighth of the dependent-load chain of a serial sum).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n_parts, int stride, int offset,
                                                              int rows, int cols, int ld_part, float* __restrict__ dst, int ld_dst,
                                                              int accumulate) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    const bool in = i < rows * cols;
    const int r = in ? i / cols : 0, c = in ? i - r * cols : 0;
    float acc = 0.f;
    if (in) {
        const float* src = partials + offset + (size_t)r * ld_part + c;
        for (int pi = ty; pi < n_parts; pi += 8) acc += src[(size_t)pi * stride];
    }
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty != 0 || !in) return;
    acc = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) acc += sh[g][tx];
    float* d = dst + (size_t)r * ld_dst + c;
    *d = accumulate ? (*d + acc) : acc;
}

struct ReduceSegs {
    gp_reduce_seg s[32];
    int n;
};
// Block = 32 lanes x 8 partial-groups: thread (tx, ty) adds the partial blocks p = ty, ty+8, ... in
// ascending order, then the 8 group sums are added in fixed order -> bit-reproducible.  VEC = 4:
// every lane owns four consecutive elements (16-byte loads; needs 4-element alignment of every
// offset / stride, checked on the host), VEC = 1 is the general path.
template <int VEC>
__global__ void __launch_bounds__(256) reduce_multi_kernel(const __grid_constant__ ReduceSegs segs) {
    __shared__ float sh[8][32 * VEC + 4];
    const gp_reduce_seg& sg = segs.s[blockIdx.y];
    const float* __restrict__ partials = sg.partials;
    const int n_parts = sg.n_parts, stride = sg.stride;
    const int total = sg.rows * sg.cols;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32 * VEC; base < total; base += gridDim.x * 32 * VEC) {
        const int i = base + tx * VEC;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        int r = 0, c = 0;
        if (i < total) {
            r = i / sg.cols;
            c = i - r * sg.cols;
            const float* src = partials + sg.offset + (size_t)r * sg.ld_part + c;
            for (int pi = ty; pi < n_parts; pi += 8) {
                if constexpr (VEC == 4) {
                    const float4 q = *reinterpret_cast<const float4*>(src + (size_t)pi * stride);
                    acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
                } else {
                    acc[0] += src[(size_t)pi * stride];
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) sh[ty][tx * VEC + v] = acc[v];
        __syncthreads();
        if (ty == 0 && i < total) {
            float* d = sg.dst + (size_t)r * sg.ld_dst + c;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float t = sh[0][tx * VEC + v];
#pragma unroll
                for (int k = 1; k < 8; ++k) t += sh[k][tx * VEC + v];
                d[v] = sg.accumulate ? (d[v] + t) : t;
            }
        }
        __syncthreads();
    }
}

template <int H>
int launch_bwd(const gp_mlp_bwd_args& a, int32_t* grid_out, cudaStream_t st) {
    size_t smem = 1024;
    smem += (size_t)((a.ka + 63) / 64) * H * 128 + (size_t)((H + 63) / 64) * a.nb * 128;
    BwdMaps maps;
    memset(&maps, 0, sizeof(maps));
    static const bool no_tma = getenv("GP_NO_TMA") != nullptr;
    if (!no_tma) {
        uint32_t use = 0;
        if (a.a_bf16 && gp::tma_map_2d(&maps.ain, a.a_bf16, a.rows, a.ka, a.lda)) use |= kMapAin;
        if (a.mode == 0 && gp::tma_map_2d(&maps.db, a.delta_b, a.rows, a.nb, a.ld_db)) use |= kMapDb;
        if (a.mode == 1 && a.gy_bf16 && gp::tma_map_2d(&maps.gy, a.gy_bf16, a.rows, H, a.ld_gy)) use |= kMapGy;
        if (a.delta_a_out && gp::tma_map_2d(&maps.da_out, a.delta_a_out, a.rows, H, H)) use |= kMapDaOut;
        if (a.ha_saved && gp::tma_map_2d(&maps.ha, a.ha_saved, a.rows, H, H)) use |= kMapHa;
        if (a.need_din && a.out_bf16 && a.ka == H && gp::tma_map_2d(&maps.out, a.out_bf16, a.rows, a.ka, a.ld_out)) {
            // the residual rides through the spare tile buffer; GIVEN mode only (NORM keeps du / q there)
            if (!a.out_resid)
                use |= kMapOut;
            else if (a.mode == 0 && gp::tma_map_2d(&maps.resid, a.out_resid, a.rows, a.ka, a.ld_out))
                use |= kMapOut | kMapResid;
        }
        maps.use = use;
    }
    smem += (size_t)((a.mode == 1 || (maps.use & kMapResid)) ? 4 : 3) * kBufBytes + 128 * 128 + 3 * 128 * 4 + 2 * BwdCfg<H>::NPART * 128 * 4 + 144 * 4;
    GP_REQUIRE((int)smem <= gp::max_smem_optin(), "gp_mlp_bwd_stage: needs %zu B of shared memory (> %d)", smem,
               gp::max_smem_optin());
    // the stages of the processor's edge and node MLPs get their own instantiations (H = 128 only)
    int fast = 0;
    if (H == 128 && a.ka == H && a.nb == H && !a.ha_saved && a.need_din) {
        const bool plain = !a.out_resid && !a.delta_a_out && !a.seg_id;
        if (a.mode == 1 && a.a_bf16 && a.out_bf16 && !a.init && a.mask_by_ain && plain) {
            if (a.gy_bf16 && !a.gy_gather && (!a.gy_gather_bf16 || a.gy_idx) && maps.use == (kMapAin | kMapGy | kMapOut))
                fast = 1;                                                                                             // edge B (+ edge encoder B)
            if (a.gy_f32 && !a.gy_gather && !a.gy_gather_bf16 && maps.use == (kMapAin | kMapOut)) fast = 3;          // node B
        }
        if (a.mode == 0 && a.init && !a.mask_by_ain && a.delta_a_out) {
            if (a.a_bf16 && a.out_bf16 && a.two_inits && a.idx0 && a.idx1 && a.out_resid && a.seg_id && a.seg_out_bf16 &&
                maps.use == (kMapAin | kMapDb | kMapResid | kMapDaOut | kMapOut))
                fast = 2;                                                                                             // edge A
            if (a.a_bf16 && a.out_bf16 && !a.two_inits && !a.idx0 && !a.out_resid && !a.seg_id &&
                maps.use == (kMapAin | kMapDb | kMapDaOut | kMapOut))
                fast = 4;                                                                                             // node A
        }
    }
    static const bool force_general = getenv("GP_BWD_GENERAL") != nullptr;      // A/B testing of the instantiations
    if (force_general) fast = 0;
    auto set_smem = [&](int bytes) -> cudaError_t {
        switch (fast) {
            case 1: return cudaFuncSetAttribute(mlp_bwd_kernel<H, (H == 128 ? 1 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            case 2: return cudaFuncSetAttribute(mlp_bwd_kernel<H, (H == 128 ? 2 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            case 3: return cudaFuncSetAttribute(mlp_bwd_kernel<H, (H == 128 ? 3 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            case 4: return cudaFuncSetAttribute(mlp_bwd_kernel<H, (H == 128 ? 4 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            default: return cudaFuncSetAttribute(mlp_bwd_kernel<H, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        }
    };
    static int smem_set[5] = {0, 0, 0, 0, 0};      // raised once per instantiation (and never inside a stream capture twice)
    if ((int)smem > smem_set[fast]) {
        GP_CHECK_CUDA(set_smem((int)smem));
        smem_set[fast] = (int)smem;
    }
    const int n_tiles = (a.rows + 127) / 128;
    // as many CTAs as keep the number of tile rounds minimal, not more: every CTA dumps (and the reduction re-reads)
    // a full block of partial weight gradients (503 node tiles: 126 CTAs x 4 tiles instead of 148 x 3.4)
    int grid = n_tiles < gp::sm_count() ? n_tiles : gp::sm_count();
    if (grid > 0) {
        const int rounds = (n_tiles + grid - 1) / grid;
        grid = (n_tiles + rounds - 1) / rounds;
    }
    const dim3 g3(grid), b3(BwdCfg<H>::NT);
    cudaError_t le;
    if (fast == 1)
        le = gp::launch_kernel(mlp_bwd_kernel<H, (H == 128 ? 1 : 0)>, g3, b3, smem, st, a, maps);
    else if (fast == 2)
        le = gp::launch_kernel(mlp_bwd_kernel<H, (H == 128 ? 2 : 0)>, g3, b3, smem, st, a, maps);
    else if (fast == 3)
        le = gp::launch_kernel(mlp_bwd_kernel<H, (H == 128 ? 3 : 0)>, g3, b3, smem, st, a, maps);
    else if (fast == 4)
        le = gp::launch_kernel(mlp_bwd_kernel<H, (H == 128 ? 4 : 0)>, g3, b3, smem, st, a, maps);
    else
        le = gp::launch_kernel(mlp_bwd_kernel<H, 0>, g3, b3, smem, st, a, maps);
    GP_CHECK_CUDA(le);
    if (grid_out) *grid_out = grid;
    return 0;
}
}  // namespace

extern "C" int gp_mlp_bwd_layout(int hidden, int ka, int nb, int32_t* out6) {
    GP_REQUIRE(out6 != nullptr, "gp_mlp_bwd_layout: null output");
    const BwdLayout L = bwd_layout(hidden, ka, nb);
    out6[0] = L.off_dwb; out6[1] = L.off_dwa; out6[2] = L.off_dbb; out6[3] = L.off_dba; out6[4] = L.off_dsc;
    out6[5] = L.stride;
    return 0;
}

extern "C" int gp_mlp_bwd_stage(const gp_mlp_bwd_args* args, int hidden, int32_t* grid_out, void* stream) {
    GP_REQUIRE(args != nullptr, "gp_mlp_bwd_stage: null args");
    const gp_mlp_bwd_args& a = *args;
    GP_REQUIRE(a.rows >= 0, "gp_mlp_bwd_stage: rows must not be negative");
    if (a.rows == 0) {           // an edge-less graph: no rows, no partial blocks (the reductions then write zeros)
        if (grid_out) *grid_out = 0;
        return 0;
    }
    GP_REQUIRE(a.ka > 0 && a.ka % 16 == 0 && a.ka <= 128, "gp_mlp_bwd_stage: bad ka=%d", a.ka);
    GP_REQUIRE(a.nb > 0 && a.nb % 16 == 0 && a.nb <= 128, "gp_mlp_bwd_stage: bad nb=%d", a.nb);
    GP_REQUIRE((a.a_bf16 != nullptr) != (a.a_f32 != nullptr), "gp_mlp_bwd_stage: exactly one of a_bf16 / a_f32");
    GP_REQUIRE(a.wa && a.wb && a.partials, "gp_mlp_bwd_stage: null weights or partials");
    if (a.mode == 1) {
        GP_REQUIRE(a.nb == hidden, "gp_mlp_bwd_stage: NORM mode needs nb == hidden");
        GP_REQUIRE(!(a.gy_bf16 && a.gy_f32) && (a.gy_bf16 || a.gy_f32 || a.gy_gather),
                   "gp_mlp_bwd_stage: NORM mode needs gy_bf16 or gy_f32 (not both), or at least gy_gather");
        GP_REQUIRE(!a.gy_gather_bf16 || (a.gy_bf16 && !a.gy_gather && a.gy_idx),
                   "gp_mlp_bwd_stage: gy_gather_bf16 needs gy_bf16 and gy_idx, and excludes gy_gather");
    } else {
        GP_REQUIRE(a.mode == 0 && a.delta_b != nullptr && a.ld_db % 8 == 0, "gp_mlp_bwd_stage: GIVEN mode needs delta_b");
    }
    if (a.need_din) GP_REQUIRE((a.out_bf16 != nullptr) != (a.out_f32 != nullptr), "gp_mlp_bwd_stage: need one output");
    if (a.seg_id) GP_REQUIRE(((a.seg_out != nullptr) != (a.seg_out_bf16 != nullptr)) && a.seg_bnd,
                             "gp_mlp_bwd_stage: segment sum needs seg_bnd and exactly one of seg_out / seg_out_bf16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (hidden) {
        case 128: return launch_bwd<128>(a, grid_out, st);
        case 64: return launch_bwd<64>(a, grid_out, st);
        case 32: return launch_bwd<32>(a, grid_out, st);
        default: gp::set_error("gp_mlp_bwd_stage: unsupported hidden size %d (32, 64, 128)", hidden); return -1;
    }
}

extern "C" int gp_reduce_partials(const float* partials, int32_t n_parts, int32_t stride, int32_t offset, int32_t rows,
                                  int32_t cols, int32_t ld_part, float* dst, int32_t ld_dst, int32_t accumulate,
                                  void* stream) {
    const int total = rows * cols;
    if (total <= 0) return 0;
    reduce_partials_kernel<<<(total + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        partials, n_parts, stride, offset, rows, cols, ld_part, dst, ld_dst, accumulate);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_reduce_partials_multi(const float* partials, int32_t n_parts, int32_t stride,
                                        const gp_reduce_seg* segs_host, int32_t n_segs, void* stream) {
    GP_REQUIRE(n_segs >= 0 && n_segs <= 32, "gp_reduce_partials_multi: at most 32 segments");
    if (n_segs == 0) return 0;
    ReduceSegs segs;
    segs.n = n_segs;
    int max_total = 0;
    bool vec4 = true;
    for (int i = 0; i < n_segs; ++i) {
        gp_reduce_seg& g = segs.s[i];
        g = segs_host[i];
        if (!g.partials) {
            g.partials = partials;
            g.n_parts = n_parts;
            g.stride = stride;
        }
        // n_parts == 0 is legal (a stage over zero rows wrote no partial block): the segment's sum is zero
        GP_REQUIRE(g.partials != nullptr && g.n_parts >= 0, "gp_reduce_partials_multi: segment %d has no partials", i);
        const int t = g.rows * g.cols;
        if (t > max_total) max_total = t;
        vec4 = vec4 && g.stride % 4 == 0 && (reinterpret_cast<uintptr_t>(g.partials) & 15u) == 0 && g.cols % 4 == 0 &&
               g.offset % 4 == 0 && g.ld_part % 4 == 0 && g.ld_dst % 4 == 0 && (reinterpret_cast<uintptr_t>(g.dst) & 15u) == 0;
    }
    const int per_block = vec4 ? 128 : 32;
    int bx = (max_total + per_block - 1) / per_block;
    if (bx > 512) bx = 512;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n_segs);
    if (vec4)
        reduce_multi_kernel<4><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(segs);
    else
        reduce_multi_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(segs);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
