// mlp_fwd.cu -- fused row-tile MLP forward on tcgen05 (see include/gp_b200.h, gp_mlp_fwd).
//
// Persistent kernel, one CTA per SM.  The packed bf16 weights of all layers stay resident in
// shared memory (SW128 row tiles) for the whole launch.  A CTA runs NG independent tile slots of
// 256 threads; each slot owns one 128-row activation buffer, one 128-column fp32 accumulator in
// TMEM and one mbarrier, and walks its own tiles.  Inside a slot the flow is bulk-synchronous
// (stage -> MMA -> epilogue per layer); the slots run out of phase, so one slot's tcgen05.mma
// and global-memory latency overlap the other slot's TMEM epilogue (16 resident warps per SM).
//
// Thread (row = t & 127, half = t >> 7) of a slot owns row `row` of the tile (== TMEM lane) and
// one half of its columns: bias, ReLU and RMSNorm are thread-local apart from one float
// exchanged between the halves.  The bias is stored into the accumulator before the MMAs, so the
// epilogues only convert (ReLU fused into the bf16 conversion).  Everything that goes to or comes
// from global memory as a tile (layer-0 operand, gathered rows, saved activation, output + residual)
// moves by TMA or in row-major 16-byte chunks, one chunk per lane, so a warp touches 4 cache lines
// per instruction instead of 32; the tile is transposed between the two mappings through the
// swizzled shared-memory buffer.  The saved activation leaves by one TMA store while the next
// layer's MMA is in flight; the next tile's row ids and gathered rows are prefetched meanwhile.
// The receiver-sorted segment sum re-partitions through shared memory (column pairs x sub-tiles
// of H/4 rows) and uses no atomics.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tile_util.cuh"

struct gp_mlp_fwd_args;
namespace gp {
int try_linear_fwd(const gp_mlp_fwd_args& a, int hidden, cudaStream_t st);
int try_edge_fwd2(const gp_mlp_fwd_args& a, cudaStream_t st);
}

namespace {
using namespace gp;

constexpr int kBiasStride = 384;
constexpr int kSlotThreads = 256;

// TMA descriptors of the tile-shaped global tensors (valid ones flagged in `use`): the layer-0
// operand is bulk-loaded and the saved activation bulk-stored by one elected thread per slot.
struct FwdMaps {
    CUtensorMap a, h[3];     // h[i]: saved output of layer i (save_h1, save_h2, save_h3)
    uint32_t use;
};
enum : uint32_t { kMapA = 1, kMapH = 2 /* << layer */ };

__device__ __forceinline__ void slot_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// kMode 1 / 2 are the instantiations for the processor's edge MLP (4 H-wide layers, two gathered
// pre-activation sources, RMSNorm, bf16 residual output, segment sum) and node MLP (bf16 aggregate
// as layer-0 operand, the node's own pre-activation row, RMSNorm, bf16 residual output): every
// optional path is resolved at compile time, which cuts the code the two out-of-phase slots stream
// through the instruction cache (a 10% larger kernel measured 13% slower).  kMode 0 is general.
template <int H, int NG, int kMode>
__global__ void __launch_bounds__(kSlotThreads * NG, 1) mlp_fwd_kernel(const gp_mlp_fwd_args p,
                                                                       const __grid_constant__ FwdMaps maps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t mma_bar[NG], tma_bar[NG];
    const bool t_a = kMode == 0 && (maps.use & kMapA);
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int g = tid / kSlotThreads;        // tile slot
    const int t = tid - g * kSlotThreads;    // thread in slot
    // The same slot / warp numbers taken from lane 0, which the compiler can prove warp-uniform; only
    // the single-thread issue blocks (tcgen05.mma, TMA) use them, so that their descriptors live in
    // uniform registers.  (Using them for the per-thread addressing as well measured 7% slower.)
    const int warp_u = warp_uniform(tid >> 5);
    const int g_u = warp_u / (kSlotThreads / 32);
    const int wis = warp_u - g_u * (kSlotThreads / 32);    // warp in slot
    const int row = t & 127;                 // row in tile == TMEM lane
    const int half = t >> 7;                 // column half
    constexpr bool kEdge = kMode == 1, kNode = kMode == 2, E = kEdge || kNode;     // E: a 4-layer H-wide processor MLP
    const int L = E ? 4 : p.n_layers;
    const bool f_norm = E || p.norm_scale != nullptr, f_resid = E || p.resid != nullptr, f_ybf = E || p.y_bf16 != nullptr;
    const bool f_seg = kEdge || (!E && p.seg_id != nullptr), f_abf = E || p.a_bf16 != nullptr;
    const bool f_idx0 = kEdge || (!E && p.idx0 != nullptr);
    const int ka = E ? H : p.ka;
    constexpr int CH = H / 2;                // columns per thread in H-wide layers
    constexpr int KC = H / 8;                // 16-byte chunks per H-wide row
    constexpr int CPT = 128 * KC / kSlotThreads;   // chunks per thread in a row-major tile copy
    const int cb = half * CH;

    // ---- carve shared memory (all offsets uniform across the CTA)
    uint32_t w_off[4];
    uint32_t off = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        w_off[l] = off;
        if (l < L) off += ((p.k[l] + 63) >> 6) * p.n[l] * 128;
    }
    const uint32_t buf_u = smem_u32(smem + off) + g_u * kBufBytes;      // slot buffer, uniform copy of its address
    uint8_t* buf = smem + off + g * kBufBytes;
    off += NG * kBufBytes;
    float* sbias = reinterpret_cast<float*>(smem + off);
    off += 4 * kBiasStride * 4;
    float* sscale = reinterpret_cast<float*>(smem + off);
    off += 128 * 4;
    float* sred = reinterpret_cast<float*>(smem + off) + g * 256;   // [2 halves][128 rows]
    off += NG * 256 * 4;
    int* sseg = reinterpret_cast<int*>(smem + off) + g * 144;       // 16-byte aligned, ids at [4..131]

    // ---- one-time staging: weights, biases, barriers, TMEM
    for (int l = 0; l < L; ++l) {
        stage_weight(smem + w_off[l], p.w[l], p.n[l], p.k[l]);
        for (int i = tid; i < kBiasStride; i += blockDim.x)
            sbias[l * kBiasStride + i] = (p.bias[l] && i < p.n[l]) ? p.bias[l][i] : 0.f;
    }
    cp_async_commit();
    if (f_norm)
        for (int i = tid; i < H; i += blockDim.x) sscale[i] = p.norm_scale[i];
    if (tid == 0) {
        for (int i = 0; i < NG; ++i) {
            mbar_init(&mma_bar[i], 1);
            mbar_init(&tma_bar[i], 1);
        }
        fence_mbar_init();
    }
    constexpr uint32_t kTmemCols = NG * 128 < 32 ? 32 : NG * 128;
    if (tid < 32) tmem_alloc(&tmem_slot, kTmemCols);
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (kNode) {
        // node update: the pre-activation rows (P) and the residual rows (x) of this slot's first tile are older than the
        // previous kernel: pull them into L2 while it finishes (every CTA runs only ~4 tiles; the first used to start cold)
        const int tile0_ = blockIdx.x * NG + g;
        if (tile0_ < ((p.rows + 127) >> 7) && t < 256) {
            const int r = min((tile0_ << 7) + (t >> 1), p.rows - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.init + (size_t)r * p.ld_init + p.init_off0 + (t & 1) * 64));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + (size_t)r * p.ld_out + (t & 1) * 64));
        }
    }
    // everything above is independent of the previous kernel in the stream (parameters only); from here on
    // its outputs are read, and the next kernel may start its own prologue
    pdl_wait();
    pdl_launch_dependents();

    const uint32_t tmem_base = tmem_slot;
    const uint32_t tacc_mma = tmem_base + g_u * 128;                             // D operand (lane 0)
    const uint32_t tacc = tmem_addr(tmem_base, (row >> 5) * 32, g * 128);         // this warp's lanes
    const uint32_t buf_s = smem_u32(buf);
    uint32_t phase = 0, tphase = 0;
    const bool has_init = E || p.init != nullptr;
    const int n_tiles = (p.rows + 127) >> 7;
    const int tile_stride = gridDim.x * NG;

#ifdef GP_MLP_PROF
    const bool prof = p.prof != nullptr && t == 0;      // phase timing (scratch/phase*.py): lib/libgp_b200_prof.so only
#else
    constexpr bool prof = false;                        // the product build carries no profiling code
#endif
    long long tk = 0;
    auto tick = [&](int slot) {   // accumulate cycles since the previous tick into counter `slot`
        if (prof) {
            const long long now = clock64();
            atomicAdd(p.prof + slot, (unsigned long long)(now - tk));
            tk = now;
        }
    };
    // the MMA issuer adds (byte offset >> 4) to descriptor templates built once
    const uint64_t adesc0 = desc_kmajor(buf_u, 128, 0);

    // gather indices of a tile (prefetched one tile ahead): i0n = this thread's row in the directly
    // loaded source; ridx[] = rows of the chunks this thread copies for the staged source
    const bool stage1 = kEdge || (!E && p.two_inits != 0);           // which source goes through buf
    const int32_t* sidx = stage1 ? p.idx1 : p.idx0;
    const int soff = stage1 ? p.init_off1 : p.init_off0;
    int i0n = 0, ridx[CPT];
    auto load_idx = [&](int tile_) {
        const int r = min((tile_ << 7) + row, p.rows - 1);
        i0n = f_idx0 ? __ldg(p.idx0 + r) : r;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int gr = min((tile_ << 7) + (t + j * kSlotThreads) / KC, p.rows - 1);
            ridx[j] = sidx ? __ldg(sidx + gr) : gr;
        }
    };
    // The accumulator always starts from the bias, so the epilogues never add it: before the MMAs of
    // (layer l, column chunk nc) are issued, bias[l][nc + c] is stored to column c of every lane.
    // A thread only overwrites columns it has itself just read: its CH columns of an H-wide
    // accumulator (`wide` = false), or its 64 of a full 128-column chunk (`wide` = true).
    auto prestore_bias = [&](int l, int nc, bool wide) {
        const int c0 = wide ? half * 64 : cb;
        const float* b = sbias + l * kBiasStride + nc + c0;
        const uint32_t ta = tmem_addr(tmem_base, (row >> 5) * 32, g * 128) + c0;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
            if (c < CH || wide) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const uint4 q = *reinterpret_cast<const uint4*>(b + c + j);
                    v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
                }
                tmem_st16(ta + c, v);
            }
        }
    };
    const int tile0 = blockIdx.x * NG + g;
    if (has_init && tile0 < n_tiles) load_idx(tile0);

    for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
        if (prof) tk = clock64();
        const int R0 = tile << 7;
        const int grow = R0 + row;
        const bool valid = grow < p.rows;

        // the residual rows are read at the very end of the tile (output pass): pull them into L2 now.  In the node update
        // they are x, which this kernel has not touched yet (the layer-0 operand is agg): read cold, the output pass took
        // 5 k of the tile's 22 k cycles
        if (f_resid && f_ybf) {
            constexpr int kLines = 128 * (H * 2 / 128 > 0 ? H * 2 / 128 : 1);
            for (int i = t; i < kLines; i += kSlotThreads) {
                const int r = i / (kLines / 128), part_ = i % (kLines / 128);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + (size_t)min(R0 + r, p.rows - 1) * p.ld_out + part_ * 64));
            }
        }
        // (a) segment ids of the tile (+ one row of context on each side): requested now, stored
        //     to shared memory later so their latency is not exposed
        int sid_me = -1, sid_prev = -1, sid_next = -1;
        if (f_seg && t < 128) {
            if (valid) sid_me = __ldg(p.seg_id + grow);
            if (row == 0) {
                if (R0 > 0) sid_prev = __ldg(p.seg_id + R0 - 1);
                if (R0 + 128 < p.rows) sid_next = __ldg(p.seg_id + R0 + 128);
            }
        }
        // (b) accumulator pre-load = fp32 sum of the gathered pre-activation rows.  The randomly
        //     indexed source (senders; or the node's own row) is gathered into buf with 16-byte
        //     chunks in row-major order -- 8 lanes per cache line, so the LSU sees ~1/8 of the
        //     wavefronts a row-per-thread load would cost -- and read back row-wise; the
        //     receiver-indexed source is sorted, its row-per-thread loads already coalesce.
        if (has_init) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = t + j * kSlotThreads;
                const int r = i / KC, ch = i % KC;
                cp_async16(buf_s + sw128_off(128, r, ch * 8), p.init + (size_t)ridx[j] * p.ld_init + soff + ch * 8);
            }
            cp_async_commit();
            // this thread's part of the directly loaded (receiver-indexed) row: requested before the wait
            uint4 dq[CH / 8];
            if (stage1) {
                const gp_bf16* dp = p.init + (size_t)i0n * p.ld_init + p.init_off0 + cb;
#pragma unroll
                for (int i = 0; i < CH / 8; ++i) dq[i] = ldg16(dp + i * 8);
            }
            // L2 prefetch of the layer-0 operand tile, which is staged right after
            if (f_abf && t < 128) {
                const gp_bf16* rp = p.a_bf16 + (size_t)min(grow, p.rows - 1) * p.lda;
                for (int c = 0; c < ka; c += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c));
            }
            cp_async_wait<0>();
            slot_sync(g);
#pragma unroll
            for (int c = 0; c < CH; c += 16) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(f + j) = *reinterpret_cast<const float4*>(sbias + cb + c + j);
                acc8(*reinterpret_cast<const uint4*>(buf + sw128_off(128, row, cb + c)), f);
                acc8(*reinterpret_cast<const uint4*>(buf + sw128_off(128, row, cb + c + 8)), f + 8);
                if (stage1) {
                    acc8(dq[c / 8], f);
                    acc8(dq[c / 8 + 1], f + 8);
                }
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(f[j]);
                tmem_st16(tacc + cb + c, v);
            }
            tmem_st_wait();
            fence_async_smem();
            slot_sync(g);            // buf is free for the layer-0 operand
        } else {
            prestore_bias(0, 0, true);
        }
        if (f_seg && t < 128) {
            sseg[4 + row] = sid_me;
            if (row == 0) {
                sseg[3] = sid_prev;
                sseg[132] = sid_next;
            }
        }
        // (c) streamed layer-0 operand -> buf (rows past the end replicate the last row; their
        //     results are never stored); bulk-loaded when a descriptor exists (rows past the end = 0)
        if (t_a) {
            if (wis == 0 && elect_one()) {
                const int nblk = (ka + 63) >> 6;
                mbar_arrive_expect_tx(&tma_bar[g_u], nblk * 16384u);
                for (int b = 0; b < nblk; ++b) tma_load_2d(buf_u + b * 16384, &maps.a, b * 64, R0, &tma_bar[g_u]);
            }
        } else {
            stage_rows(buf, p.a_bf16, E ? nullptr : p.a_f32, ka, p.lda, R0, p.rows, t, kSlotThreads);
            cp_async_commit();
        }
        tick(0);                 // issue of loads + gathers + TMEM pre-load (this thread)
        tmem_st_wait();
        cp_async_wait<0>();
        if (t_a) {
            mbar_wait(&tma_bar[g], tphase);
            tphase ^= 1;
        }
        fence_async_smem();
        tc_fence_before();
        slot_sync(g);
        tick(1);                 // waiting for the tile's loads / the slowest thread

        // (d) layers
        for (int l = 0; l < L; ++l) {
            const int K = E ? H : p.k[l];
            const int Nl = E ? H : p.n[l];
            const bool last = (l == L - 1);
            for (int nc = 0; nc < Nl; nc += 128) {
                const int ncols = E ? H : min(128, Nl - nc);
                if (wis == 0 && elect_one()) {
                    tc_fence_after();
                    const uint32_t idesc = idesc_bf16(ncols, false, false);
                    const uint64_t bdesc0 = desc_kmajor(smem_u32(smem + w_off[l]) + nc * 128, Nl, 0);
                    const uint32_t bblk = (uint32_t)(Nl * 128) >> 4;     // next 64-column block of W, in 16-B units
                    const int nks = K >> 4;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t ko = (ks & 3) * 2;                // 32 bytes per k-step inside a block
                        mma_ss(tacc_mma, adesc0 + (uint64_t)((ks >> 2) * 1024u + ko), bdesc0 + (uint64_t)((ks >> 2) * bblk + ko),
                               idesc, 1u);
                    }
                    mma_commit(&mma_bar[g_u]);
                }
                tick(9 + l);     // MMA issue (elected thread of the slot), per layer
                // while the first MMA runs: indices of this slot's next tile, and its gathered rows
                // into L2, so the next tile's pre-load does not start with two dependent misses
                if (l == 0 && nc == 0 && has_init && tile + tile_stride < n_tiles) load_idx(tile + tile_stride);
                if (l == 1 && nc == 0 && has_init && tile + tile_stride < n_tiles) {
#pragma unroll
                    for (int j = 0; j < CPT; ++j)
                        if (((t + j * kSlotThreads) & 7) == 0)     // one prefetch per 128-byte line of the staged rows
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.init + (size_t)ridx[j] * p.ld_init + soff +
                                                                         ((t + j * kSlotThreads) % KC) * 8));
                }
                // while the MMA of layer l runs: copy the saved activation (output of layer l-1, still
                // intact in buf as this MMA's A operand) to global memory -- one bulk store, or
                // row-major 16-byte chunks
                gp_bf16* const sv = (nc != 0 || l == 0) ? nullptr : (l == 1 ? p.save_h1 : (l == 2 ? p.save_h2 : p.save_h3));
                if (sv) {
                    if (maps.use & (kMapH << (l - 1))) {
                        if (wis == 0 && elect_one()) {
                            // (constant indices keep the descriptors in parameter space)
                            const void* mp = l == 1 ? &maps.h[0] : (l == 2 ? &maps.h[1] : &maps.h[2]);
                            for (int b = 0; b < (H + 63) >> 6; ++b) tma_store_2d(mp, b * 64, R0, buf_u + b * 16384);
                            tma_store_commit();
                            tma_store_wait_read<0>();     // buf has been read: the epilogue may overwrite it
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) {
                            const int i = t + j * kSlotThreads;
                            const int r = i / KC, ch = i % KC;
                            if (R0 + r < p.rows)
                                *reinterpret_cast<uint4*>(sv + (size_t)(R0 + r) * H + ch * 8) =
                                    *reinterpret_cast<const uint4*>(buf + sw128_off(128, r, ch * 8));
                        }
                    }
                    slot_sync(g);    // the tile is out (or read) before any thread overwrites buf in the epilogue
                }
                mbar_wait(&mma_bar[g], phase);
                phase ^= 1;
                tc_fence_after();
                tick(2);         // MMA issue + wait (+ overlapped copy-out)

                if (!last) {
                    // hidden layer: relu(acc + b) -> bf16 -> next A operand (in place over buf)
                    uint32_t v[CH];
#pragma unroll
                    for (int c = 0; c < CH; c += 16) tmem_ld16(tacc + cb + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
                    tmem_ld_wait();
                    prestore_bias(l + 1, 0, false);     // hidden layers are H wide: one chunk, and so is the next layer
#pragma unroll
                    for (int c = 0; c < CH; c += 8)
                        *reinterpret_cast<uint4*>(buf + sw128_off(128, row, cb + c)) =
                            pack8_relu(reinterpret_cast<const float*>(&v[c]));
                    fence_async_smem();
                    tmem_st_wait();
                    tick(3);
                } else if (f_norm) {
                    // RMSNorm (layers.py:104-129): u = scale * m / (||m||/sqrt(H) + 1e-8), rounded to
                    // bf16 once; that value feeds the segment sum and the residual alike.
                    uint32_t v[CH];
#pragma unroll
                    for (int c = 0; c < CH; c += 16) tmem_ld16(tacc + cb + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
                    tmem_ld_wait();
                    float ss = 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const float m = __uint_as_float(v[c]);
                        ss = fmaf(m, m, ss);
                    }
                    sred[half * 128 + row] = ss;
                    slot_sync(g);
                    ss = sred[row] + sred[128 + row];
                    const float rinv = 1.f / (sqrtf(ss * (1.f / H)) + 1e-8f);
#pragma unroll
                    for (int c = 0; c < CH; c += 8) {
                        float u[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) u[j] = sscale[cb + c + j] * (__uint_as_float(v[c + j]) * rinv);
                        *reinterpret_cast<uint4*>(buf + sw128_off(128, row, cb + c)) = pack8(u);
                    }
                    tick(4);
                } else {
                    // plain last layer (decoder / projection): acc + b, first n_valid columns
                    const int chh = ncols >= 32 ? ncols / 2 : ncols;      // half 1 idles on narrow outputs
                    if (ncols >= 32 || half == 0) {
                        for (int c0 = half * chh; c0 < half * chh + chh; c0 += 16) {
                            uint32_t v[16];
                            tmem_ld16(tacc + c0, v);
                            tmem_ld_wait();
                            float u[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) u[j] = __uint_as_float(v[j]);
                            if (valid) {
                                const int cg = nc + c0;
                                if (p.y_bf16 && cg + 16 <= p.n_valid) {
                                    uint4* d = reinterpret_cast<uint4*>(p.y_bf16 + (size_t)grow * p.ld_out + cg);
                                    d[0] = pack8(u);
                                    d[1] = pack8(u + 8);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (cg + j < p.n_valid) {
                                            if (p.y_bf16)
                                                reinterpret_cast<__nv_bfloat16*>(p.y_bf16)[(size_t)grow * p.ld_out + cg + j] =
                                                    __float2bfloat16(u[j]);
                                            else
                                                p.y_f32[(size_t)grow * p.ld_out + cg + j] = u[j];
                                        }
                                }
                            }
                        }
                    }
                    if (nc + 128 < Nl) {      // the next column chunk of this layer reuses the accumulator
                        prestore_bias(l, nc + 128, true);
                        tmem_st_wait();
                    }
                    tick(4);
                }
                tc_fence_before();
                slot_sync(g);
                tick(5);         // waiting for the slowest thread of the slot
            }
        }

        if (f_norm) {
            // (e) receiver-sorted segment sum of bf16(u) (fp32 accumulate, fixed order, no atomics)
            if (f_seg) tile_segment_sum<H, kSlotThreads>(buf, sseg, R0, t, kEdge ? nullptr : p.seg_out, p.seg_bnd, p.seg_out_bf16);
            tick(6);
            // (f) output: y = resid + bf16(u), row-major chunks; the residual chunks (L2 hits: the tile
            //     was this kernel's layer-0 operand) are requested together
            uint4 rq[CPT];
            if (f_resid) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int i = t + j * kSlotThreads;
                    const int r = i / KC, ch = i % KC;
                    rq[j] = ldg16(p.resid + (size_t)min(R0 + r, p.rows - 1) * p.ld_out + ch * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int i = t + j * kSlotThreads;
                const int r = i / KC, ch = i % KC;
                if (R0 + r < p.rows) {
                    const uint4 uq = *reinterpret_cast<const uint4*>(buf + sw128_off(128, r, ch * 8));
                    if (f_ybf) {
                        // bf16(u) + bf16 residual, exact sum rounded once to bf16
                        *reinterpret_cast<uint4*>(p.y_bf16 + (size_t)(R0 + r) * p.ld_out + ch * 8) =
                            f_resid ? add8_bf16(uq, rq[j]) : uq;
                    } else {
                        float u[8];
                        unpack8(uq, u);
                        if (f_resid) acc8(rq[j], u);
                        float4* d = reinterpret_cast<float4*>(p.y_f32 + (size_t)(R0 + r) * p.ld_out + ch * 8);
                        d[0] = make_float4(u[0], u[1], u[2], u[3]);
                        d[1] = make_float4(u[4], u[5], u[6], u[7]);
                    }
                }
            }
            slot_sync(g);        // buf and sseg are free for the next tile
            tick(7);
        }
        if (prof) atomicAdd(p.prof + 15, 1ull);
    }

    if ((maps.use & (7u * kMapH)) && wis == 0 && elect_one()) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base, kTmemCols);
}

// One thread per segment classifies it (complete inside one sub-tile: nothing to do; empty: zero row;
// spanning sub-tiles: sum of its boundary partials) from two coalesced rowptr reads; the warp then
// walks its flagged segments together, 16 bytes of the row per lane.
template <typename OutT>
__global__ void __launch_bounds__(256) seg_fixup_kernel(const int32_t* __restrict__ rowptr, int num_segments, int H, int sub_shift,
                                                        const float* __restrict__ bnd, OutT* __restrict__ out) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int s = 0, e = 0;
    bool flagged = false;
    if (seg < num_segments) {
        s = rowptr[seg];
        e = rowptr[seg + 1];
        flagged = (s == e) || ((s >> sub_shift) != ((e - 1) >> sub_shift));
    }
    // the graph layout above is older than the previous kernel; its boundary partials are read from here on
    pdl_wait();
    pdl_launch_dependents();
    unsigned todo = __ballot_sync(0xffffffffu, flagged);
    const int c4 = lane * 4;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int sg = __shfl_sync(0xffffffffu, seg, src);
        const int ss = __shfl_sync(0xffffffffu, s, src), ee = __shfl_sync(0xffffffffu, e, src);
        if (c4 < H) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ss != ee) {
                const int t0 = ss >> sub_shift, t1 = (ee - 1) >> sub_shift;
                acc = *reinterpret_cast<const float4*>(bnd + ((size_t)t0 * 2 + 1) * H + c4);
                for (int t = t0 + 1; t <= t1; ++t) {
                    const float4 b = *reinterpret_cast<const float4*>(bnd + ((size_t)t * 2) * H + c4);
                    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                }
            }
            if constexpr (sizeof(OutT) == 4)
                *reinterpret_cast<float4*>(out + (size_t)sg * H + c4) = acc;
            else
                *reinterpret_cast<uint2*>(out + (size_t)sg * H + c4) = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
        }
    }
}

template <int H, int NG, int kMode>
int launch_fwd(const gp_mlp_fwd_args& a, cudaStream_t st) {
    size_t smem = 1024;
    for (int l = 0; l < a.n_layers; ++l) smem += (size_t)((a.k[l] + 63) / 64) * a.n[l] * 128;
    smem += (size_t)NG * kBufBytes + 4 * kBiasStride * 4 + 128 * 4 + NG * 256 * 4 + NG * 144 * 4;
    GP_REQUIRE((int)smem <= gp::max_smem_optin(), "gp_mlp_fwd: needs %zu B of shared memory (> %d)", smem,
               gp::max_smem_optin());
    static int smem_set = 0;      // raised once per instantiation (and never inside a stream capture twice)
    if ((int)smem > smem_set) {
        GP_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<H, NG, kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = (int)smem;
    }
    const int n_tiles = (a.rows + 127) / 128;
    int grid = (n_tiles + NG - 1) / NG;
    if (grid > gp::sm_count()) grid = gp::sm_count();
    if (grid < 1) grid = 1;
    FwdMaps maps;
    memset(&maps, 0, sizeof(maps));
    static const bool no_tma = getenv("GP_NO_TMA") != nullptr;
    if (!no_tma) {
        // with gathered pre-activations the tile buffer is busy until just before layer 0, where the
        // shorter latency of per-thread cp.async (L2-prefetched rows) beats one bulk copy
        if (a.a_bf16 && !a.init && gp::tma_map_2d(&maps.a, a.a_bf16, a.rows, a.ka, a.lda)) maps.use |= kMapA;
        gp_bf16* const sv[3] = {a.save_h1, a.save_h2, a.save_h3};
        for (int i = 0; i < 3; ++i)
            if (sv[i] && gp::tma_map_2d(&maps.h[i], sv[i], a.rows, H, H)) maps.use |= kMapH << i;
    }
    GP_CHECK_CUDA(gp::launch_kernel(mlp_fwd_kernel<H, NG, kMode>, dim3(grid), dim3(kSlotThreads * NG), smem, st, a, maps));
    return 0;
}
}  // namespace

extern "C" int gp_mlp_fwd(const gp_mlp_fwd_args* args, int hidden, void* stream) {
    GP_REQUIRE(args != nullptr, "gp_mlp_fwd: null args");
    const gp_mlp_fwd_args& a = *args;
    GP_REQUIRE(a.rows >= 0, "gp_mlp_fwd: rows < 0");
    if (a.rows == 0) return 0;
    GP_REQUIRE(a.n_layers >= 1 && a.n_layers <= 4, "gp_mlp_fwd: n_layers must be 1..4");
    GP_REQUIRE(a.ka > 0 && a.ka % 16 == 0 && a.ka <= 128 && a.k[0] == a.ka, "gp_mlp_fwd: bad ka=%d (k[0]=%d)", a.ka, a.k[0]);
    GP_REQUIRE((a.a_bf16 != nullptr) != (a.a_f32 != nullptr), "gp_mlp_fwd: exactly one of a_bf16 / a_f32");
    GP_REQUIRE(a.lda % 8 == 0, "gp_mlp_fwd: lda must be a multiple of 8");
    for (int l = 0; l < a.n_layers; ++l) {
        GP_REQUIRE(a.w[l] != nullptr && a.k[l] % 16 == 0 && a.n[l] % 16 == 0 && a.n[l] <= 384 && a.k[l] <= 128,
                   "gp_mlp_fwd: layer %d has bad shape n=%d k=%d", l, a.n[l], a.k[l]);
        if (l + 1 < a.n_layers)
            GP_REQUIRE(a.n[l] == hidden && a.k[l + 1] == hidden, "gp_mlp_fwd: hidden layers must be %d wide", hidden);
    }
    if (a.init) GP_REQUIRE(a.n[0] == hidden && a.ld_init % 8 == 0 && a.init_off0 % 8 == 0 && a.init_off1 % 8 == 0,
                           "gp_mlp_fwd: init rows need n[0]==hidden and 16-byte aligned offsets");
    if (a.norm_scale) GP_REQUIRE(a.n[a.n_layers - 1] == hidden, "gp_mlp_fwd: RMSNorm needs n_last == hidden");
    if (a.seg_id) GP_REQUIRE(a.norm_scale && ((a.seg_out != nullptr) != (a.seg_out_bf16 != nullptr)) && a.seg_bnd,
                             "gp_mlp_fwd: segment sum needs norm, seg_bnd and exactly one of seg_out / seg_out_bf16");
    if (a.save_h1) GP_REQUIRE(a.n_layers >= 2, "gp_mlp_fwd: save_h1 needs at least 2 layers");
    if (a.save_h2) GP_REQUIRE(a.n_layers >= 3, "gp_mlp_fwd: save_h2 needs at least 3 layers");
    if (a.save_h3) GP_REQUIRE(a.n_layers >= 4, "gp_mlp_fwd: save_h3 needs 4 layers");
    GP_REQUIRE(a.n_layers == 1 || a.n[a.n_layers - 1] <= hidden, "gp_mlp_fwd: a multi-layer MLP cannot widen in its last layer");
    GP_REQUIRE((a.y_bf16 != nullptr) != (a.y_f32 != nullptr), "gp_mlp_fwd: exactly one of y_bf16 / y_f32");
    if (a.norm_scale) GP_REQUIRE(a.ld_out % 8 == 0, "gp_mlp_fwd: ld_out must be a multiple of 8 with RMSNorm");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {   // wide single-layer projections have their own kernel (linear_fwd.cu)
        const int taken = gp::try_linear_fwd(a, hidden, st);
        if (taken != 0) return taken < 0 ? taken : 0;
    }
    // the processor's edge and node MLPs get instantiations with every option fixed at compile time
    bool proc = a.init && a.n_layers == 4 && a.ka == hidden && a.norm_scale && a.resid && a.y_bf16;
    for (int l = 0; l < a.n_layers && proc; ++l) proc = a.k[l] == hidden && a.n[l] == hidden;
    const bool edge = proc && a.two_inits && a.idx0 && a.idx1 && a.a_bf16 && a.seg_id && a.seg_out_bf16;
    const bool node = proc && !a.two_inits && !a.idx0 && a.a_bf16 && !a.seg_id;
    if (edge && hidden == 128) {      // the CTA-pair, warp-specialised edge kernel (edge_fwd2.cu)
        const int taken = gp::try_edge_fwd2(a, st);
        if (taken != 0) return taken < 0 ? taken : 0;
    }
    switch (hidden) {
        case 128:
            return edge ? launch_fwd<128, 2, 1>(a, st) : node ? launch_fwd<128, 2, 2>(a, st) : launch_fwd<128, 2, 0>(a, st);
        case 64: return launch_fwd<64, 2, 0>(a, st);
        case 32: return launch_fwd<32, 2, 0>(a, st);
        default: gp::set_error("gp_mlp_fwd: unsupported hidden size %d (32, 64, 128)", hidden); return -1;
    }
}

extern "C" int gp_seg_sub_rows(int32_t hidden, int32_t backward) {
    // rows per sub-tile = 64 * hidden / (threads walking a tile): forward 256 threads, backward 512 at hidden 128
    return (backward && hidden >= 128) ? hidden / 8 : hidden / 4;
}

template <typename OutT>
static int seg_fixup_launch(const int32_t* rowptr, int32_t num_segments, int32_t hidden, int32_t sub_rows, const float* seg_bnd,
                            OutT* seg_out, void* stream) {
    if (num_segments <= 0) return 0;
    GP_REQUIRE(sub_rows > 0 && (sub_rows & (sub_rows - 1)) == 0, "gp_seg_fixup: sub_rows must be a power of two");
    GP_REQUIRE(hidden % 4 == 0 && hidden <= 128, "gp_seg_fixup: hidden must be a multiple of 4, at most 128");
    int shift = 0;
    while ((1 << shift) < sub_rows) ++shift;
    const int threads = 256;
    const int blocks = (num_segments + threads - 1) / threads;
    GP_CHECK_CUDA(gp::launch_kernel(seg_fixup_kernel<OutT>, dim3(blocks), dim3(threads), 0, static_cast<cudaStream_t>(stream), rowptr,
                                    (int)num_segments, (int)hidden, shift, seg_bnd, seg_out));
    return 0;
}

extern "C" int gp_seg_fixup(const int32_t* rowptr, int32_t num_segments, int32_t hidden, int32_t sub_rows,
                            const float* seg_bnd, float* seg_out, void* stream) {
    return seg_fixup_launch<float>(rowptr, num_segments, hidden, sub_rows, seg_bnd, seg_out, stream);
}

extern "C" int gp_seg_fixup_bf16(const int32_t* rowptr, int32_t num_segments, int32_t hidden, int32_t sub_rows,
                                 const float* seg_bnd, gp_bf16* seg_out_bf16, void* stream) {
    return seg_fixup_launch<gp_bf16>(rowptr, num_segments, hidden, sub_rows, seg_bnd, seg_out_bf16, stream);
}
