// tile_util.cuh -- small per-thread helpers shared by the row-tile kernels.
#pragma once
#include "tc5.cuh"
#include "../../include/gp_b200.h"

namespace gp {
using namespace tc5;

constexpr int kBufBytes = 128 * 128 * 2;   // one activation buffer: 128 rows x up to 128 bf16

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
    f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
    f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
// f[0..7] += the eight bf16 values of q (one FHADD each, no separate unpack)
__device__ __forceinline__ void acc8(const uint4& q, float* f) {
    f[0] = add_bf16_lo(q.x, f[0]); f[1] = add_bf16_hi(q.x, f[1]); f[2] = add_bf16_lo(q.y, f[2]); f[3] = add_bf16_hi(q.y, f[3]);
    f[4] = add_bf16_lo(q.z, f[4]); f[5] = add_bf16_hi(q.z, f[5]); f[6] = add_bf16_lo(q.w, f[6]); f[7] = add_bf16_hi(q.w, f[7]);
}
__device__ __forceinline__ uint4 pack8_relu(const float* f) {
    return make_uint4(pack_bf16_relu(f[0], f[1]), pack_bf16_relu(f[2], f[3]), pack_bf16_relu(f[4], f[5]),
                      pack_bf16_relu(f[6], f[7]));
}
__device__ __forceinline__ uint4 mask8_pos(const uint4& v, const uint4& h) {
    return make_uint4(mask_pos_bf16x2(v.x, h.x), mask_pos_bf16x2(v.y, h.y), mask_pos_bf16x2(v.z, h.z), mask_pos_bf16x2(v.w, h.w));
}
__device__ __forceinline__ uint4 add8_bf16(const uint4& a, const uint4& b) {
    return make_uint4(add_bf16x2(a.x, b.x), add_bf16x2(a.y, b.y), add_bf16x2(a.z, b.z), add_bf16x2(a.w, b.w));
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// Stage a packed bf16 weight [n][k] (row-major) into an SW128 row tile with n rows.
__device__ __forceinline__ void stage_weight(uint8_t* tile, const gp_bf16* w, int n, int k) {
    const int kc = k >> 3, total = n * kc;
    const uint32_t ws = smem_u32(tile);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / kc, ch = i - r * kc;
        cp_async16(ws + sw128_off(n, r, ch * 8), w + (size_t)r * k + ch * 8);
    }
}

// Stage 128 rows x ka columns of a row-major global matrix (bf16, or fp32 converted on the
// fly) into an SW128 row tile.  Rows past `rows` replicate the last valid row.
__device__ __forceinline__ void stage_rows(uint8_t* buf, const gp_bf16* a_bf16, const float* a_f32, int ka, int lda,
                                           int R0, int rows, int t, int nthreads) {
    const int kc = ka >> 3;                       // 16-byte chunks per row: 2, 4, 8 or 16
    const int sh = 31 - __clz(kc);
    const uint32_t buf_s = smem_u32(buf);
    if (a_bf16) {
        for (int i = t; i < 128 * kc; i += nthreads) {
            const int r = i >> sh, ch = i & (kc - 1);
            const int gr = min(R0 + r, rows - 1);
            cp_async16(buf_s + sw128_off(128, r, ch * 8), a_bf16 + (size_t)gr * lda + ch * 8);
        }
    } else if (a_f32) {
        // four chunks per round: their eight 16-byte loads are in flight together
        for (int i0 = t; i0 < 128 * kc; i0 += 4 * nthreads) {
            float4 u[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * nthreads;
                if (i < 128 * kc) {
                    const int r = i >> sh, ch = i & (kc - 1);
                    const float4* s = reinterpret_cast<const float4*>(a_f32 + (size_t)min(R0 + r, rows - 1) * lda + ch * 8);
                    u[2 * j] = __ldg(s);
                    u[2 * j + 1] = __ldg(s + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * nthreads;
                if (i < 128 * kc) {
                    const int r = i >> sh, ch = i & (kc - 1);
                    *reinterpret_cast<uint4*>(buf + sw128_off(128, r, ch * 8)) =
                        make_uint4(pack_bf16(u[2 * j].x, u[2 * j].y), pack_bf16(u[2 * j].z, u[2 * j].w),
                                   pack_bf16(u[2 * j + 1].x, u[2 * j + 1].y), pack_bf16(u[2 * j + 1].z, u[2 * j + 1].w));
                }
            }
        }
    }
}

// Accumulator pre-load: fp32 sum of one or two gathered bf16 rows -> NC TMEM columns starting at
// `tacc`.  All row loads are issued before the first use so one memory round trip covers them.
template <int NC>
__device__ __forceinline__ void init_rows_to_tmem(uint32_t tacc, const gp_bf16* r0p, const gp_bf16* r1p) {
    uint4 q0[NC / 8], q1[NC / 8];
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) q0[i] = ldg16(r0p + i * 8);
    if (r1p) {
#pragma unroll
        for (int i = 0; i < NC / 8; ++i) q1[i] = ldg16(r1p + i * 8);
    }
#pragma unroll
    for (int c = 0; c < NC; c += 16) {
        float f[16];
        unpack8(q0[c / 8], f);
        unpack8(q0[c / 8 + 1], f + 8);
        if (r1p) {
            float h[16];
            unpack8(q1[c / 8], h);
            unpack8(q1[c / 8 + 1], h + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += h[j];
        }
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(f[j]);
        tmem_st16(tacc + c, v);
    }
    tmem_st_wait();
}

// Same pre-load, but one of the two sources has already been gathered into the SW128 tile `buf`
// (row-major 16-byte chunks, one cache line per 8 lanes) and is read back row-wise from there;
// `dp` (may be null) is this thread's part of the directly loaded row.  Columns [c0, c0+NC).
template <int NC>
__device__ __forceinline__ void init_staged_to_tmem(uint32_t tacc, const uint8_t* buf, int row, int c0,
                                                    const gp_bf16* dp) {
    uint4 q1[NC / 8];
    if (dp) {
#pragma unroll
        for (int i = 0; i < NC / 8; ++i) q1[i] = ldg16(dp + i * 8);
    }
#pragma unroll
    for (int c = 0; c < NC; c += 16) {
        float f[16];
        unpack8(*reinterpret_cast<const uint4*>(buf + sw128_off(128, row, c0 + c)), f);
        unpack8(*reinterpret_cast<const uint4*>(buf + sw128_off(128, row, c0 + c + 8)), f + 8);
        if (dp) {
            float h[16];
            unpack8(q1[c / 8], h);
            unpack8(q1[c / 8 + 1], h + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += h[j];
        }
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(f[j]);
        tmem_st16(tacc + c, v);
    }
    tmem_st_wait();
}

// fp32 staging tile [128 rows x ncols] that turns a row-per-lane TMEM read-out into row-major,
// fully coalesced global stores: 16-byte chunk c of row r lives at chunk c ^ (r & (SW-1)),
// SW = min(largest power of two dividing the chunks per row, 32) (conflict-free for the lane-per-row writes and the row-major reads).
__device__ __forceinline__ uint32_t stage_f32_off(int r, int chunk, int ncols) {
    const int cpr = ncols >> 2;
    const int low = cpr & -cpr;                     // largest power of two dividing the chunk count
    const int sw = low < 32 ? low : 32;
    return (uint32_t)r * (uint32_t)(ncols * 4) + (uint32_t)((chunk ^ (r & (sw - 1))) << 4);
}

// Copy a TMEM-resident fp32 matrix (lane r = row r, columns [col0, col0 + ncols) of this warp's
// lanes at `tlane`) to global memory, rows [0, nrows), through the staging tile `stg`
// (>= 128 * ncols * 4 bytes).  Called by all NT threads of the CTA; part = tid >> 7.
template <int NT>
__device__ __forceinline__ void tmem_rows_to_global(uint32_t tlane, uint32_t col0, int nrows, int ncols, float* dst,
                                                    int ld_dst, uint8_t* stg, int tid) {
    const int row = tid & 127, part = tid >> 7;
    for (int c0 = part * 16; c0 < ncols; c0 += (NT / 128) * 16) {
        uint32_t v[16];
        tmem_ld16(tlane + col0 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stg + stage_f32_off(row, c0 / 4 + q, ncols)) =
                make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncthreads();
    const int cpr = ncols >> 2;
    for (int i = tid; i < nrows * cpr; i += NT) {
        const int r = i / cpr, ch = i - r * cpr;
        reinterpret_cast<float4*>(dst + (size_t)r * ld_dst)[ch] = *reinterpret_cast<const float4*>(stg + stage_f32_off(r, ch, ncols));
    }
    __syncthreads();
}

// Segment sum over the rows of one 128-row tile held in `buf` (bf16, SW128 layout, H columns),
// by NT threads: thread t owns column pair (t % (H/2)) of the sub-tile of SUB = 64*H/NT rows
// number t / (H/2).  sseg[3] = segment id of the row before the tile, sseg[4..131] = rows,
// sseg[132] = row after; -1 marks "no row".  A piece that neither continues from the previous
// sub-tile nor into the next one is a complete segment and goes to seg_out (fp32) or seg_out_bf16; other pieces go to
// seg_bnd[sub-tile][0 = continues from before | 1 = continues after] for gp_seg_fixup.
template <int H, int NT>
__device__ __forceinline__ void tile_segment_sum(const uint8_t* buf, const int* sseg, int R0, int t, float* seg_out,
                                                 float* seg_bnd, gp_bf16* seg_out_bf16 = nullptr,
                                                 uint32_t block_stride = 128 * 128) {
    constexpr int NP = H / 2;            // column pairs
    constexpr int SUB = 128 * NP / NT;   // rows per sub-tile
    static_assert(SUB % 8 == 0, "sub-tile must be a multiple of 8 rows");
    const int part = t / NP, cp = t - part * NP;
    const int rb = part * SUB;
    const int c = cp * 2;
    const size_t sub_index = (size_t)(R0 + rb) / SUB;
    const uint32_t chunk = (c & 63) >> 3;
    const uint8_t* colbase = buf + (c >> 6) * block_stride + (c & 7) * 2;     // block_stride: bytes between 64-column blocks
    const int seg_prev = sseg[3 + rb], seg_next = sseg[4 + rb + SUB];
    int cur = sseg[4 + rb];
    bool first_piece = true;
    float s0 = 0.f, s1 = 0.f;
    auto flush = [&](int seg, bool last_piece) {
        if (seg >= 0) {
            const bool before = first_piece && (seg_prev == seg);
            const bool after = last_piece && (seg_next == seg);
            if (!before && !after && seg_out_bf16) {       // complete segment, bf16 output: rounded once, here
                *reinterpret_cast<uint32_t*>(seg_out_bf16 + (size_t)seg * H + c) = pack_bf16(s0, s1);
            } else {
                float* d = (!before && !after) ? seg_out + (size_t)seg * H + c
                                               : seg_bnd + (sub_index * 2 + (before ? 0 : 1)) * H + c;
                *reinterpret_cast<float2*>(d) = make_float2(s0, s1);
            }
        }
        first_piece = false;
    };
#pragma unroll 1
    for (int r8 = rb; r8 < rb + SUB; r8 += 8) {
        const int4 sa = *reinterpret_cast<const int4*>(sseg + 4 + r8);
        const int4 sb = *reinterpret_cast<const int4*>(sseg + 8 + r8);
        const int sid[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const uint8_t* rowbase = colbase + r8 * 128;
        uint32_t w[8];           // all eight rows requested before the first (store-carrying) branch
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const uint32_t*>(rowbase + j * 128 + ((chunk ^ j) << 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (sid[j] != cur) {
                flush(cur, false);
                cur = sid[j]; s0 = 0.f; s1 = 0.f;
            }
            s0 = add_bf16_lo(w[j], s0);
            s1 = add_bf16_hi(w[j], s1);
        }
    }
    flush(cur, true);
}


// Predicated global stores (no branch): a taken branch costs an instruction-fetch bubble in these large kernels.
__device__ __forceinline__ void st_global_u32_if(void* ptr, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.b32 [%0], %1;\n\t}" ::"l"(ptr), "r"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_global_f2_if(void* ptr, float a, float b, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(ptr), "f"(a), "f"(b),
                 "r"((uint32_t)pred)
                 : "memory");
}

// tile_segment_sum for bf16 segment outputs, straight-line: same partition (thread = column pair x sub-tile of
// SUB = 64*H/NT rows), same summation order and the same outputs bit for bit, but the row loop is fully unrolled with
// predicated stores instead of one branch per segment change (the branchy walk stalls on instruction fetch at every
// reconvergence point: ncu source view of the round-2 edge kernels).
template <int H, int NT>
__device__ __forceinline__ void tile_segment_sum_bf16_flat(const uint8_t* buf, const int* sseg, int R0, int t, float* seg_bnd,
                                                           gp_bf16* seg_out_bf16, uint32_t block_stride = 128 * 128) {
    constexpr int NP = H / 2;
    constexpr int SUB = 128 * NP / NT;
    static_assert(SUB % 8 == 0, "sub-tile must be a multiple of 8 rows");
    const int part = t / NP, cp = t - part * NP;
    const int rb = part * SUB, c = cp * 2;
    const size_t sub_index = (size_t)(R0 + rb) / SUB;
    const uint32_t chunk = (c & 63) >> 3;
    const uint8_t* colbase = buf + (c >> 6) * block_stride + (c & 7) * 2;
    const int seg_prev = sseg[3 + rb], seg_next = sseg[4 + rb + SUB];
    float* const bnd0 = seg_bnd + (sub_index * 2) * H + c;           // piece continuing from the previous sub-tile
    int cur = sseg[4 + rb];
    bool first = true;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r8 = 0; r8 < SUB; r8 += 8) {
        const int4 sa = *reinterpret_cast<const int4*>(sseg + 4 + rb + r8);
        const int4 sb = *reinterpret_cast<const int4*>(sseg + 8 + rb + r8);
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

}  // namespace gp
