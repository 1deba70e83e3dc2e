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
    const int kc = ka >> 3;
    const uint32_t buf_s = smem_u32(buf);
    if (a_bf16) {
        for (int i = t; i < 128 * kc; i += nthreads) {
            const int r = i / kc, ch = i - r * kc;
            const int gr = min(R0 + r, rows - 1);
            cp_async16(buf_s + sw128_off(128, r, ch * 8), a_bf16 + (size_t)gr * lda + ch * 8);
        }
    } else if (a_f32) {
        for (int i = t; i < 128 * kc; i += nthreads) {
            const int r = i / kc, ch = i - r * kc;
            const int gr = min(R0 + r, rows - 1);
            const float4* s = reinterpret_cast<const float4*>(a_f32 + (size_t)gr * lda + ch * 8);
            const float4 u0 = __ldg(s), u1 = __ldg(s + 1);
            *reinterpret_cast<uint4*>(buf + sw128_off(128, r, ch * 8)) =
                make_uint4(pack_bf16(u0.x, u0.y), pack_bf16(u0.z, u0.w), pack_bf16(u1.x, u1.y), pack_bf16(u1.z, u1.w));
        }
    }
}

// Accumulator pre-load: fp32 sum of one or two gathered bf16 rows -> TMEM columns [c_begin, c_end).
__device__ __forceinline__ void init_rows_to_tmem(uint32_t tacc, const gp_bf16* r0p, const gp_bf16* r1p, int c_begin,
                                                  int c_end) {
    for (int c = c_begin; c < c_end; c += 16) {
        float f[16];
        unpack8(ldg16(r0p + c), f);
        unpack8(ldg16(r0p + c + 8), f + 8);
        if (r1p) {
            float h[16];
            unpack8(ldg16(r1p + c), h);
            unpack8(ldg16(r1p + c + 8), h + 8);
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

// Segment sum over the rows of one 128-row tile held in `buf` (bf16, SW128 layout, H columns).
// sseg[0] = segment of the row before the tile, sseg[1..128] = rows, sseg[129] = row after;
// -1 marks "no row".  Threads t in [0,128) take part.  See gp_mlp_fwd in gp_b200.h.
template <int H>
__device__ __forceinline__ void tile_segment_sum(const uint8_t* buf, const int* sseg, int R0, int t, float* seg_out,
                                                 float* seg_bnd) {
    constexpr int SUB = H / 2;   // rows per sub-tile == number of column pairs
    const int part = t / SUB, cp = t - part * SUB;
    const int rb = part * SUB, re = rb + SUB;
    const int c = cp * 2;
    const size_t sub_index = (size_t)(R0 + rb) / SUB;
    auto flush = [&](int seg, int a, int b, float s0, float s1) {
        if (seg < 0) return;
        const bool before = (a == rb) && (sseg[a] == seg);   // sseg[a] is row a-1
        const bool after = (b == re) && (sseg[1 + b] == seg);
        float* d = (!before && !after) ? seg_out + (size_t)seg * H + c
                                       : seg_bnd + (sub_index * 2 + (before ? 0 : 1)) * H + c;
        *reinterpret_cast<float2*>(d) = make_float2(s0, s1);
    };
    int cur = sseg[1 + rb], a = rb;
    float s0 = 0.f, s1 = 0.f;
    for (int r = rb; r < re; ++r) {
        const int s = sseg[1 + r];
        if (s != cur) {
            flush(cur, a, r, s0, s1);
            cur = s; a = r; s0 = 0.f; s1 = 0.f;
        }
        const uint32_t w = *reinterpret_cast<const uint32_t*>(buf + sw128_off(128, r, c & ~7) + (c & 7) * 2);
        s0 += bf16_lo(w);
        s1 += bf16_hi(w);
    }
    flush(cur, a, re, s0, s1);
}

}  // namespace gp
