// mlp_fwd.cu -- fused row-tile MLP forward on tcgen05 (see include/gp_b200.h, gp_mlp_fwd).
//
// Persistent kernel, one CTA per SM.  The packed bf16 weights of all layers stay resident in
// shared memory (SW128 row tiles) for the whole launch.  A CTA runs NG independent worker
// groups of 128 threads; each group owns one 128-row activation buffer, one 128-column fp32
// accumulator in TMEM and one mbarrier, and walks its own tiles.  Inside a group the flow is
// bulk-synchronous (stage -> MMA -> epilogue per layer); the two groups run out of phase, so
// one group's tcgen05.mma overlaps the other group's TMEM epilogue.
//
// Thread r of a group owns row r of the tile == TMEM lane r: bias, ReLU, RMSNorm and the
// residual are thread-local.  The receiver-sorted segment sum re-partitions through shared
// memory (column pairs x sub-tiles of H/2 rows) and uses no atomics.
#include "common.cuh"
#include "tile_util.cuh"

namespace {
using namespace gp;

constexpr int kBiasStride = 384;

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }




template <int H, int NG>
__global__ void __launch_bounds__(128 * NG, 1) mlp_fwd_kernel(const gp_mlp_fwd_args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = gp::align1024(smem_raw);
    __shared__ uint64_t mma_bar[NG];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int g = tid >> 7;           // worker group
    const int row = tid & 127;        // row in tile == TMEM lane
    const int L = p.n_layers;

    // ---- carve shared memory (all offsets uniform across the CTA)
    uint32_t w_off[4];
    uint32_t off = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        w_off[l] = off;
        if (l < L) off += ((p.k[l] + 63) >> 6) * p.n[l] * 128;
    }
    uint8_t* buf = smem + off + g * kBufBytes;
    off += NG * kBufBytes;
    float* sbias = reinterpret_cast<float*>(smem + off);
    off += 4 * kBiasStride * 4;
    float* sscale = reinterpret_cast<float*>(smem + off);
    off += 128 * 4;
    int* sseg = reinterpret_cast<int*>(smem + off) + g * 136;

    // ---- one-time staging: weights, biases, barriers, TMEM
    for (int l = 0; l < L; ++l) {
        const int kc = p.k[l] >> 3, total = p.n[l] * kc;
        const uint32_t ws = smem_u32(smem + w_off[l]);
        for (int i = tid; i < total; i += blockDim.x) {
            const int r = i / kc, ch = i - r * kc;
            cp_async16(ws + sw128_off(p.n[l], r, ch * 8), p.w[l] + (size_t)r * p.k[l] + ch * 8);
        }
        for (int i = tid; i < p.n[l]; i += blockDim.x) sbias[l * kBiasStride + i] = p.bias[l] ? p.bias[l][i] : 0.f;
    }
    cp_async_commit();
    if (p.norm_scale)
        for (int i = tid; i < H; i += blockDim.x) sscale[i] = p.norm_scale[i];
    if (tid == 0) {
        for (int i = 0; i < NG; ++i) mbar_init(&mma_bar[i], 1);
        fence_mbar_init();
    }
    if (tid < 32) tmem_alloc(&tmem_slot, NG * 128);
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tmem_base = tmem_slot;
    const uint32_t tacc_mma = tmem_base + g * 128;                               // D operand (lane 0)
    const uint32_t tacc = tmem_addr(tmem_base, (row >> 5) * 32, g * 128);         // this warp's lanes
    const uint32_t buf_s = smem_u32(buf);
    uint32_t phase = 0;
    const bool has_init = p.init != nullptr;
    const int n_tiles = (p.rows + 127) >> 7;

    for (int tile = blockIdx.x * NG + g; tile < n_tiles; tile += gridDim.x * NG) {
        const int R0 = tile << 7;
        const int grow = R0 + row;
        const bool valid = grow < p.rows;
        const int crow = valid ? grow : p.rows - 1;

        // (a) streamed layer-0 operand -> buf (rows past the end replicate the last row; their
        //     results are never stored)
        {
            const int kc = p.ka >> 3;
            if (p.a_bf16) {
                for (int i = row; i < 128 * kc; i += 128) {
                    const int r = i / kc, ch = i - r * kc;
                    const int gr = min(R0 + r, p.rows - 1);
                    cp_async16(buf_s + sw128_off(128, r, ch * 8), p.a_bf16 + (size_t)gr * p.lda + ch * 8);
                }
            } else if (p.a_f32) {
                for (int i = row; i < 128 * kc; i += 128) {
                    const int r = i / kc, ch = i - r * kc;
                    const int gr = min(R0 + r, p.rows - 1);
                    const float4* s = reinterpret_cast<const float4*>(p.a_f32 + (size_t)gr * p.lda + ch * 8);
                    const float4 u0 = __ldg(s), u1 = __ldg(s + 1);
                    const uint4 pk = make_uint4(pack_bf16(u0.x, u0.y), pack_bf16(u0.z, u0.w), pack_bf16(u1.x, u1.y),
                                                pack_bf16(u1.z, u1.w));
                    *reinterpret_cast<uint4*>(buf + sw128_off(128, r, ch * 8)) = pk;
                }
            }
            cp_async_commit();
        }
        // (b) segment ids of the tile (+ one row of context on each side)
        if (p.seg_id) {
            sseg[1 + row] = valid ? __ldg(p.seg_id + grow) : -1;
            if (row == 0) {
                sseg[0] = R0 > 0 ? __ldg(p.seg_id + R0 - 1) : -1;
                sseg[129] = (R0 + 128 < p.rows) ? __ldg(p.seg_id + R0 + 128) : -1;
            }
        }
        // (c) accumulator pre-load: gathered pre-activation rows (fp32 sum of bf16 rows)
        if (has_init) {
            const int i0 = p.idx0 ? __ldg(p.idx0 + crow) : crow;
            const gp_bf16* r0p = p.init + (size_t)i0 * p.ld_init + p.init_off0;
            const gp_bf16* r1p = nullptr;
            if (p.two_inits) {
                const int i1 = p.idx1 ? __ldg(p.idx1 + crow) : crow;
                r1p = p.init + (size_t)i1 * p.ld_init + p.init_off1;
            }
#pragma unroll 2
            for (int c = 0; c < H; c += 16) {
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
        cp_async_wait<0>();
        fence_async_smem();
        tc_fence_before();
        group_sync(g);

        // (d) layers
        for (int l = 0; l < L; ++l) {
            const int K = p.k[l];
            const int Nl = p.n[l];
            const bool last = (l == L - 1);
            const float* bl = sbias + l * kBiasStride;
            for (int nc = 0; nc < Nl; nc += 128) {
                const int ncols = min(128, Nl - nc);
                if (row == 0) {
                    tc_fence_after();
                    const uint32_t idesc = idesc_bf16(ncols, false, false);
                    const uint32_t ws = smem_u32(smem + w_off[l]) + nc * 128;
                    for (int ks = 0; ks < (K >> 4); ++ks)
                        mma_ss(tacc_mma, desc_kmajor(buf_s, 128, ks), desc_kmajor(ws, Nl, ks), idesc,
                               (ks > 0 || (l == 0 && has_init)) ? 1u : 0u);
                    mma_commit(&mma_bar[g]);
                }
                mbar_wait(&mma_bar[g], phase);
                phase ^= 1;
                tc_fence_after();

                if (!last) {
                    // hidden layer: relu(acc + b) -> bf16 -> next A operand (in place over buf)
                    for (int c0 = 0; c0 < H; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tacc + c0, v);
                        tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(__uint_as_float(v[j]) + bl[c0 + j], 0.f);
                        const uint4 q0 = pack8(f), q1 = pack8(f + 8);
                        *reinterpret_cast<uint4*>(buf + sw128_off(128, row, c0)) = q0;
                        *reinterpret_cast<uint4*>(buf + sw128_off(128, row, c0 + 8)) = q1;
                        if (l == 1 && p.save_h2 && valid) {
                            uint4* d = reinterpret_cast<uint4*>(p.save_h2 + (size_t)grow * H + c0);
                            d[0] = q0;
                            d[1] = q1;
                        }
                    }
                    fence_async_smem();
                } else if (p.norm_scale) {
                    // RMSNorm (layers.py:104-129) + residual + optional segment-sum staging
                    float ss = 0.f;
                    for (int c0 = 0; c0 < H; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tacc + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float m = __uint_as_float(v[j]) + bl[c0 + j];
                            ss = fmaf(m, m, ss);
                        }
                    }
                    const float rinv = 1.f / (sqrtf(ss * (1.f / H)) + 1e-8f);
                    for (int c0 = 0; c0 < H; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tacc + c0, v);
                        tmem_ld_wait();
                        float u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            u[j] = sscale[c0 + j] * ((__uint_as_float(v[j]) + bl[c0 + j]) * rinv);
                        if (p.seg_id) {
                            *reinterpret_cast<uint4*>(buf + sw128_off(128, row, c0)) = pack8(u);
                            *reinterpret_cast<uint4*>(buf + sw128_off(128, row, c0 + 8)) = pack8(u + 8);
                        }
                        if (valid) {
                            if (p.resid) {
                                float e[16];
                                const gp_bf16* rp = p.resid + (size_t)grow * p.ld_out + c0;
                                unpack8(ldg16(rp), e);
                                unpack8(ldg16(rp + 8), e + 8);
#pragma unroll
                                for (int j = 0; j < 16; ++j) u[j] += e[j];
                            }
                            if (p.y_bf16) {
                                uint4* d = reinterpret_cast<uint4*>(p.y_bf16 + (size_t)grow * p.ld_out + c0);
                                d[0] = pack8(u);
                                d[1] = pack8(u + 8);
                            } else {
                                float4* d = reinterpret_cast<float4*>(p.y_f32 + (size_t)grow * p.ld_out + c0);
#pragma unroll
                                for (int j = 0; j < 4; ++j) d[j] = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                            }
                        }
                    }
                } else {
                    // plain last layer (decoder / projection): acc + b, first n_valid columns
                    for (int c0 = 0; c0 < ncols; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tacc + c0, v);
                        tmem_ld_wait();
                        float u[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) u[j] = __uint_as_float(v[j]) + bl[nc + c0 + j];
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
                tc_fence_before();
                group_sync(g);
            }
        }

        // (e) receiver-sorted segment sum of bf16(u) (fp32 accumulate, fixed order, no atomics)
        if (p.seg_id) {
            constexpr int SUB = H / 2;             // rows per sub-tile == column pairs
            const int part = row / SUB, cp = row - part * SUB;
            const int rb = part * SUB, re = rb + SUB;
            const int c = cp * 2;
            const size_t sub_index = (size_t)(R0 + rb) / SUB;
            auto flush = [&](int seg, int a, int b, float s0, float s1) {
                if (seg < 0) return;
                const bool before = (a == rb) && (sseg[a] == seg);          // sseg[a] is row a-1
                const bool after = (b == re) && (sseg[1 + b] == seg);
                float* d = (!before && !after) ? p.seg_out + (size_t)seg * H + c
                                               : p.seg_bnd + (sub_index * 2 + (before ? 0 : 1)) * H + c;
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
            group_sync(g);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base, NG * 128);
}

__global__ void seg_fixup_kernel(const int32_t* __restrict__ rowptr, int num_segments, int H, int SUB,
                                 const float* __restrict__ bnd, float* __restrict__ out) {
    // one warp per segment; lanes stride the H columns
    const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (seg >= num_segments) return;
    const int s = rowptr[seg], e = rowptr[seg + 1];
    if (s == e) {
        for (int c = lane; c < H; c += 32) out[(size_t)seg * H + c] = 0.f;
        return;
    }
    const int t0 = s / SUB, t1 = (e - 1) / SUB;
    if (t0 == t1) return;
    for (int c = lane; c < H; c += 32) {
        float acc = bnd[((size_t)t0 * 2 + 1) * H + c];
        for (int t = t0 + 1; t <= t1; ++t) acc += bnd[((size_t)t * 2) * H + c];
        out[(size_t)seg * H + c] = acc;
    }
}

template <int H, int NG>
int launch_fwd(const gp_mlp_fwd_args& a, cudaStream_t st) {
    size_t smem = 1024;
    for (int l = 0; l < a.n_layers; ++l) smem += (size_t)((a.k[l] + 63) / 64) * a.n[l] * 128;
    smem += (size_t)NG * kBufBytes + 4 * kBiasStride * 4 + 128 * 4 + NG * 136 * 4;
    GP_REQUIRE((int)smem <= gp::max_smem_optin(), "gp_mlp_fwd: needs %zu B of shared memory (> %d)", smem,
               gp::max_smem_optin());
    GP_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<H, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = (a.rows + 127) / 128;
    int grid = (n_tiles + NG - 1) / NG;
    if (grid > gp::sm_count()) grid = gp::sm_count();
    if (grid < 1) grid = 1;
    mlp_fwd_kernel<H, NG><<<grid, 128 * NG, smem, st>>>(a);
    GP_CHECK_CUDA(cudaGetLastError());
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
    if (a.seg_id) GP_REQUIRE(a.norm_scale && a.seg_out && a.seg_bnd, "gp_mlp_fwd: segment sum needs norm + outputs");
    GP_REQUIRE((a.y_bf16 != nullptr) != (a.y_f32 != nullptr), "gp_mlp_fwd: exactly one of y_bf16 / y_f32");
    GP_REQUIRE(a.ld_out % 8 == 0 || !a.y_bf16 || !a.norm_scale, "gp_mlp_fwd: ld_out must be a multiple of 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (hidden) {
        case 128: return launch_fwd<128, 2>(a, st);
        case 64: return launch_fwd<64, 2>(a, st);
        case 32: return launch_fwd<32, 2>(a, st);
        default: gp::set_error("gp_mlp_fwd: unsupported hidden size %d (32, 64, 128)", hidden); return -1;
    }
}

extern "C" int gp_seg_fixup(const int32_t* rowptr, int32_t num_segments, int32_t hidden, const float* seg_bnd,
                            float* seg_out, void* stream) {
    if (num_segments <= 0) return 0;
    const int threads = 256;
    const int blocks = (int)(((size_t)num_segments * 32 + threads - 1) / threads);
    seg_fixup_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, num_segments, hidden, hidden / 2,
                                                                                 seg_bnd, seg_out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
