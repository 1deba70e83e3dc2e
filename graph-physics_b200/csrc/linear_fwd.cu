// linear_fwd.cu -- single-layer row-tile GEMM  Y = X . W^T (+ b), bf16 in / bf16 out, for wide outputs
// (the per-node projection P = x . [W1d; W1s; W1x]^T of every message-passing step, N = 3H).
// gp_mlp_fwd routes 1-layer calls with N a multiple of 128 here.
//
// One persistent CTA per SM, 256 threads.  W stays in shared memory; X tiles arrive by TMA into a
// two-deep ring (the next tile's load is issued before this tile's MMAs), all N/128 column chunks
// of a tile are issued as one MMA batch into separate TMEM column ranges, and every chunk leaves
// through a bf16 staging tile and one TMA store (two staging tiles, so a store drains while the
// next chunk is converted).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tile_util.cuh"

namespace {
using namespace gp;

struct LinFwdMaps {
    CUtensorMap x, y;
};

template <int H>
__global__ void __launch_bounds__(256, 1) linear_fwd_kernel(const gp_mlp_fwd_args p, const __grid_constant__ LinFwdMaps maps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = GP_SMEM_ALIGNED(smem_raw);
    __shared__ uint64_t mma_bar, x_bar[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, row = tid & 127, half = tid >> 7;
    const int warp = warp_uniform(tid >> 5);
    const int N = p.n[0];                      // multiple of 128, <= 384
    const int nchunk = N >> 7;
    constexpr int KB = (H + 63) >> 6;          // 64-column blocks of the K dimension
    uint32_t off = 0;
    uint8_t* w_t = smem + off;   off += KB * N * 128;
    uint8_t* xbuf = smem + off;  off += 2 * kBufBytes;
    uint8_t* stg = smem + off;   off += 2 * kBufBytes;
    float* sbias = reinterpret_cast<float*>(smem + off);

    stage_weight(w_t, p.w[0], N, H);
    cp_async_commit();
    for (int i = tid; i < N; i += 256) sbias[i] = p.bias[0] ? p.bias[0][i] : 0.f;
    if (tid == 0) {
        mbar_init(&mma_bar, 1);
        mbar_init(&x_bar[0], 1);
        mbar_init(&x_bar[1], 1);
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
    const uint32_t x_s = smem_u32(xbuf), w_s = smem_u32(w_t), stg_s = smem_u32(stg);
    const int n_tiles = (p.rows + 127) >> 7;
    const bool has_bias = p.bias[0] != nullptr;
    uint32_t phase = 0, xphase[2] = {0, 0};
    int chunk_count = 0;                        // output chunks issued so far (staging tile = count & 1)

    auto load_x = [&](int tile, int slot) {     // elected thread only
        mbar_arrive_expect_tx(&x_bar[slot], KB * 16384u);
        for (int b = 0; b < KB; ++b) tma_load_2d(x_s + slot * kBufBytes + b * 16384, &maps.x, b * 64, tile << 7, &x_bar[slot]);
    };
    if ((int)blockIdx.x < n_tiles && warp == 0 && elect_one()) load_x(blockIdx.x, 0);

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int slot = it & 1;
        const int R0 = tile << 7;
        // the other ring slot was last read by the previous tile's MMAs, which have completed
        if (tile + (int)gridDim.x < n_tiles && warp == 0 && elect_one()) load_x(tile + gridDim.x, slot ^ 1);
        mbar_wait(&x_bar[slot], xphase[slot]);
        xphase[slot] ^= 1;
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            const uint32_t idesc = idesc_bf16(128, false, false);
            for (int c = 0; c < nchunk; ++c)
                for (int ks = 0; ks < (H >> 4); ++ks)
                    mma_ss(tmem + c * 128, desc_kmajor(x_s + slot * kBufBytes, 128, ks), desc_kmajor(w_s + c * 128 * 128, N, ks), idesc,
                           ks > 0 ? 1u : 0u);
            mma_commit(&mma_bar);
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        for (int c = 0; c < nchunk; ++c, ++chunk_count) {
            uint8_t* sb = stg + (chunk_count & 1) * kBufBytes;
            // this staging tile was handed to a bulk store two chunks ago: its reads must be done
            if (chunk_count >= 2) {
                if (warp == 0 && elect_one()) tma_store_wait_read<1>();
                __syncthreads();
            }
            uint32_t v[64];
#pragma unroll
            for (int q = 0; q < 64; q += 16) tmem_ld16(tlane + c * 128 + half * 64 + q, *reinterpret_cast<uint32_t(*)[16]>(&v[q]));
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 64; q += 8) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q + j]) + (has_bias ? sbias[c * 128 + half * 64 + q + j] : 0.f);
                *reinterpret_cast<uint4*>(sb + sw128_off(128, row, half * 64 + q)) = pack8(f);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) {
                const uint32_t ss = stg_s + (chunk_count & 1) * kBufBytes;
                tma_store_2d(&maps.y, c * 128, R0, ss);
                tma_store_2d(&maps.y, c * 128 + 64, R0, ss + 16384);
                tma_store_commit();
            }
        }
        tc_fence_before();
        __syncthreads();       // every accumulator has been read: the next tile's MMAs may overwrite them
    }
    if (warp == 0 && elect_one()) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}
}  // namespace

namespace gp {
// Returns 1 when the call was taken (launched), 0 when it does not fit this kernel, < 0 on error.
int try_linear_fwd(const gp_mlp_fwd_args& a, int hidden, cudaStream_t st) {
    if (hidden != 128 || a.n_layers != 1 || a.init || a.norm_scale || a.resid || a.seg_id || !a.a_bf16 || !a.y_bf16 ||
        a.ka != hidden || a.k[0] != hidden || a.n[0] % 128 != 0 || a.n[0] > 384 || a.n_valid != a.n[0] || a.save_h1 || a.save_h2 ||
        a.save_h3 || a.prof || getenv("GP_NO_TMA"))
        return 0;
    LinFwdMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (!tma_map_2d(&maps.x, a.a_bf16, a.rows, hidden, a.lda) || !tma_map_2d(&maps.y, a.y_bf16, a.rows, a.n[0], a.ld_out)) return 0;
    const size_t smem = 1024 + (size_t)2 * a.n[0] * 128 + 4 * kBufBytes + 384 * 4;
    if ((int)smem > max_smem_optin()) return 0;
    static int smem_set = 0;
    if ((int)smem > smem_set) {
        GP_CHECK_CUDA(cudaFuncSetAttribute(linear_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = (int)smem;
    }
    const int n_tiles = (a.rows + 127) / 128;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    GP_CHECK_CUDA(gp::launch_kernel(linear_fwd_kernel<128>, dim3(grid), dim3(256), smem, st, a, maps));
    return 1;
}
}  // namespace gp
