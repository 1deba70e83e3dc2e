// tc5.cuh -- thin inline-PTX layer over the Blackwell (sm_100a) tensor path:
// tcgen05.mma / TMEM / mbarrier / async-proxy fences, plus the one shared-memory
// operand layout every kernel in this library uses.
//
// Operand layout ("SW128 row tile"): a [R x C] bf16 tile is stored as ceil(C/64)
// column blocks; block b lives at byte offset b*R*128 and holds R rows of 128 B
// (64 bf16).  Inside a row the eight 16-byte chunks are XOR-swizzled with the row
// index (chunk' = chunk ^ (row & 7)), i.e. the hardware 128B swizzle (address bits
// [4,7) ^= bits [7,10)); tile bases must be 1024-byte aligned.  The same bytes can
// be handed to tcgen05.mma either as a K-major operand (K = C) or as an MN-major
// operand (MN = C, K = R), which is what lets one buffer serve forward, dgrad and
// wgrad without a transposed copy.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One probe of the phase; the thread may be suspended by the hardware for up to `hint_ns`.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 100000u) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (after ~2 s) instead of hanging the GPU.  The time-out
// bookkeeping only starts once the first probe has failed, so the common path is probe + branch.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 32)) {
            printf("tc5: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- TMA (bulk tensor copies)
// 2-D tiled tensor map over a row-major bf16 matrix, box = 64 columns x 128 rows, SWIZZLE_128B:
// one instruction moves one 64-column block of an SW128 row tile (16 KB).  Coordinates are
// {column, row}; rows past the end of the tensor are zero-filled on load and clipped on store.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int32_t col, int32_t row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(col), "r"(row)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, int32_t col, int32_t row, uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(col), "r"(row)
                 : "memory");
}
// bring one box into L2 only (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int32_t col, int32_t row) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(row)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory sources of all committed stores have been read (the buffers may be rewritten)
template <int N = 0>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch
// A kernel launched with the programmatic-stream-serialization attribute may start while the previous
// kernel of the stream is still draining; everything it does before pdl_wait() must be independent of
// that kernel (it is used for the weight staging / TMEM allocation prologue).  No-ops otherwise.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- warp-uniform helpers
// Value of lane 0, which the compiler can treat as warp-uniform (kept in uniform registers: the
// tcgen05.mma descriptors are then built with uniform-datapath adds instead of a per-MMA
// vector-to-uniform broadcast loop).
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
// True in exactly one lane of a converged warp.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_async_smem() {   // generic-proxy smem writes -> async proxy (UMMA/TMA)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
// One full warp calls alloc / dealloc.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMEM address = (lane << 16) | column.  A warp may only touch lanes 32*(warp%4) .. +31.
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
    return base + (lane << 16) + col;
}

// 32 lanes x 32-bit, 8 / 16 consecutive columns per thread (thread i of the warp <-> lane base+i).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 format): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout [61,64) with 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major view of an SW128 row tile with `rows` rows: descriptor for the 16-wide k-step `ks`
// (ks counts 16-element steps along C).  8-row groups are 1024 B apart; LBO is unused (=16).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, uint32_t rows, uint32_t ks) {
    uint32_t a = tile_saddr + (ks >> 2) * rows * 128u + (ks & 3u) * 32u;
    return smem_desc(a, 16u, 1024u);
}
// MN-major view of the same tile (MN runs along C, K along the rows): descriptor for the
// 16-row k-step `ks`.  64-wide MN blocks are rows*128 B apart (LBO); 8-row k groups 1024 B (SBO).
// `mn_block0` selects the first 64-column block the instruction starts at.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, uint32_t rows, uint32_t ks,
                                                 uint32_t mn_block0 = 0, uint32_t lbo_override = 0xFFFFFFFFu) {
    uint32_t a = tile_saddr + mn_block0 * rows * 128u + ks * 2048u;
    uint32_t lbo = (lbo_override == 0xFFFFFFFFu) ? rows * 128u : lbo_override;
    return smem_desc(a, lbo, 1024u);
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32, M = 128.
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t n, bool a_mn_major, bool b_mn_major, uint32_t m = 128) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All MMAs issued so far by this thread arrive on `bar` when they have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- SW128 tile addressing
// Byte offset of 16-byte chunk `chunk` (0..7) of row r inside one 64-column block.
__device__ __forceinline__ uint32_t sw128_chunk_off(uint32_t r, uint32_t chunk) {
    return r * 128u + ((chunk ^ (r & 7u)) << 4);
}
// Byte offset of element column c (multiple of 8) of row r in a tile with `rows` rows.
__device__ __forceinline__ uint32_t sw128_off(uint32_t rows, uint32_t r, uint32_t c) {
    return (c >> 6) * rows * 128u + sw128_chunk_off(r, (c & 63u) >> 3);
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// Mixed-precision helpers (sm_100: FHADD.BF16 / F2FP.RELU / HFMA2.BF16_V2, one instruction each).
// c + float(low / high bf16 half of p): the bf16 -> fp32 conversion is exact, the add rounds once in fp32.
__device__ __forceinline__ float add_bf16_lo(uint32_t p, float c) {
    float d;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, lo, %2;\n\t}" : "=f"(d) : "r"(p), "f"(c));
    return d;
}
__device__ __forceinline__ float add_bf16_hi(uint32_t p, float c) {
    float d;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, hi, %2;\n\t}" : "=f"(d) : "r"(p), "f"(c));
    return d;
}
// bf16x2( max(lo, 0), max(hi, 0) ): ReLU fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// element-wise  h > 0 ? v : 0  on bf16 pairs (HSET2 + HMUL2): the ReLU-derivative mask without unpacking
__device__ __forceinline__ uint32_t mask_pos_bf16x2(uint32_t v, uint32_t h) {
    uint32_t m, d;
    asm("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(h), "r"(0u));
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(m));
    return d;
}
// element-wise a + b of two bf16 pairs, exact sum rounded once to bf16
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

}  // namespace tc5
