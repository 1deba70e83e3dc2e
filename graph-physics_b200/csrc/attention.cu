// attention.cu -- adjacency-masked multi-head attention over a CSR graph (gp_csr_attention_*).
//
// Replaces the DGL-sparse path of graphphysics/models/layers.py:493-561 (bsddmm -> row softmax ->
// bspmm): for every stored (i = edge_index[0], j = edge_index[1]) and head h,
//     s_ij^h = (q_i^h . k_j^h) / sqrt(d),   a_i.^h = softmax_j(s_i.^h),   y_i^h = sum_j a_ij^h v_j^h
// with the reference's head layout: q.reshape(N, d, heads) puts the HEAD index innermost, channel
// c = d_idx * heads + h (layers.py:673-675).  No self loops are added; a row without entries gives 0.
//
// One warp per row, online softmax, lane l owns channels [l*VPT, (l+1)*VPT).  Channels of one head
// are strided across lanes, so the per-head dot product is a butterfly all-reduce over the lanes
// that hold the same head.  fp32 throughout.  The backward is two gather passes (rows, then
// columns) -- no atomics, bit-reproducible.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/gp_b200.h"

namespace {

template <int VPT, int HH>
struct HeadReduce {
    static constexpr int LOCAL = VPT > HH ? VPT / HH : 1;    // elements of one head inside a lane
    static constexpr int S = HH > VPT ? HH / VPT : 1;        // lanes l, l+S, l+2S.. hold the same heads
    // in: p[t]; out: r[t] = sum of p over every channel with the same head as element t (all lanes)
    __device__ static void run(const float (&p)[VPT], float (&r)[VPT]) {
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            float acc = 0.f;
            if constexpr (VPT > HH) {
#pragma unroll
                for (int u = t % HH; u < VPT; u += HH) acc += p[u];
            } else {
                acc = p[t];
            }
            r[t] = acc;
        }
#pragma unroll
        for (int o = S; o < 32; o <<= 1) {
#pragma unroll
            for (int t = 0; t < VPT; ++t) r[t] += __shfl_xor_sync(0xffffffffu, r[t], o);
        }
    }
};

// bf16 rows (q, k, v, y in the Transformer path: half the gathered bytes, SURVEY §8d E*(4H+8)) converted on load
template <int VPT>
__device__ __forceinline__ void load_row(const __nv_bfloat16* __restrict__ base, size_t row, int H, int lane, float (&out)[VPT]) {
    const __nv_bfloat16* p = base + row * H + lane * VPT;
    if constexpr (VPT == 4) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        out[0] = __uint_as_float(v.x << 16); out[1] = __uint_as_float(v.x & 0xFFFF0000u);
        out[2] = __uint_as_float(v.y << 16); out[3] = __uint_as_float(v.y & 0xFFFF0000u);
    } else if constexpr (VPT == 2) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
        out[0] = __uint_as_float(v << 16); out[1] = __uint_as_float(v & 0xFFFF0000u);
    } else {
        out[0] = __bfloat162float(p[0]);
    }
}
__device__ __forceinline__ void store_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_elem(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int VPT>
__device__ __forceinline__ void load_row(const float* __restrict__ base, size_t row, int H, int lane, float (&out)[VPT]) {
    const float* p = base + row * H + lane * VPT;
    if constexpr (VPT == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else if constexpr (VPT == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p));
        out[0] = v.x; out[1] = v.y;
    } else {
        out[0] = __ldg(p);
    }
}

template <int VPT, int HH, typename T>
__global__ void attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, float scale,
                                T* __restrict__ y, float* __restrict__ y32, float* __restrict__ lse, int ldq) {
    constexpr int H = 32 * VPT;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    float qv[VPT], m[VPT], s[VPT], acc[VPT];
    load_row<VPT>(q, i, ldq, lane, qv);
#pragma unroll
    for (int t = 0; t < VPT; ++t) { qv[t] *= scale; m[t] = -INFINITY; s[t] = 0.f; acc[t] = 0.f; }
    const int b = rowptr[i], e = rowptr[i + 1];
    for (int p = b; p < e; ++p) {
        const int j = __ldg(col + p);
        float kv[VPT], vv[VPT], pr[VPT], dot[VPT];
        load_row<VPT>(k, j, ldq, lane, kv);
        load_row<VPT>(v, j, ldq, lane, vv);
#pragma unroll
        for (int t = 0; t < VPT; ++t) pr[t] = qv[t] * kv[t];
        HeadReduce<VPT, HH>::run(pr, dot);
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            const float mn = fmaxf(m[t], dot[t]);
            const float corr = __expf(m[t] - mn), w = __expf(dot[t] - mn);
            s[t] = s[t] * corr + w;
            acc[t] = acc[t] * corr + w * vv[t];
            m[t] = mn;
        }
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        const int c = lane * VPT + t;
        const float yv = (e > b) ? acc[t] / s[t] : 0.f;
        store_elem(y + (size_t)i * H + c, yv);
        if (y32) y32[(size_t)i * H + c] = yv;
        if (c < HH) lse[(size_t)i * HH + c] = (e > b) ? m[t] + __logf(s[t]) : 0.f;   // channel c < HH has head c
    }
}

// Backward pass 1 (rows): dq_i, and per stored entry the softmax weight a and ds = a * (dy.v - dy.y),
// written at the entry's position in the COLUMN-sorted list (pos[p]) for pass 2.
template <int VPT, int HH, typename T>
__global__ void attn_bwd_rows_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                     const T* __restrict__ y, const float* __restrict__ y32, const float* __restrict__ dy,
                                     const float* __restrict__ lse, const int32_t* __restrict__ rowptr,
                                     const int32_t* __restrict__ col, const int32_t* __restrict__ pos, int n, float scale,
                                     float* __restrict__ dq, float* __restrict__ ea, float* __restrict__ eds, int ldq, int ldg) {
    constexpr int H = 32 * VPT;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    float qv[VPT], yv[VPT], dyv[VPT], l[VPT], pr[VPT], D[VPT], dqa[VPT];
    load_row<VPT>(q, i, ldq, lane, qv);
    if (y32) load_row<VPT>(y32, i, H, lane, yv);
    else load_row<VPT>(y, i, H, lane, yv);
    load_row<VPT>(dy, i, H, lane, dyv);
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        qv[t] *= scale;
        l[t] = __ldg(lse + (size_t)i * HH + (lane * VPT + t) % HH);
        pr[t] = dyv[t] * yv[t];
        dqa[t] = 0.f;
    }
    HeadReduce<VPT, HH>::run(pr, D);
    const int b = rowptr[i], e = rowptr[i + 1];
    for (int p = b; p < e; ++p) {
        const int j = __ldg(col + p);
        float kv[VPT], vv[VPT], dot[VPT], dp[VPT];
        load_row<VPT>(k, j, ldq, lane, kv);
        load_row<VPT>(v, j, ldq, lane, vv);
#pragma unroll
        for (int t = 0; t < VPT; ++t) pr[t] = qv[t] * kv[t];
        HeadReduce<VPT, HH>::run(pr, dot);
#pragma unroll
        for (int t = 0; t < VPT; ++t) pr[t] = dyv[t] * vv[t];
        HeadReduce<VPT, HH>::run(pr, dp);
        const size_t o = (size_t)__ldg(pos + p) * HH;
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            const float a = __expf(dot[t] - l[t]);
            const float ds = a * (dp[t] - D[t]);
            dqa[t] = fmaf(ds * scale, kv[t], dqa[t]);
            const int c = lane * VPT + t;
            if (c < HH) { ea[o + c] = a; eds[o + c] = ds; }
        }
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) dq[(size_t)i * ldg + lane * VPT + t] = dqa[t];
}

// Backward pass 2 (columns): dk_j = sum_i ds_ij q_i / sqrt(d),  dv_j = sum_i a_ij dy_i over the entries of
// column j (contiguous in the column-sorted list; `row` holds their row index).
template <int VPT, int HH, typename T>
__global__ void attn_bwd_cols_kernel(const T* __restrict__ q, const float* __restrict__ dy,
                                     const float* __restrict__ ea, const float* __restrict__ eds,
                                     const int32_t* __restrict__ colptr, const int32_t* __restrict__ row, int n,
                                     float scale, float* __restrict__ dk, float* __restrict__ dv, int ldq, int ldg) {
    constexpr int H = 32 * VPT;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= n) return;
    float dka[VPT], dva[VPT];
#pragma unroll
    for (int t = 0; t < VPT; ++t) { dka[t] = 0.f; dva[t] = 0.f; }
    const int b = colptr[j], e = colptr[j + 1];
    for (int p = b; p < e; ++p) {
        const int i = __ldg(row + p);
        float qv[VPT], dyv[VPT];
        load_row<VPT>(q, i, ldq, lane, qv);
        load_row<VPT>(dy, i, H, lane, dyv);
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            const int h = (lane * VPT + t) % HH;
            const float a = __ldg(ea + (size_t)p * HH + h), ds = __ldg(eds + (size_t)p * HH + h);
            dka[t] = fmaf(ds * scale, qv[t], dka[t]);
            dva[t] = fmaf(a, dyv[t], dva[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        dk[(size_t)j * ldg + lane * VPT + t] = dka[t];
        dv[(size_t)j * ldg + lane * VPT + t] = dva[t];
    }
}

template <int VPT, int HH, typename T>
void launch_typed(int which, const gp_attention_args& a, cudaStream_t st) {
    const int threads = 256, blocks = (int)(((size_t)a.n * 32 + threads - 1) / threads);
    const float scale = 1.f / sqrtf((float)(32 * VPT / HH));
    const T *q = reinterpret_cast<const T*>(a.q), *k = reinterpret_cast<const T*>(a.k), *v = reinterpret_cast<const T*>(a.v);
    T* y = reinterpret_cast<T*>(a.y);
    const int ldq = a.ld_qkv > 0 ? a.ld_qkv : 32 * VPT, ldg = a.ld_dqkv > 0 ? a.ld_dqkv : 32 * VPT;
    if (which == 0)
        attn_fwd_kernel<VPT, HH, T><<<blocks, threads, 0, st>>>(q, k, v, a.rowptr, a.col, a.n, scale, y, a.y_f32, a.lse, ldq);
    else if (which == 1)
        attn_bwd_rows_kernel<VPT, HH, T><<<blocks, threads, 0, st>>>(q, k, v, y, a.y_f32, a.dy, a.lse, a.rowptr, a.col, a.pos, a.n, scale, a.dq,
                                                                   a.edge_a, a.edge_ds, ldq, ldg);
    else
        attn_bwd_cols_kernel<VPT, HH, T><<<blocks, threads, 0, st>>>(q, a.dy, a.edge_a, a.edge_ds, a.colptr, a.row, a.n, scale, a.dk, a.dv, ldq,
                                                                   ldg);
}
template <int VPT, int HH>
void launch_all(int which, const gp_attention_args& a, cudaStream_t st) {
    if (a.io_bf16) launch_typed<VPT, HH, __nv_bfloat16>(which, a, st);
    else launch_typed<VPT, HH, float>(which, a, st);
}

template <int VPT>
int dispatch_heads(int which, const gp_attention_args& a, cudaStream_t st) {
    switch (a.num_heads) {
        case 1: launch_all<VPT, 1>(which, a, st); break;
        case 2: launch_all<VPT, 2>(which, a, st); break;
        case 4: launch_all<VPT, 4>(which, a, st); break;
        case 8: launch_all<VPT, 8>(which, a, st); break;
        case 16: launch_all<VPT, 16>(which, a, st); break;
        default: gp::set_error("gp_csr_attention: num_heads must be 1, 2, 4, 8 or 16 (got %d)", a.num_heads); return -1;
    }
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int dispatch(int which, const gp_attention_args* args, void* stream) {
    GP_REQUIRE(args != nullptr, "gp_csr_attention: null args");
    const gp_attention_args& a = *args;
    if (a.n <= 0) return 0;
    GP_REQUIRE(a.hidden % a.num_heads == 0, "gp_csr_attention: hidden must be divisible by num_heads");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (a.hidden) {
        case 32: return dispatch_heads<1>(which, a, st);
        case 64: return dispatch_heads<2>(which, a, st);
        case 128: return dispatch_heads<4>(which, a, st);
        default: gp::set_error("gp_csr_attention: hidden must be 32, 64 or 128 (got %d)", a.hidden); return -1;
    }
}
}  // namespace

extern "C" int gp_csr_attention_fwd(const gp_attention_args* args, void* stream) {
    // (col / row / pos / edge_* are empty -- possibly NULL -- for an adjacency without entries: no row loop then runs)
    GP_REQUIRE(args && args->q && args->k && args->v && args->y && args->lse && args->rowptr, "gp_csr_attention_fwd: null pointer");
    return dispatch(0, args, stream);
}
extern "C" int gp_csr_attention_bwd(const gp_attention_args* args, void* stream) {
    GP_REQUIRE(args && args->dy && args->dq && args->dk && args->dv && args->rowptr && args->colptr, "gp_csr_attention_bwd: null pointer");
    const int rc = dispatch(1, args, stream);
    return rc ? rc : dispatch(2, args, stream);
}
