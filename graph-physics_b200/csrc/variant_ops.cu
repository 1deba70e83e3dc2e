// variant_ops.cu -- row-wise kernels of the variant flags of the path: what GraphNetBlock / Attention do
// around their dense layers when use_silu_activation, use_gated_mlp, use_gate / use_gated_attention or use_rope(_embeddings)
// are switched on (graphphysics/models/layers.py:213-249, 410-491, 637-697, 989-1149).  The dense layers themselves are
// gp_gemm; gathers and segment sums reuse gp_halo_pack / gp_segsum_gather.  fp32 arithmetic, no atomics.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "../../include/gp_b200.h"

namespace {
constexpr int kTB = 256;
inline unsigned nblk(long long n) { return (unsigned)((n + kTB - 1) / kTB); }

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + __expf(-z)); }
// kind: 1 ReLU, 2 SiLU, 3 GELU (exact, erf)
__device__ __forceinline__ float act_f(float z, int kind) {
    if (kind == 1) return fmaxf(z, 0.f);
    if (kind == 2) return z * sigmoidf_(z);
    return 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
}
__device__ __forceinline__ float act_df(float z, int kind) {
    if (kind == 1) return z > 0.f ? 1.f : 0.f;
    if (kind == 2) {
        const float s = sigmoidf_(z);
        return s * (1.f + z * (1.f - s));
    }
    return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.39894228040143268f * __expf(-0.5f * z * z);
}

__global__ void act_fwd_kernel(const float* __restrict__ z, long long n, int kind, __nv_bfloat16* __restrict__ ob, float* __restrict__ of) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h = act_f(z[i], kind);
    if (ob) ob[i] = __float2bfloat16_rn(h);
    else of[i] = h;
}
__global__ void act_bwd_kernel(const float* __restrict__ z, long long n, int kind, float* __restrict__ d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] *= act_df(z[i], kind);
}
// g = act(a1) * a2 ; da1 = dg * a2 * act'(a1), da2 = dg * act(a1)   (a1, a2: row stride ld, `cols` valid columns)
__global__ void glu_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int ld, long long rows, int cols, int kind,
                               __nv_bfloat16* __restrict__ ob, float* __restrict__ of) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = act_f(a1[r * ld + c], kind) * a2[r * ld + c];
    if (ob) ob[i] = __float2bfloat16_rn(v);
    else of[i] = v;
}
__global__ void glu_bwd_kernel(const float* __restrict__ a1, const float* __restrict__ a2, int ld, const float* __restrict__ dg, long long rows,
                               int cols, int kind, float* __restrict__ da1, float* __restrict__ da2, int ld_d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float x1 = a1[r * ld + c], d = dg[i];
    da1[r * ld_d + c] = d * a2[r * ld + c] * act_df(x1, kind);
    da2[r * ld_d + c] = d * act_f(x1, kind);
}
// out = v * sigmoid(logits) ; dlogits = dout * v * s * (1 - s), dv = dout * s
__global__ void sigmoid_mul_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ v, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[i] * sigmoidf_(logits[i]);
}
__global__ void sigmoid_mul_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ v, const float* __restrict__ dout, long long n,
                                       float* __restrict__ dlogits, float* __restrict__ dv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = sigmoidf_(logits[i]), d = dout[i];
    dlogits[i] = d * v[i] * s * (1.f - s);
    dv[i] = d * s;
}
// logits[r, c] += phi[r] * gate_pos[c]
__global__ void add_outer_kernel(float* __restrict__ logits, const float* __restrict__ phi, const float* __restrict__ gp_, long long rows, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    logits[i] += phi[r] * gp_[i - r * cols];
}

// Relative RoPE on gathered sender rows (layers.py:1104-1149): channel pair (2p, 2p+1) of axis a -- columns a*2*pc + 2p, +1 --
// is rotated by sign * (pos[src, a] - pos[dst, a]) * base^(-p / pc); columns past axes*2*pc are copied.
__global__ void rope_rel_kernel(const float* __restrict__ x, const float* __restrict__ pos, int ld_pos, const int32_t* __restrict__ src,
                                const int32_t* __restrict__ dst, long long E, int H, int axes, int pc, float base, float sign,
                                float* __restrict__ out) {
    const int half = H / 2 + (H & 1);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * half) return;
    const long long e = i / half;
    const int pr = (int)(i - e * half);
    const int c = 2 * pr;
    const float* xr = x + e * H;
    float* o = out + e * H;
    if (c + 1 < axes * 2 * pc) {
        const int a = c / (2 * pc), p = (c - a * 2 * pc) >> 1;
        const float delta = pos[(size_t)src[e] * ld_pos + a] - pos[(size_t)dst[e] * ld_pos + a];
        const float theta = sign * delta * powf(base, -(float)p / fmaxf((float)pc, 1.f));
        float s, co;
        sincosf(theta, &s, &co);
        const float ev = xr[c], od = xr[c + 1];
        o[c] = ev * co - od * s;
        o[c + 1] = ev * s + od * co;
    } else {
        o[c] = xr[c];
        if (c + 1 < H) o[c + 1] = xr[c + 1];
    }
}
// RoPE on per-node q / k in the (N, D, heads) layout (layers.py:420-491): for axis a, frequency i and head h the channels
// d = a*2m + 2i and d + 1 (flat index d*heads + h) rotate by sign * pos[n, a] * inv_freq[i]; in place.
__global__ void rope_nodes_kernel(float* __restrict__ t, const float* __restrict__ pos, int ld_pos, long long N, int D, int heads, int pd, int m,
                                  float log_base, float sign) {
    const long long per = (long long)pd * m * heads;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * per) return;
    const long long n = i / per;
    int r = (int)(i - n * per);
    const int a = r / (m * heads);
    r -= a * m * heads;
    const int f = r / heads, h = r - f * heads;
    const float inv = __expf(-(float)f * (log_base / fmaxf((float)m, 1.f)));
    float s, co;
    sincosf(sign * pos[n * ld_pos + a] * inv, &s, &co);
    float* p0 = t + n * (long long)D * heads + (long long)(a * 2 * m + 2 * f) * heads + h;
    const float ev = p0[0], od = p0[heads];
    p0[0] = ev * co - od * s;
    p0[heads] = ev * s + od * co;
}
// out[r] = [a[r] | b[r] | c[r]]  (fp32; b / c may be absent)
__global__ void concat_kernel(const float* __restrict__ a, int wa, const float* __restrict__ b, int wb, const float* __restrict__ c, int wc,
                              long long rows, float* __restrict__ out) {
    const int W = wa + wb + wc;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * W) return;
    const long long r = i / W;
    const int col = (int)(i - r * W);
    out[i] = col < wa ? a[r * wa + col] : (col < wa + wb ? b[r * wb + col - wa] : c[r * wc + col - wa - wb]);
}
// a[r] += d[r, 0:wa], ... the transpose of concat: three strided row copies with accumulation flags
__global__ void split_add_kernel(const float* __restrict__ d, int W, int col0, int w, long long rows, float* __restrict__ out, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * w) return;
    const long long r = i / w;
    const int c = (int)(i - r * w);
    const float v = d[r * W + col0 + c];
    out[i] = accumulate ? out[i] + v : v;
}
// out[n, :] = sum over p in [rowptr[n], rowptr[n+1]) of src[perm ? perm[p] : p, :], ascending p (fixed order, no atomics);
// one thread per (segment, column)
__global__ void segsum_f32_kernel(const float* __restrict__ src, int ld, const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
                                  long long num_segments, int H, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_segments * H) return;
    const long long n = i / H;
    const int c = (int)(i - n * H);
    float a = 0.f;
    for (int p = rowptr[n]; p < rowptr[n + 1]; ++p) a += src[(size_t)(perm ? perm[p] : p) * ld + c];
    out[i] = a;
}
}  // namespace

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int gp_segsum_rows_f32(const float* src, int32_t ld, const int32_t* perm, const int32_t* rowptr, int64_t num_segments, int32_t hidden,
                                  float* out, void* stream) {
    if (num_segments <= 0 || hidden <= 0) return 0;
    GP_REQUIRE(rowptr && out, "gp_segsum_rows_f32: null pointer");        // (src is NULL when there are no rows to add: all sums are zero)
    segsum_f32_kernel<<<nblk(num_segments * hidden), kTB, 0, ST(stream)>>>(src, ld, perm, rowptr, num_segments, hidden, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_act_fwd(const float* z, int64_t n, int32_t kind, gp_bf16* out_bf16, float* out_f32, void* stream) {
    if (n <= 0) return 0;
    GP_REQUIRE(z && kind >= 1 && kind <= 3 && ((out_bf16 != nullptr) != (out_f32 != nullptr)), "gp_act_fwd: bad arguments (kind 1 relu, 2 silu, 3 gelu)");
    act_fwd_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(z, n, kind, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_act_bwd(const float* z, int64_t n, int32_t kind, float* d, void* stream) {
    if (n <= 0) return 0;
    GP_REQUIRE(z && d && kind >= 1 && kind <= 3, "gp_act_bwd: bad arguments");
    act_bwd_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(z, n, kind, d);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_glu_fwd(const float* a1, const float* a2, int32_t ld, int64_t rows, int32_t cols, int32_t kind, gp_bf16* out_bf16, float* out_f32,
                          void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    GP_REQUIRE(a1 && a2 && (kind == 2 || kind == 3) && ((out_bf16 != nullptr) != (out_f32 != nullptr)), "gp_glu_fwd: bad arguments (kind 2 silu, 3 gelu)");
    glu_fwd_kernel<<<nblk(rows * cols), kTB, 0, ST(stream)>>>(a1, a2, ld, rows, cols, kind, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_glu_bwd(const float* a1, const float* a2, int32_t ld, const float* dg, int64_t rows, int32_t cols, int32_t kind, float* da1,
                          float* da2, int32_t ld_d, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    GP_REQUIRE(a1 && a2 && dg && da1 && da2 && (kind == 2 || kind == 3), "gp_glu_bwd: bad arguments");
    glu_bwd_kernel<<<nblk(rows * cols), kTB, 0, ST(stream)>>>(a1, a2, ld, dg, rows, cols, kind, da1, da2, ld_d);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_sigmoid_mul_fwd(const float* logits, const float* v, int64_t n, float* out, void* stream) {
    if (n <= 0) return 0;
    sigmoid_mul_fwd_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(logits, v, n, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_sigmoid_mul_bwd(const float* logits, const float* v, const float* dout, int64_t n, float* dlogits, float* dv, void* stream) {
    if (n <= 0) return 0;
    sigmoid_mul_bwd_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(logits, v, dout, n, dlogits, dv);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_add_outer(float* logits, const float* phi, const float* vec, int64_t rows, int32_t cols, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    add_outer_kernel<<<nblk(rows * cols), kTB, 0, ST(stream)>>>(logits, phi, vec, rows, cols);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_rope_rel(const float* x, const float* pos, int32_t ld_pos, const int32_t* src, const int32_t* dst, int64_t num_edges, int32_t hidden,
                           int32_t axes, int32_t pair_count, float base, int32_t inverse, float* out, void* stream) {
    if (num_edges <= 0) return 0;
    GP_REQUIRE(x && pos && src && dst && out && (axes == 2 || axes == 3) && axes * 2 * pair_count <= hidden, "gp_rope_rel: bad arguments");
    const long long n = num_edges * (long long)(hidden / 2 + (hidden & 1));
    rope_rel_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(x, pos, ld_pos, src, dst, num_edges, hidden, axes, pair_count, base, inverse ? -1.f : 1.f, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_rope_nodes(float* t, const float* pos, int32_t ld_pos, int64_t num_nodes, int32_t head_dim, int32_t num_heads, int32_t pos_dim,
                             int32_t m, float base, int32_t inverse, void* stream) {
    if (num_nodes <= 0 || m <= 0) return 0;
    GP_REQUIRE(t && pos && pos_dim >= 1 && pos_dim * 2 * m <= head_dim, "gp_rope_nodes: bad arguments");
    const long long n = num_nodes * (long long)pos_dim * m * num_heads;
    rope_nodes_kernel<<<nblk(n), kTB, 0, ST(stream)>>>(t, pos, ld_pos, num_nodes, head_dim, num_heads, pos_dim, m, logf(base), inverse ? -1.f : 1.f);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_concat_rows(const float* a, int32_t wa, const float* b, int32_t wb, const float* c, int32_t wc, int64_t rows, float* out, void* stream) {
    if (rows <= 0) return 0;
    GP_REQUIRE(a && out && wa > 0 && (wb == 0 || b) && (wc == 0 || c), "gp_concat_rows: bad arguments");
    concat_kernel<<<nblk(rows * (wa + wb + wc)), kTB, 0, ST(stream)>>>(a, wa, b, wb, c, wc, rows, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_split_cols(const float* d, int32_t width, int32_t col0, int32_t w, int64_t rows, float* out, int32_t accumulate, void* stream) {
    if (rows <= 0 || w <= 0) return 0;
    GP_REQUIRE(d && out && col0 >= 0 && col0 + w <= width, "gp_split_cols: bad arguments");
    split_add_kernel<<<nblk(rows * w), kTB, 0, ST(stream)>>>(d, width, col0, w, rows, out, accumulate);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
