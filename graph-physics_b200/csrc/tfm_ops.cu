// tfm_ops.cu -- row-wise kernels of the graph-Transformer block around the tensor-core GEMMs (gp_gemm) and the CSR
// attention kernels: RMSNorm (single or the double norm in front of the gated MLP) forward / backward, the GELU gate
// of GatedMLP forward / backward, and column sums (bias gradients).  Reference: graphphysics/models/layers.py:104-129
// (RMSNorm), 213-249 (GatedMLP: GELU(W1 x) * (W2 x), exact erf GELU), 766-819 (Transformer.forward).
// One warp per row (H <= 128 values, H/32 per lane), fp32 arithmetic, no atomics: per-block partial sums are reduced in
// fixed order by gp_reduce_partials.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "../../include/gp_b200.h"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = s1 * x / (||x||/sqrt(h) + eps);  if s2: y = s2 * y / (||y||/sqrt(h) + eps)
template <int VPT>
__global__ void rmsnorm_fwd_kernel(const float* __restrict__ x, int ldx, int rows, const float* __restrict__ s1,
                                   const float* __restrict__ s2, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                   int ld_out) {
    constexpr int H = 32 * VPT;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float v[VPT];
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        v[t] = x[(size_t)r * ldx + lane + 32 * t];
        ss = fmaf(v[t], v[t], ss);
    }
    float inv = 1.f / (sqrtf(warp_sum(ss) * (1.f / H)) + 1e-8f);
    ss = 0.f;
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        v[t] = s1[lane + 32 * t] * (v[t] * inv);
        ss = fmaf(v[t], v[t], ss);
    }
    if (s2) {
        inv = 1.f / (sqrtf(warp_sum(ss) * (1.f / H)) + 1e-8f);
#pragma unroll
        for (int t = 0; t < VPT; ++t) v[t] = s2[lane + 32 * t] * (v[t] * inv);
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        if (out_bf16) out_bf16[(size_t)r * ld_out + lane + 32 * t] = __float2bfloat16_rn(v[t]);
        else out_f32[(size_t)r * ld_out + lane + 32 * t] = v[t];
    }
}

// backward of one norm on register rows: given x, s and g = dL/dy returns dL/dx in g; accumulates dscale into ds
template <int VPT>
__device__ __forceinline__ void norm_bwd_row(const float (&x)[VPT], const float* __restrict__ s, float (&g)[VPT], float (&ds)[VPT], int lane) {
    constexpr int H = 32 * VPT;
    float ss = 0.f, dot = 0.f;
    float sc[VPT];
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        sc[t] = s[lane + 32 * t];
        ss = fmaf(x[t], x[t], ss);
        dot = fmaf(g[t] * sc[t], x[t], dot);
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float rms = sqrtf(ss * (1.f / H));
    const float inv = 1.f / (rms + 1e-8f);
    const float coef = rms > 0.f ? dot * inv * inv / (rms * H) : 0.f;
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        ds[t] += g[t] * (x[t] * inv);
        g[t] = fmaf(sc[t] * g[t], inv, -(coef * x[t]));
    }
}

// dx = [add +] d/dx [norm(norm(x; s1); s2)] . dy   (s2, add optional; add may alias dx);  per-block partial sums of dscale1 / dscale2
template <int VPT>
__global__ void rmsnorm_bwd_kernel(const float* __restrict__ x, int ldx, int rows, const float* __restrict__ s1,
                                   const float* __restrict__ s2, const float* __restrict__ dy, int ld_dy, float* dx,
                                   int ld_dx, const float* add, int ld_add, float* __restrict__ part1, float* __restrict__ part2) {
    constexpr int H = 32 * VPT;
    __shared__ float sh[2][8][H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float d1[VPT], d2[VPT];
#pragma unroll
    for (int t = 0; t < VPT; ++t) { d1[t] = 0.f; d2[t] = 0.f; }
    // a block walks its rows in ascending order, warp w takes rows w, w+8, ...: fixed summation order
    const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    for (int r = r_begin + warp; r < r_end; r += 8) {
        float xv[VPT], g[VPT];
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            xv[t] = x[(size_t)r * ldx + lane + 32 * t];
            g[t] = dy[(size_t)r * ld_dy + lane + 32 * t];
        }
        if (s2) {
            // recompute y1 = norm(x; s1), then back through the second and the first norm
            float ss = 0.f;
#pragma unroll
            for (int t = 0; t < VPT; ++t) ss = fmaf(xv[t], xv[t], ss);
            const float inv = 1.f / (sqrtf(warp_sum(ss) * (1.f / H)) + 1e-8f);
            float y1[VPT];
#pragma unroll
            for (int t = 0; t < VPT; ++t) y1[t] = s1[lane + 32 * t] * (xv[t] * inv);
            norm_bwd_row<VPT>(y1, s2, g, d2, lane);
        }
        norm_bwd_row<VPT>(xv, s1, g, d1, lane);
#pragma unroll
        for (int t = 0; t < VPT; ++t) {
            const float a = add ? add[(size_t)r * ld_add + lane + 32 * t] : 0.f;
            dx[(size_t)r * ld_dx + lane + 32 * t] = a + g[t];
        }
    }
#pragma unroll
    for (int t = 0; t < VPT; ++t) {
        sh[0][warp][lane + 32 * t] = d1[t];
        sh[1][warp][lane + 32 * t] = d2[t];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += sh[0][w][c]; b += sh[1][w][c]; }
        part1[(size_t)blockIdx.x * H + c] = a;
        if (part2) part2[(size_t)blockIdx.x * H + c] = b;
    }
}


// ---- any width (the 2H / 3H-wide norm in front of the gated edge / node MLP, layers.py:252-278): one norm, columns strided over the lanes
__global__ void rmsnorm_fwd_generic_kernel(const float* __restrict__ x, int ldx, int rows, int H, const float* __restrict__ s1,
                                           __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int ld_out) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + (size_t)r * ldx;
    float ss = 0.f;
    for (int c = lane; c < H; c += 32) ss = fmaf(xr[c], xr[c], ss);
    const float inv = 1.f / (sqrtf(warp_sum(ss) / (float)H) + 1e-8f);
    for (int c = lane; c < H; c += 32) {
        const float v = s1[c] * (xr[c] * inv);
        if (out_bf16) out_bf16[(size_t)r * ld_out + c] = __float2bfloat16_rn(v);
        else out_f32[(size_t)r * ld_out + c] = v;
    }
}
__global__ void rmsnorm_bwd_generic_kernel(const float* __restrict__ x, int ldx, int rows, int H, const float* __restrict__ s1,
                                           const float* __restrict__ dy, int ld_dy, float* dx, int ld_dx, const float* add, int ld_add,
                                           float* __restrict__ part1) {
    extern __shared__ float shg[];              // [8 warps][H]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* mine = shg + (size_t)warp * H;
    for (int c = lane; c < H; c += 32) mine[c] = 0.f;
    const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    for (int r = r_begin + warp; r < r_end; r += 8) {
        const float* xr = x + (size_t)r * ldx;
        const float* g = dy + (size_t)r * ld_dy;
        float ss = 0.f, dot = 0.f;
        for (int c = lane; c < H; c += 32) {
            ss = fmaf(xr[c], xr[c], ss);
            dot = fmaf(g[c] * s1[c], xr[c], dot);
        }
        ss = warp_sum(ss);
        dot = warp_sum(dot);
        const float rms = sqrtf(ss / (float)H);
        const float inv = 1.f / (rms + 1e-8f);
        const float coef = rms > 0.f ? dot * inv * inv / (rms * H) : 0.f;
        for (int c = lane; c < H; c += 32) {
            mine[c] += g[c] * (xr[c] * inv);
            const float a = add ? add[(size_t)r * ld_add + c] : 0.f;
            dx[(size_t)r * ld_dx + c] = a + fmaf(s1[c] * g[c], inv, -(coef * xr[c]));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += shg[(size_t)w * H + c];
        part1[(size_t)blockIdx.x * H + c] = a;
    }
}

__device__ __forceinline__ float gelu_exact(float a) { return 0.5f * a * (1.f + erff(a * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float a) {
    return 0.5f * (1.f + erff(a * 0.70710678118654752f)) + a * 0.39894228040143268f * __expf(-0.5f * a * a);
}
// g = gelu(a1) * a2  (elementwise over contiguous fp32 arrays)
__global__ void gelu_gate_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ a2, long long n, __nv_bfloat16* __restrict__ g,
                                     float* __restrict__ g32) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = gelu_exact(a1[i]) * a2[i];
    if (g) g[i] = __float2bfloat16_rn(v);
    else g32[i] = v;
}
// da1 = dg * a2 * gelu'(a1);  da2 = dg * gelu(a1)
__global__ void gelu_gate_bwd_kernel(const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ dg, long long n,
                                     float* __restrict__ da1, float* __restrict__ da2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x1 = a1[i], d = dg[i];
    da1[i] = d * a2[i] * gelu_grad(x1);
    da2[i] = d * gelu_exact(x1);
}
// in-place ReLU backward: d[i] = h[i] > 0 ? d[i] : 0
template <typename T>
__global__ void relu_bwd_kernel(float* __restrict__ d, const T* __restrict__ h, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(static_cast<float>(h[i]) > 0.f)) d[i] = 0.f;
}

// per-block partial column sums (block b takes a contiguous slice of rows, each thread one column, ascending rows)
__global__ void colsum_kernel(const float* __restrict__ src, int ld, int rows, int cols, int round_bf16, float* __restrict__ part) {
    const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float a = 0.f;
        for (int r = r_begin; r < r_end; ++r) {
            const float v = src[(size_t)r * ld + c];
            a += round_bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
        }
        part[(size_t)blockIdx.x * cols + c] = a;
    }
}
}  // namespace

extern "C" int gp_rmsnorm_fwd(const float* x, int32_t ldx, int32_t rows, int32_t hidden, const float* scale1, const float* scale2,
                              gp_bf16* out_bf16, float* out_f32, int32_t ld_out, void* stream) {
    if (rows <= 0) return 0;
    GP_REQUIRE(x && scale1 && ((out_bf16 != nullptr) != (out_f32 != nullptr)), "gp_rmsnorm_fwd: need x, scale1 and exactly one output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = 256, blocks = (int)(((size_t)rows * 32 + threads - 1) / threads);
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    switch (hidden) {
        case 32: rmsnorm_fwd_kernel<1><<<blocks, threads, 0, st>>>(x, ldx, rows, scale1, scale2, ob, out_f32, ld_out); break;
        case 64: rmsnorm_fwd_kernel<2><<<blocks, threads, 0, st>>>(x, ldx, rows, scale1, scale2, ob, out_f32, ld_out); break;
        case 128: rmsnorm_fwd_kernel<4><<<blocks, threads, 0, st>>>(x, ldx, rows, scale1, scale2, ob, out_f32, ld_out); break;
        default:
            GP_REQUIRE(scale2 == nullptr && hidden > 0 && hidden <= 1024, "gp_rmsnorm_fwd: the double norm needs hidden 32, 64 or 128 (got %d)", hidden);
            rmsnorm_fwd_generic_kernel<<<blocks, threads, 0, st>>>(x, ldx, rows, hidden, scale1, ob, out_f32, ld_out);
    }
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_rmsnorm_bwd_blocks(int32_t rows) {
    const int b = (rows + 63) / 64;      // 8 rows per warp: short dependent chains, every SM busy
    return b < 1 ? 1 : (b > 1184 ? 1184 : b);
}

extern "C" int gp_rmsnorm_bwd(const float* x, int32_t ldx, int32_t rows, int32_t hidden, const float* scale1, const float* scale2,
                              const float* dy, int32_t ld_dy, float* dx, int32_t ld_dx, const float* add, int32_t ld_add, float* part1,
                              float* part2, void* stream) {
    if (rows <= 0) return 0;
    GP_REQUIRE(x && scale1 && dy && dx && part1 && ((scale2 != nullptr) == (part2 != nullptr)), "gp_rmsnorm_bwd: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = gp_rmsnorm_bwd_blocks(rows);
    switch (hidden) {
        case 32: rmsnorm_bwd_kernel<1><<<blocks, 256, 0, st>>>(x, ldx, rows, scale1, scale2, dy, ld_dy, dx, ld_dx, add, ld_add, part1, part2); break;
        case 64: rmsnorm_bwd_kernel<2><<<blocks, 256, 0, st>>>(x, ldx, rows, scale1, scale2, dy, ld_dy, dx, ld_dx, add, ld_add, part1, part2); break;
        case 128: rmsnorm_bwd_kernel<4><<<blocks, 256, 0, st>>>(x, ldx, rows, scale1, scale2, dy, ld_dy, dx, ld_dx, add, ld_add, part1, part2); break;
        default:
            GP_REQUIRE(scale2 == nullptr && hidden > 0 && hidden <= 1024, "gp_rmsnorm_bwd: the double norm needs hidden 32, 64 or 128 (got %d)", hidden);
            rmsnorm_bwd_generic_kernel<<<blocks, 256, (size_t)8 * hidden * sizeof(float), st>>>(x, ldx, rows, hidden, scale1, dy, ld_dy, dx, ld_dx, add,
                                                                                             ld_add, part1);
    }
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_gelu_gate_fwd(const float* a1, const float* a2, int64_t n, gp_bf16* g_bf16, float* g_f32, void* stream) {
    if (n <= 0) return 0;
    GP_REQUIRE(a1 && a2 && ((g_bf16 != nullptr) != (g_f32 != nullptr)), "gp_gelu_gate_fwd: need a1, a2 and exactly one output");
    gelu_gate_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a1, a2, n, reinterpret_cast<__nv_bfloat16*>(g_bf16), g_f32);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_gelu_gate_bwd(const float* a1, const float* a2, const float* dg, int64_t n, float* da1, float* da2, void* stream) {
    if (n <= 0) return 0;
    gelu_gate_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a1, a2, dg, n, da1, da2);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_relu_bwd(float* d, const void* h, int32_t h_bf16, int64_t n, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h_bf16) relu_bwd_kernel<__nv_bfloat16><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, static_cast<const __nv_bfloat16*>(h), n);
    else relu_bwd_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, static_cast<const float*>(h), n);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int gp_colsum_blocks(int32_t rows) {
    const int b = (rows + 511) / 512;
    return b < 1 ? 1 : (b > 148 ? 148 : b);
}
extern "C" int gp_colsum(const float* src, int32_t ld, int32_t rows, int32_t cols, int32_t round_bf16, float* partials, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    colsum_kernel<<<gp_colsum_blocks(rows), 128, 0, static_cast<cudaStream_t>(stream)>>>(src, ld, rows, cols, round_bf16, partials);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
