// train_ops.cu -- the small, bandwidth-trivial pieces of the training step, kept on the device and
// deterministic: masked L2 loss + gradient, global gradient norm, clip + AdamW on the flat
// parameter buffer, and the fp32 -> packed-bf16 weight refresh.
#include "common.cuh"
#include "../../include/gp_b200.h"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
    // fixed-shape tree: warp shuffle then one warp over the per-warp sums (deterministic)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sh[0] = v;
    __syncthreads();
    return sh[0];
}

// L2Loss.forward (graphphysics/utils/loss.py:45-75): mean over masked rows x D of (out-target)^2,
// and its gradient.  Three small launches: per-block partial sums, fixed-order final sum, gradient.
constexpr int kMseBlocks = 128;
__global__ void masked_mse_partial_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                          const uint8_t* __restrict__ mask, int n, int d, float* __restrict__ ws) {
    __shared__ float sh[32];
    float se = 0.f, cnt = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (mask[r]) {
            cnt += 1.f;
            for (int c = 0; c < d; ++c) {
                const float e = out[(size_t)r * d + c] - tgt[(size_t)r * d + c];
                se = fmaf(e, e, se);
            }
        }
    }
    se = block_sum(se, sh);
    cnt = block_sum(cnt, sh);
    if (threadIdx.x == 0) {
        ws[blockIdx.x] = se;
        ws[kMseBlocks + blockIdx.x] = cnt;
    }
}
__global__ void masked_mse_final_kernel(float* __restrict__ ws, int d, float* __restrict__ loss) {
    __shared__ float sh[32];
    float se = threadIdx.x < kMseBlocks ? ws[threadIdx.x] : 0.f;
    float cnt = threadIdx.x < kMseBlocks ? ws[kMseBlocks + threadIdx.x] : 0.f;
    se = block_sum(se, sh);
    cnt = block_sum(cnt, sh);
    if (threadIdx.x == 0) {
        const float denom = cnt * d;
        loss[0] = se / denom;          // 0/0 -> NaN like torch.mean of an empty selection
        ws[2 * kMseBlocks] = denom;
    }
}
__global__ void masked_mse_grad_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                       const uint8_t* __restrict__ mask, int n, int d, const float* __restrict__ ws,
                                       float* __restrict__ grad, float grad_scale) {
    const float k = 2.f * grad_scale / ws[2 * kMseBlocks];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * d; i += gridDim.x * blockDim.x)
        grad[i] = mask[i / d] ? k * (out[i] - tgt[i]) : 0.f;
}

__global__ void sqnorm_partial_kernel(const float* __restrict__ g, size_t n, float* __restrict__ partial) {
    __shared__ float sh[32];
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        s = fmaf(g[i], g[i], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void sqnorm_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    __shared__ float sh[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = s;
}

// clip_grad_norm_(max_norm) followed by torch.optim.AdamW (lightning_module.py:494-511,
// train.py:288 gradient_clip_val=1.0).
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2, float max_norm, const float* __restrict__ sqnorm) {
    float coef = 1.f;
    if (max_norm > 0.f) {
        const float nrm = sqrtf(sqnorm[0]);
        coef = fminf(max_norm / (nrm + 1e-6f), 1.f);
    }
    const float step = lr / bc1, rs = rsqrtf(bc2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        pi -= step * mi / (sqrtf(vi) * rs + eps);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// Same update with the step-dependent scalars taken from device memory, so the whole training step
// can be captured once in a CUDA graph and replayed: state[0] = number of optimizer steps done so
// far.  The learning rate is base_lr * CosineWarmupScheduler.get_lr_factor(last_epoch = state[0])
// (graphphysics/utils/scheduler.py:51-67).
__global__ void adamw_sched_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                   float* __restrict__ v, size_t n, const int* __restrict__ state, float base_lr,
                                   int warmup, int max_iters, float min_factor, float b1, float b2, float eps, float wd,
                                   float max_norm, const float* __restrict__ sqnorm) {
    const int t = state[0] + 1;
    float f = 0.5f * (1.f + cospif((float)t / (float)max_iters));
    if (t <= warmup) f *= (float)t / (float)warmup;
    const float lr = base_lr * fmaxf(f, min_factor);
    const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(max_norm / (sqrtf(sqnorm[0]) + 1e-6f), 1.f);
    const float step = lr / bc1, rs = rsqrtf(bc2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        pi -= step * mi / (sqrtf(vi) * rs + eps);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
__global__ void advance_step_kernel(int* state) { state[0] += 1; }

// fp32 master weights -> packed bf16 operands.  One table entry per matrix; blockIdx.y = entry.
__global__ void pack_weights_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ packed,
                                    const gp_pack_entry* __restrict__ table) {
    const gp_pack_entry e = table[blockIdx.y];
    const int total = e.n * e.k;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / e.k, c = i - r * e.k;
        packed[(size_t)e.dst_off + (size_t)(e.dst_row0 + r) * e.ld_dst + e.dst_col0 + c] =
            __float2bfloat16(params[(size_t)e.src_off + (size_t)r * e.ld_src + e.src_col0 + c]);
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = __float2bfloat16(s[i]);
}
}  // namespace

extern "C" int gp_masked_mse(const float* out, const float* target, const uint8_t* mask, int32_t n, int32_t d,
                             float* loss, float* grad, float grad_scale, float* workspace, void* stream) {
    GP_REQUIRE(out && target && mask && loss && n > 0 && d > 0, "gp_masked_mse: bad arguments");
    GP_REQUIRE(workspace != nullptr, "gp_masked_mse: workspace of 260 floats required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    masked_mse_partial_kernel<<<kMseBlocks, 256, 0, st>>>(out, target, mask, n, d, workspace);
    masked_mse_final_kernel<<<1, 128, 0, st>>>(workspace, d, loss);
    if (grad) masked_mse_grad_kernel<<<256, 256, 0, st>>>(out, target, mask, n, d, workspace, grad, grad_scale);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_sqnorm(const float* g, int64_t n, float* workspace, float* out, void* stream) {
    GP_REQUIRE(g && workspace && out && n > 0, "gp_sqnorm: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = 256;
    sqnorm_partial_kernel<<<blocks, 256, 0, st>>>(g, (size_t)n, workspace);
    sqnorm_final_kernel<<<1, 256, 0, st>>>(workspace, blocks, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int32_t step, float max_norm,
                        const float* sqnorm, void* stream) {
    GP_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "gp_adamw: bad arguments");
    GP_REQUIRE(max_norm <= 0.f || sqnorm != nullptr, "gp_adamw: clipping needs the squared gradient norm");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    int blocks = (int)((n + 255) / 256);
    if (blocks > gp::sm_count() * 8) blocks = gp::sm_count() * 8;
    adamw_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, (size_t)n, lr,
                                                                         beta1, beta2, eps, weight_decay, bc1, bc2,
                                                                         max_norm, sqnorm);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_adamw_sched(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              int32_t* state, float base_lr, int32_t warmup, int32_t max_iters, float min_lr_factor,
                              float beta1, float beta2, float eps, float weight_decay, float max_norm,
                              const float* sqnorm, void* stream) {
    GP_REQUIRE(params && grads && exp_avg && exp_avg_sq && state && n > 0 && max_iters > 0, "gp_adamw_sched: bad arguments");
    GP_REQUIRE(max_norm <= 0.f || sqnorm != nullptr, "gp_adamw_sched: clipping needs the squared gradient norm");
    int blocks = (int)((n + 255) / 256);
    if (blocks > gp::sm_count() * 8) blocks = gp::sm_count() * 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    adamw_sched_kernel<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (size_t)n, state, base_lr, warmup,
                                               max_iters, min_lr_factor, beta1, beta2, eps, weight_decay, max_norm, sqnorm);
    advance_step_kernel<<<1, 1, 0, st>>>(state);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_pack_weights(const float* params, gp_bf16* packed, const gp_pack_entry* table, int32_t n_entries,
                               void* stream) {
    if (n_entries <= 0) return 0;
    GP_REQUIRE(params && packed && table, "gp_pack_weights: null pointer");
    dim3 grid(16, n_entries);
    pack_weights_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        params, reinterpret_cast<__nv_bfloat16*>(packed), table);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_cast_bf16(const float* src, gp_bf16* dst, int64_t n, void* stream) {
    if (n <= 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > gp::sm_count() * 8) blocks = gp::sm_count() * 8;
    cast_f32_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                                 (size_t)n);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Online feature normaliser (graphphysics/models/layers.py:281-408) and the Simulator's node-feature assembly
// (graphphysics/models/simulator.py:112-143) as three / one launches instead of ~20 elementwise ones per normaliser.
//   stats  : per-block column sums of x and x*x over a strided row range (fixed-shape trees: deterministic)
//   update : fixed-order sum of the block partials -> [sum | sum of squares | rows]; optionally the gated accumulation
//            of Normalizer._accumulate (layers.py:363-377; the gate repeats the freeze test of :347 on the device)
//   apply  : (x - mean) / max(std, eps)   or the inverse   x * max(std, eps) + mean,  mean / std from the accumulators
//            with the operation order of layers.py:379-408
namespace {
constexpr int kNormBlocks = 296, kNormThreads = 256, kNormMaxSize = 64;

__global__ void __launch_bounds__(kNormThreads) norm_stats_kernel(const float* __restrict__ x, long long rows, int size, long long ld,
                                                                  float* __restrict__ partials) {
    // thread -> (column, row phase): consecutive threads read consecutive columns of a row (rows are `size` floats wide)
    __shared__ float sh[2][kNormThreads];
    const int per = kNormThreads / size;                 // rows per pass of this block
    const int c = threadIdx.x % size, rp = threadIdx.x / size;
    float s = 0.f, q = 0.f;
    if (rp < per)
        for (long long r = (long long)blockIdx.x * per + rp; r < rows; r += (long long)gridDim.x * per) {
            const float v = x[r * ld + c];
            s += v;
            q = fmaf(v, v, q);
        }
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < size) {
        float ts = 0.f, tq = 0.f;
        for (int k = 0; k < per; ++k) {                   // fixed order
            ts += sh[0][k * size + threadIdx.x];
            tq += sh[1][k * size + threadIdx.x];
        }
        partials[(size_t)blockIdx.x * 2 * size + threadIdx.x] = ts;
        partials[(size_t)blockIdx.x * 2 * size + size + threadIdx.x] = tq;
    }
}

__global__ void __launch_bounds__(1024) norm_update_kernel(const float* __restrict__ partials, int n_blocks, int size, float rows,
                                                           float* __restrict__ stats, float* acc_sum, float* acc_sq, float* acc_count,
                                                           float* num_acc, float max_acc) {
    // thread (group g, column c): group g adds the block partials g, g + 8, ... in ascending order (independent loads, four
    // in flight), then the eight group sums are added in fixed order -- a serial walk over all blocks is one long chain of
    // L2 latencies (16 us for 296 blocks)
    __shared__ float sh[8][2 * kNormMaxSize];
    const int c = threadIdx.x & (2 * kNormMaxSize - 1), g = threadIdx.x / (2 * kNormMaxSize);
    float v = 0.f;
    if (c < 2 * size) {
        int b = g;
        for (; b + 24 < n_blocks; b += 32) {
            const float p0 = partials[(size_t)b * 2 * size + c], p1 = partials[(size_t)(b + 8) * 2 * size + c];
            const float p2 = partials[(size_t)(b + 16) * 2 * size + c], p3 = partials[(size_t)(b + 24) * 2 * size + c];
            v += p0; v += p1; v += p2; v += p3;
        }
        for (; b < n_blocks; b += 8) v += partials[(size_t)b * 2 * size + c];
    }
    sh[g][c] = v;
    __syncthreads();
    if (g != 0) return;
    v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sh[k][c];
    if (stats) {
        if (c < 2 * size) stats[c] = v;
        if (c == 0) stats[2 * size] = rows;
    }
    if (num_acc) {
        const float gate = (num_acc[0] < max_acc) ? 1.f : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(2 * kNormMaxSize) : "memory");      // group 0: everyone has read the counter before it moves
        if (c < size) acc_sum[c] += gate * v;
        else if (c < 2 * size) acc_sq[c - size] += gate * v;
        if (c == 0) {
            acc_count[0] += gate * rows;
            num_acc[0] += gate;
        }
    }
}

__global__ void norm_accumulate_kernel(const float* __restrict__ stats, int size, float* acc_sum, float* acc_sq, float* acc_count,
                                       float* num_acc, float max_acc) {
    const int c = threadIdx.x;
    const float gate = (num_acc[0] < max_acc) ? 1.f : 0.f;
    __syncthreads();
    if (c < size) acc_sum[c] += gate * stats[c];
    else if (c < 2 * size) acc_sq[c - size] += gate * stats[c];
    if (c == 0) {
        acc_count[0] += gate * stats[2 * size];
        num_acc[0] += gate;
    }
}

__global__ void norm_apply_kernel(const float* __restrict__ x, long long rows, int size, long long ld, const float* __restrict__ acc_sum,
                                  const float* __restrict__ acc_sq, const float* __restrict__ acc_count, float eps, int inverse,
                                  float* __restrict__ out, long long ld_out) {
    __shared__ float s_mean[kNormMaxSize], s_std[kNormMaxSize];
    if (threadIdx.x < size) {
        const float cnt = fmaxf(acc_count[0], 1.f);
        const float mean = __fdiv_rn(acc_sum[threadIdx.x], cnt);
        const float var = __fsub_rn(__fdiv_rn(acc_sq[threadIdx.x], cnt), __fmul_rn(mean, mean));
        s_mean[threadIdx.x] = mean;
        s_std[threadIdx.x] = fmaxf(__fsqrt_rn(fmaxf(var, 0.f)), eps);
    }
    __syncthreads();
    const long long n = rows * size;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / size;
        const int c = (int)(i - r * size);
        const float v = x[r * ld + c];
        out[r * ld_out + c] = inverse ? __fadd_rn(__fmul_rn(v, s_std[c]), s_mean[c]) : __fdiv_rn(__fsub_rn(v, s_mean[c]), s_std[c]);
    }
}

// out[r] = [ x[r, f0:f1] | one_hot(x[r, type_col], n_types) ]   (simulator.py:112-143)
__global__ void node_features_kernel(const float* __restrict__ x, long long rows, long long ld, int f0, int f1, int type_col, int n_types,
                                     float* __restrict__ out) {
    const int w = (f1 - f0) + n_types;
    const long long n = rows * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w;
        const int c = (int)(i - r * w);
        float v;
        if (c < f1 - f0) v = x[r * ld + f0 + c];
        else v = ((int)x[r * ld + type_col] == c - (f1 - f0)) ? 1.f : 0.f;
        out[i] = v;
    }
}
int norm_grid(long long n) {
    long long b = (n + 255) / 256;
    if (b > gp::sm_count() * 8) b = gp::sm_count() * 8;
    return b < 1 ? 1 : (int)b;
}
}  // namespace

extern "C" int32_t gp_normalizer_blocks(void) { return kNormBlocks; }

extern "C" int gp_normalizer_stats(const float* x, int64_t rows, int32_t size, int64_t ld, float* partials, void* stream) {
    if (size < 1 || size > kNormMaxSize || rows < 0 || !partials || (rows > 0 && !x)) {
        gp::set_error("gp_normalizer_stats: bad arguments (size 1..%d)", kNormMaxSize);
        return -1;
    }
    norm_stats_kernel<<<kNormBlocks, kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, size, ld, partials);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_normalizer_update(const float* partials, int32_t size, int64_t rows, float* stats, float* acc_sum, float* acc_sum_squared,
                                    float* acc_count, float* num_accumulations, float max_accumulations, void* stream) {
    if (size < 1 || size > kNormMaxSize || !partials || (num_accumulations && (!acc_sum || !acc_sum_squared || !acc_count))) {
        gp::set_error("gp_normalizer_update: bad arguments");
        return -1;
    }
    norm_update_kernel<<<1, 16 * kNormMaxSize, 0, static_cast<cudaStream_t>(stream)>>>(partials, kNormBlocks, size, (float)rows, stats, acc_sum,
                                                                                      acc_sum_squared, acc_count, num_accumulations,
                                                                                      max_accumulations);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_normalizer_accumulate(const float* stats, int32_t size, float* acc_sum, float* acc_sum_squared, float* acc_count,
                                        float* num_accumulations, float max_accumulations, void* stream) {
    if (size < 1 || size > kNormMaxSize || !stats || !acc_sum || !acc_sum_squared || !acc_count || !num_accumulations) {
        gp::set_error("gp_normalizer_accumulate: bad arguments");
        return -1;
    }
    norm_accumulate_kernel<<<1, 2 * kNormMaxSize, 0, static_cast<cudaStream_t>(stream)>>>(stats, size, acc_sum, acc_sum_squared, acc_count,
                                                                                          num_accumulations, max_accumulations);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_normalizer_apply(const float* x, int64_t rows, int32_t size, int64_t ld, const float* acc_sum, const float* acc_sum_squared,
                                   const float* acc_count, float std_epsilon, int32_t inverse, float* out, int64_t ld_out, void* stream) {
    if (size < 1 || size > kNormMaxSize || rows < 0 || !acc_sum || !acc_sum_squared || !acc_count || (rows > 0 && (!x || !out))) {
        gp::set_error("gp_normalizer_apply: bad arguments (size 1..%d)", kNormMaxSize);
        return -1;
    }
    if (rows == 0) return 0;
    norm_apply_kernel<<<norm_grid(rows * size), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, size, ld, acc_sum, acc_sum_squared,
                                                                                             acc_count, std_epsilon, inverse, out, ld_out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_node_features(const float* x, int64_t rows, int64_t ld, int32_t feature_start, int32_t feature_end, int32_t node_type_col,
                                int32_t num_types, float* out, void* stream) {
    if (rows < 0 || feature_end < feature_start || num_types < 0 || (rows > 0 && (!x || !out))) {
        gp::set_error("gp_node_features: bad arguments");
        return -1;
    }
    if (rows == 0) return 0;
    const long long n = rows * (long long)((feature_end - feature_start) + num_types);
    node_features_kernel<<<norm_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, ld, feature_start, feature_end, node_type_col,
                                                                                       num_types, out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
