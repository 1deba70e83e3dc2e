/* gp_b200.h -- C ABI of libgp_b200.so: the B200 (sm_100a) kernels behind the graph-physics
 * message-passing hot path.
 *
 * The reference (DonsetPG/graph-physics) has no FFI of its own: the path sits behind Python
 * classes and reaches third-party kernels.  Each entry point below names the reference call
 * site it replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless stated otherwise; the library never
 *     allocates or frees user-visible memory;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), no internal sync;
 *   - return value 0 = ok, <0 = error; gp_last_error() gives the thread-local message;
 *   - indices are int32, activations bf16 (raw uint16 storage), accumulators/statistics fp32;
 *   - "packed weight" = bf16 row-major [n_out][k_in] with n_out, k_in padded to multiples of 16.
 */
#ifndef GP_B200_H
#define GP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t gp_bf16; /* raw bfloat16 bits */

int gp_version(void);
const char* gp_last_error(void);
/* out[0]=SM count, out[1]=max dynamic smem per block, out[2]=compute capability major*10+minor */
int gp_device_info(int* out3);
/* sizeof(struct `name`) as compiled into the library ("gp_mlp_fwd_args", ...; -1 = unknown name): lets a
 * binding (ctypes / cffi stub) verify its own struct layout before the first call. */
int gp_sizeof_struct(const char* name);
/* Launch overlap (programmatic dependent launch): when enabled (returns the previous setting; default off),
 * gp_mlp_fwd / gp_mlp_bwd_stage / gp_linear_bwd start while the previous kernel of the stream is still
 * draining and stage their packed weights, biases and scales before waiting for it.  Enable only when those
 * parameter tensors are never written by the kernel launched immediately before one of these calls (the
 * engine packs the weights once per step, several launches earlier). */
int gp_set_launch_overlap(int enabled);

/* ---------------------------------------------------------------------------------------------
 * Fused row-tile MLP forward (tcgen05).  Replaces the nn.Sequential built by build_mlp
 * (graphphysics/models/layers.py:163-210) together with what surrounds it in
 * GraphNetBlock.forward (layers.py:989-1102): the x[col]/x[row] gathers (1016-1018), the concat
 * (1058), RMSNorm (104-129), the residual (1039-1040) and the PyG sum aggregation at
 * edge_index[1] (926, 1031-1037).
 *
 * For every row r (an edge in receiver-sorted order, or a node):
 *   z1 = a_in[r] . W0^T  +  init[idx0[r]][off0:] + init[idx1[r]][off1:]  + b0
 *   h  = relu(z1); ... hidden layers ...; m = h . W_last^T + b_last
 *   u  = norm_scale ? norm_scale * m / (||m||/sqrt(n) + 1e-8) : m
 *   y[r] = (resid ? resid[r] : 0) + u                     (bf16 or fp32)
 *   seg_out[seg_id[r]] += bf16(u)   (rows with equal seg_id are contiguous; no atomics: complete
 *       segments are stored directly, pieces cut by a sub-tile boundary go to seg_bnd and are
 *       combined in fixed order by gp_seg_fixup)
 * --------------------------------------------------------------------------------------------- */
typedef struct gp_mlp_fwd_args {
    int32_t rows;
    /* layer-0 streamed operand: exactly one of a_bf16 / a_f32 (or neither when ka == 0) */
    const gp_bf16* a_bf16;
    const float* a_f32;
    int32_t ka;  /* multiple of 16, <= 128 */
    int32_t lda; /* row stride in elements */
    /* optional pre-activation rows added to layer 0 (bf16, ld_init elements per row) */
    const gp_bf16* init;
    int32_t ld_init;
    int32_t init_off0, init_off1;
    const int32_t* idx0; /* NULL -> row id */
    const int32_t* idx1; /* used when two_inits */
    int32_t two_inits;
    /* layers */
    int32_t n_layers;    /* 1..4 */
    const gp_bf16* w[4]; /* packed [n[l]][k[l]] */
    const float* bias[4];
    int32_t k[4], n[4];
    const float* norm_scale; /* NULL -> no RMSNorm */
    const gp_bf16* resid;    /* NULL -> none; row stride ld_out */
    gp_bf16* y_bf16;
    float* y_f32;
    int32_t ld_out;
    int32_t n_valid;       /* columns of the last layer actually written */
    gp_bf16* save_h2;      /* optional [rows][hidden]: output of layer index 1 (after relu) */
    gp_bf16* save_h1;      /* optional [rows][hidden]: output of layer index 0; lets gp_mlp_bwd_stage skip its recompute */
    gp_bf16* save_h3;      /* optional [rows][hidden]: output of layer index 2 (4-layer MLPs) */
    const int32_t* seg_id; /* optional [rows], non-decreasing */
    float* seg_out;        /* [num_segments][hidden] */
    gp_bf16* seg_out_bf16; /* ... or the same sums rounded once to bf16 (exactly one of the two with seg_id) */
    float* seg_bnd;        /* [ceil(rows/sub)][2][hidden], sub = gp_seg_sub_rows(hidden, 0) */
    unsigned long long* prof; /* optional [16] device counters: SM cycles per phase, summed over tiles (tuning aid) */
} gp_mlp_fwd_args;

int gp_mlp_fwd(const gp_mlp_fwd_args* args, int hidden, void* stream);

/* Combine the boundary pieces left by a kernel's segment sum: for every segment n with rows
 * [rowptr[n], rowptr[n+1]) that is empty or spans more than one sub-tile, write seg_out[n].
 * sub_rows = rows per sub-tile of the producing kernel: hidden/4 for gp_mlp_fwd, hidden/4 or hidden/8
 * (hidden = 128) for gp_mlp_bwd_stage -- gp_seg_sub_rows() tells. */
int gp_seg_sub_rows(int32_t hidden, int32_t backward);
int gp_seg_fixup(const int32_t* rowptr, int32_t num_segments, int32_t hidden, int32_t sub_rows, const float* seg_bnd,
                 float* seg_out, void* stream);
/* Same for a kernel that wrote seg_out_bf16 (the pieces are combined in fp32 and rounded once). */
int gp_seg_fixup_bf16(const int32_t* rowptr, int32_t num_segments, int32_t hidden, int32_t sub_rows, const float* seg_bnd,
                      gp_bf16* seg_out_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Two-layer backward stage of the same MLP (tcgen05): recomputes layer `a` from its streamed
 * input (gathers are re-done, no per-row activation is read back except the one tensor the
 * forward saved between layer pairs), then does dgrad and wgrad of layers b and a.
 * Replaces torch.autograd through build_mlp / RMSNorm / the gathers / scatter_add
 * (graphphysics/models/layers.py:104-129, 163-210, 1016-1060) for
 * LightningModule.training_step (graphphysics/training/lightning_module.py:270-342).
 *
 *   h_a   = relu(a_in . Wa^T + init rows + ba)     (recomputed -- or read back from ha_saved, the copy the
 *                                                   forward wrote (save_h1 / save_h3); then init / idx are unused)
 *   mode 1 (NORM):  m = h_a . Wb^T + bb;  du = gy (+ gy_gather[gy_idx]);  delta_b = d RMSNorm(m)/dm . du
 *   mode 0 (GIVEN): delta_b read from memory
 *   dWb += delta_b^T h_a ;  dbb += sum delta_b ;  dscale += sum du * m/(rms+eps)
 *   delta_a = (delta_b . Wb) * (h_a > 0)
 *   dWa += delta_a^T a_in ;  dba += sum delta_a
 *   d_in = delta_a . Wa  [* (a_in > 0)] [+ out_resid]   -> out
 *   delta_a optionally stored, and segment-summed by seg_id like gp_mlp_fwd.
 * Weight-gradient sums live in TMEM for the whole launch; each CTA then writes one partial
 * block (layout from gp_mlp_bwd_layout) and gp_reduce_partials adds them in CTA order.
 * --------------------------------------------------------------------------------------------- */
typedef struct gp_mlp_bwd_args {
    int32_t rows;
    const gp_bf16* a_bf16;
    const float* a_f32;
    int32_t ka, lda;
    const gp_bf16* init;
    int32_t ld_init, init_off0, init_off1;
    const int32_t* idx0;
    const int32_t* idx1;
    int32_t two_inits;
    const gp_bf16* ha_saved; /* optional [rows][hidden] bf16: h_a as stored by the forward */
    const gp_bf16* wa; /* packed [hidden][ka] */
    const float* ba;
    const gp_bf16* wb; /* packed [nb][hidden] */
    const float* bb;
    int32_t nb;
    int32_t mode; /* 0 GIVEN, 1 NORM */
    const gp_bf16* delta_b;
    int32_t ld_db;
    const float* norm_scale;
    const gp_bf16* gy_bf16;
    const float* gy_f32;
    int32_t ld_gy;
    const float* gy_gather; /* optional fp32 [.][hidden] rows added to gy */
    const gp_bf16* gy_gather_bf16; /* the same rows stored as bf16 (needs gy_bf16; at most one of the two) */
    const int32_t* gy_idx;
    int32_t need_din;
    int32_t mask_by_ain;
    const gp_bf16* out_resid;
    gp_bf16* out_bf16;
    float* out_f32;
    int32_t ld_out;
    gp_bf16* delta_a_out; /* optional [rows][hidden] */
    const int32_t* seg_id;
    float* seg_out;
    gp_bf16* seg_out_bf16; /* alternative to seg_out: sums rounded once to bf16 */
    float* seg_bnd;
    float* partials; /* [grid][stride] floats, grid <= SM count */
    unsigned long long* prof; /* optional [16] device counters: SM cycles per phase (tuning aid) */
} gp_mlp_bwd_args;

/* out6 = {off_dWb, off_dWa, off_dbb, off_dba, off_dscale, stride} in floats.
 * dWb is [nb][hidden], dWa is [hidden][ka] (row-major, padded shapes). */
int gp_mlp_bwd_layout(int hidden, int ka, int nb, int32_t* out6);
/* Launches the stage; *grid_out (host) receives the number of partial blocks written. */
int gp_mlp_bwd_stage(const gp_mlp_bwd_args* args, int hidden, int32_t* grid_out, void* stream);
/* dst[r*ld_dst + c] (+)= sum_p partials[p*stride + offset + r*ld_part + c], p ascending. */
int gp_reduce_partials(const float* partials, int32_t n_parts, int32_t stride, int32_t offset, int32_t rows,
                       int32_t cols, int32_t ld_part, float* dst, int32_t ld_dst, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the bias-free per-node projection P = x . Wp^T (tcgen05).  W1 of the edge MLP is
 * [W1e | W1d | W1s] over the concat [e, x[dst], x[src]] (layers.py:1016-1018, 1058) and W1 of the
 * node MLP is [W1x | W1a] over [x, agg] (layers.py:1100-1102); by linearity the x-dependent
 * column blocks are applied once per node (Wp = [W1d; W1s; W1x], P = x.Wp^T) and P's rows are
 * gathered into the fused kernels as pre-activations.  This entry point is autograd through that
 * projection:  dx_out = dx_in + sum_s src_s . Wp_s ;  dWp_s = src_s^T . x  (per-CTA partials,
 * [grid][n_src*hidden*hidden] floats, reduce with gp_reduce_partials).
 * --------------------------------------------------------------------------------------------- */
typedef struct gp_linear_bwd_args {
    int32_t rows;
    int32_t n_src;
    const float* src_f32[3];
    const gp_bf16* src_bf16[3];
    int32_t ld_src[3];
    const gp_bf16* w; /* packed [n_src*hidden][hidden] */
    const gp_bf16* x; /* [rows][ldx] */
    int32_t ldx;
    const float* dx_in; /* optional [rows][hidden] */
    float* dx_out;      /* [rows][hidden] */
    float* partials;
} gp_linear_bwd_args;
int gp_linear_bwd(const gp_linear_bwd_args* args, int hidden, int32_t* grid_out, void* stream);

/* out[n][:] = sum_{j in [rowptr[n], rowptr[n+1])} src[perm[j]][:]  -- the sender-side sum of the
 * edge gradients (the transpose of the x[row] gather, layers.py:1018), fixed order, no atomics.
 * perm may be NULL (identity). */
int gp_segsum_gather(const gp_bf16* src, int32_t ld, const int32_t* perm, const int32_t* rowptr, int32_t num_segments,
                     int32_t hidden, float* out, void* stream);
/* Same sums (fp32 accumulate), rounded once to bf16. */
int gp_segsum_gather_bf16(const gp_bf16* src, int32_t ld, const int32_t* perm, const int32_t* rowptr, int32_t num_segments,
                          int32_t hidden, gp_bf16* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step glue (graphphysics/training/lightning_module.py:270-342, 494-511;
 * graphphysics/utils/loss.py:19-75; train.py:276-290).
 * --------------------------------------------------------------------------------------------- */
/* loss[0] = mean over rows with mask!=0 and all d columns of (out-target)^2; grad (optional) =
 * grad_scale * dloss/dout.  workspace: 260 floats. */
int gp_masked_mse(const float* out, const float* target, const uint8_t* mask, int32_t n, int32_t d, float* loss,
                  float* grad, float grad_scale, float* workspace, void* stream);
/* out[0] = sum g[i]^2 (two fixed-shape passes; workspace >= 256 floats). */
int gp_sqnorm(const float* g, int64_t n, float* workspace, float* out, void* stream);
/* clip-by-global-norm (max_norm <= 0 disables; sqnorm = device scalar from gp_sqnorm) + AdamW. */
int gp_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
             float beta2, float eps, float weight_decay, int32_t step, float max_norm, const float* sqnorm, void* stream);
/* Same, with the step counter on the device (state[0] = optimizer steps done; incremented here) and the
 * cosine-warm-up factor of graphphysics/utils/scheduler.py:51-67 computed on the device: no host value
 * changes between steps, so a whole training step can be replayed as a CUDA graph. */
int gp_adamw_sched(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t* state,
                   float base_lr, int32_t warmup, int32_t max_iters, float min_lr_factor, float beta1, float beta2,
                   float eps, float weight_decay, float max_norm, const float* sqnorm, void* stream);
/* fp32 master matrices -> packed bf16 operands, one launch for the whole model. */
typedef struct gp_pack_entry {
    int64_t src_off; /* floats into params */
    int32_t ld_src, src_col0;
    int32_t n, k; /* block copied: n rows x k columns */
    int64_t dst_off; /* elements into packed */
    int32_t ld_dst, dst_row0, dst_col0;
} gp_pack_entry;
int gp_pack_weights(const float* params, gp_bf16* packed, const gp_pack_entry* table, int32_t n_entries, void* stream);
int gp_cast_bf16(const float* src, gp_bf16* dst, int64_t n, void* stream);

/* One launch for the gradient pieces of one or several stages (at most 32 segments): for each segment s,
 * dst_s[r*ld_dst + c] (+)= sum_p partials_s[p*stride_s + offset_s + r*ld_part_s + c], p < n_parts_s.
 * A segment with partials == NULL uses the call-level partials / n_parts / stride. */
typedef struct gp_reduce_seg {
    int32_t offset, rows, cols, ld_part;
    float* dst;
    int32_t ld_dst, accumulate;
    const float* partials;
    int32_t n_parts, stride;
} gp_reduce_seg;
int gp_reduce_partials_multi(const float* partials, int32_t n_parts, int32_t stride, const gp_reduce_seg* segs_host,
                             int32_t n_segs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Adjacency-masked multi-head attention over a CSR graph: the DGL-sparse path of
 * graphphysics/models/layers.py:493-561 (bsddmm, row softmax, bspmm) as used by Attention.forward
 * (layers.py:637-697).  q, k, v, y (fp32, or bf16 with io_bf16) and dy, dq, dk, dv (fp32) are [n][hidden] with the reference's head
 * layout (channel c = d_idx*num_heads + h).  Rows are edge_index[0], columns edge_index[1].
 *   rowptr/col : the entries sorted by row (CSR);      pos[p] = index of row-sorted entry p in the
 *   colptr/row : the entries sorted by column (CSC);             column-sorted list
 *   lse        : [n][num_heads] log-sum-exp saved by the forward
 *   edge_a/ds  : [nnz][num_heads] scratch written by the backward (column-sorted order)
 * Forward and backward are gather-only (no atomics) and bit-reproducible.
 * --------------------------------------------------------------------------------------------- */
typedef struct gp_attention_args {
    int32_t n, hidden, num_heads;
    const float* q;
    const float* k;
    const float* v;
    const int32_t* rowptr;
    const int32_t* col;
    float* y;
    float* lse;
    /* backward only */
    const float* dy;
    const int32_t* pos;
    const int32_t* colptr;
    const int32_t* row;
    float* dq;
    float* dk;
    float* dv;
    float* edge_a;
    float* edge_ds;
    int32_t io_bf16; /* != 0: q, k, v and y are bf16 (raw uint16) instead of fp32 -- the Transformer path; everything else stays fp32 */
    int32_t ld_qkv;  /* row stride (elements) of q, k, v; 0 = hidden.  > hidden: q | k | v are column blocks of one [N, 3*hidden] buffer */
    int32_t ld_dqkv; /* row stride of dq, dk, dv; 0 = hidden */
    float* y_f32;    /* optional, with io_bf16: the forward also writes y unrounded here and the backward reads it (dy.y is the
                        softmax-gradient offset of every entry of the row: kept exact, 4H bytes per NODE) */
} gp_attention_args;
int gp_csr_attention_fwd(const gp_attention_args* args, void* stream);
int gp_csr_attention_bwd(const gp_attention_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Graph layout (integer work).  The reference hands COO edge lists (edge_index [2][E] int64, PyG order: sorted by
 * sender) and aggregates at edge_index[1] (graphphysics/models/layers.py:926, 1016-1018, 1031-1037); the fused
 * kernels want every receiver's edges contiguous.  gp_csr_from_coo builds, deterministically and without a host
 * round trip (capturable in a CUDA graph):
 *   perm_dst[p]   original edge id of sorted position p (STABLE sort by receiver)
 *   src_sorted / dst_sorted [E]   endpoints in sorted order;   rowptr_dst [N+1]   receiver segments
 *   perm_src[p]   positions of the sorted list, stably sorted by sender;   rowptr_src [N+1]   sender segments
 *   att_col[p]    (optional) dst_sorted[perm_src[p]]: column of row-sorted entry p (attention kernels)
 * workspace: gp_csr_workspace_bytes(E, N) bytes.  prev_edge_index ([2][E] int64) + state ([2] int32, zero-initialised),
 * both optional and owned by the caller across calls: when given, the call first compares edge_index with the copy of
 * the previous call and every kernel returns at once if nothing changed (static meshes under graph replay).
 * --------------------------------------------------------------------------------------------- */
int64_t gp_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes);
int gp_csr_from_coo(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes, int32_t* perm_dst, int32_t* src_sorted,
                    int32_t* dst_sorted, int32_t* rowptr_dst, int32_t* perm_src, int32_t* rowptr_src, int32_t* att_col,
                    void* workspace, int64_t* prev_edge_index, int32_t* state, void* stream);

/* Halo exchange of the node-partitioned mode: rows of a [.][ld] matrix (bf16: elem_bytes 2, fp32: 4).
 *   gp_halo_pack:        out[r] = x[idx[r]]                      (send buffer, contiguous rows of `cols` elements)
 *   gp_halo_unpack:      x[idx[r]] = in[r]                       (ghost rows overwritten by their owners' values)
 *   gp_halo_unpack_add:  x[dst_rows[d]] += sum_{j in [rowptr[d], rowptr[d+1])} in[order[j]]   (fp32; the transpose of
 *       pack for the backward: a row sent to several peers collects all of them in fixed order, no atomics) */
int gp_halo_pack(const void* x, int32_t ld, int32_t elem_bytes, const int32_t* idx, int32_t n, int32_t cols, void* out, void* stream);
int gp_halo_unpack(void* x, int32_t ld, int32_t elem_bytes, const int32_t* idx, int32_t n, int32_t cols, const void* in, void* stream);
int gp_halo_unpack_add(float* x, int32_t ld, const int32_t* dst_rows, const int32_t* rowptr, const int32_t* order, int32_t n_dst,
                       int32_t cols, const float* in, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Strided row-tile GEMM (tcgen05): the dense building block of the graph-Transformer path and of the "tight" mode.
 *   C(m, n) = [C(m, n) +] [resid(m, n) +] bias[n] + sum_k A(m, k) * B(n, k),  then max(., 0) if relu
 *   A(m, k) = a[m*a_sm + k*a_sk],  B(n, k) = b[n*b_sn + k*b_sk],  C(m, n) = c[m*c_sm + n*c_sn]  (element strides)
 * a / b / c are fp32 or bf16 (a_bf16 / b_bf16 / c_bf16 != 0); resid is fp32, indexed like C.  By choice of strides
 * this is torch.nn.functional.linear (graphphysics/models/layers.py:163-210, 213-278, 637-697), its dgrad or its wgrad.
 *   terms = 1: operands rounded to bf16, one MMA per k-step (Transformer block);
 *   terms = 3: fp32 operands split into three bf16 terms hi + mid + lo (24 mantissa bits), six MMAs per k-step, smallest
 *              terms first: fp32-grade products on the bf16 tensor path (precision="tight", SURVEY §7 iii).
 * split_k > 1 cuts K over CTAs (partials: split_k * M * ((N + 3) & ~3) floats of scratch, reduced in fixed order).
 * `flags` is filled by the library.
 * --------------------------------------------------------------------------------------------- */
typedef struct gp_gemm_args {
    int32_t M, N, K;
    const void* a;
    int64_t a_sm, a_sk;
    const void* b;
    int64_t b_sn, b_sk;
    void* c;
    int64_t c_sm, c_sn;
    const float* bias;
    const float* resid;
    int32_t a_bf16, b_bf16, c_bf16;
    int32_t relu, accumulate;
    int32_t terms;
    int32_t split_k;
    float* partials;
    int32_t flags;   /* filled in by gp_gemm (staging modes) */
    int32_t b_ones;  /* != 0: B has one more row than its memory, reading as 1.0: row N-1 of the result's N axis is sum_k A(m,k)
                        (a bias gradient as the last column of a weight gradient); N counts that row */
} gp_gemm_args;
int gp_gemm(const gp_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise kernels of the graph-Transformer block (graphphysics/models/layers.py:104-129 RMSNorm, 213-249 GatedMLP,
 * 766-819 Transformer.forward) around gp_gemm and gp_csr_attention_*.  fp32 arithmetic, no atomics; the *_blocks()
 * helpers give the number of per-block partial rows that gp_reduce_partials then adds in fixed order.
 *   gp_rmsnorm_fwd : y = s1*x/(rms(x)+1e-8); with scale2 the second norm of the gated MLP is applied on top
 *                    (Transformer.norm2 followed by build_gated_mlp's own RMSNorm); output bf16 (MMA operand) or fp32
 *   gp_rmsnorm_bwd : dx = [add +] J^T dy through the same one or two norms (recomputed from x; `add`, e.g. the gradient
 *                    of the residual path, may be NULL or alias dx);
 *                    part1 / part2: [gp_rmsnorm_bwd_blocks(rows)][hidden] partial sums of dscale1 / dscale2
 *   gp_gelu_gate_* : g = GELU(a1) * a2 elementwise over n contiguous fp32 values (a1 = linear1(x), a2 = linear2(x)), exact
 *                    (erf) GELU, bf16 result (the operand of the last Linear); backward
 *   gp_relu_bwd    : d[i] = h[i] > 0 ? d[i] : 0 in place (h fp32 or bf16)
 *   gp_colsum      : partials[gp_colsum_blocks(rows)][cols] of the column sums of src (bias gradients); round_bf16 sums
 *                    the bf16-rounded values (the delta is a bf16 MMA operand of the matching dgrad / wgrad)
 * --------------------------------------------------------------------------------------------- */
int gp_rmsnorm_fwd(const float* x, int32_t ldx, int32_t rows, int32_t hidden, const float* scale1, const float* scale2,
                   gp_bf16* out_bf16, float* out_f32, int32_t ld_out, void* stream);
int gp_rmsnorm_bwd_blocks(int32_t rows);
int gp_rmsnorm_bwd(const float* x, int32_t ldx, int32_t rows, int32_t hidden, const float* scale1, const float* scale2,
                   const float* dy, int32_t ld_dy, float* dx, int32_t ld_dx, const float* add, int32_t ld_add, float* part1,
                   float* part2, void* stream);
int gp_gelu_gate_fwd(const float* a1, const float* a2, int64_t n, gp_bf16* g_bf16, float* g_f32, void* stream);
int gp_gelu_gate_bwd(const float* a1, const float* a2, const float* dg, int64_t n, float* da1, float* da2, void* stream);
int gp_relu_bwd(float* d, const void* h, int32_t h_bf16, int64_t n, void* stream);
int gp_colsum_blocks(int32_t rows);
int gp_colsum(const float* src, int32_t ld, int32_t rows, int32_t cols, int32_t round_bf16, float* partials, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Graph construction on the device -- what the reference's preprocessing does on the host for every
 * sample: graphphysics/dataset/preprocessing.py:410-424 T.FaceToEdge (+ graphphysics/utils/torch_graph.py:194-210
 * tetrahedra -> triangles), :16-23 T.Cartesian + T.Distance, :92-140 add_world_edges (cKDTree.query_pairs + node-type
 * mask + to_undirected), :143-175 add_world_pos_features, :177-238 add_noise.  All index arrays int64 like edge_index.
 *   gp_cell_edge_candidates : cells (3 or 4 vertices; vertex-major [verts][n] like PyG `face`, or cell-major [n][verts])
 *                             -> 6 (12) directed candidate pairs per cell in cand_row / cand_col
 *   gp_coalesce_count       : buckets / marks the candidates, writes the number of unique directed pairs to *num_unique
 *                             (device int32); gp_coalesce_write then writes them sorted by (row, col) -- PyG coalesce.
 *                             workspace: gp_coalesce_workspace_bytes(n_cand, num_nodes), shared by the two calls
 *   gp_edge_features        : out[e] = [pos[row]-pos[col] (dim values), ||pos[col]-pos[row]||_2], fp32, out row stride ld_out
 *   gp_world_pairs_count    : OBSTACLE-NORMAL node pairs within `radius` (fp64 squared distance, <=); *num_pairs (device
 *                             int32) = number of unordered pairs; gp_world_pairs_fill writes both directions of every
 *                             pair (2 * num_pairs entries).  workspace: gp_world_pairs_workspace_bytes(num_nodes)
 *   gp_add_noise            : x[r, col_start:col_end] += noise[r, :] * scale where x[r, node_type_col] == normal_type
 *   gp_khop_candidates      : one hop of compute_k_hop_edge_index (graphphysics/utils/torch_graph.py:14-54): for every entry (i, j)
 *                             of adj_k the candidates (i, j) and (i, l), l in adj[j] (a self loop is written as (i, j) again), at
 *                             offsets[e] = exclusive prefix sum of 1 + deg_adj(col); gp_coalesce_* then gives adj_k + adj_k . adj
 * --------------------------------------------------------------------------------------------- */
int gp_cell_edge_candidates(const int64_t* cells, int64_t n_cells, int32_t verts_per_cell, int32_t cell_major, int64_t* cand_row,
                            int64_t* cand_col, void* stream);
int64_t gp_coalesce_workspace_bytes(int64_t n_cand, int32_t num_nodes);
int gp_coalesce_count(const int64_t* cand_row, const int64_t* cand_col, int64_t n_cand, int32_t num_nodes, void* workspace,
                      int32_t* num_unique, void* stream);
int gp_coalesce_write(const int64_t* cand_row, const int64_t* cand_col, int64_t n_cand, int32_t num_nodes, void* workspace,
                      int64_t* out_row, int64_t* out_col, void* stream);
int gp_edge_features(const float* pos, int32_t ld_pos, int32_t dim, const int64_t* row, const int64_t* col, int64_t num_edges,
                     float* out, int32_t ld_out, void* stream);
int64_t gp_world_pairs_workspace_bytes(int32_t num_nodes);
int gp_world_pairs_count(const float* pos, int32_t ld_pos, const float* node_type, int32_t ld_type, int32_t num_nodes, double radius,
                         int32_t normal_type, int32_t obstacle_type, void* workspace, int32_t* num_pairs, void* stream);
int gp_world_pairs_fill(const float* pos, int32_t ld_pos, const float* node_type, int32_t ld_type, int32_t num_nodes, double radius,
                        int32_t obstacle_type, void* workspace, int64_t* out_row, int64_t* out_col, void* stream);
int gp_khop_candidates(const int64_t* rowk, const int64_t* colk, int64_t num_entries, const int64_t* adj_rowptr, const int64_t* adj_col,
                       const int64_t* offsets, int64_t* cand_row, int64_t* cand_col, void* stream);
int gp_add_noise(float* x, int32_t ld, int32_t rows, int32_t col_start, int32_t col_end, int32_t node_type_col, int32_t normal_type,
                 const float* noise, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Online feature normaliser and the Simulator's node features (graphphysics/models/layers.py:281-408 Normalizer,
 * graphphysics/models/simulator.py:112-143 one-hot node type + feature slice).  fp32, rows of `size` <= 64 values.
 *   gp_normalizer_stats      : per-block column sums of x and x^2 -> partials [gp_normalizer_blocks()][2 * size]
 *   gp_normalizer_update     : fixed-order sum of the partials -> stats [sum | sum of squares | rows] (2 * size + 1 floats,
 *                              may be NULL) and, when num_accumulations is given, Normalizer._accumulate (layers.py:363-377)
 *                              gated on the device by num_accumulations < max_accumulations (the freeze of layers.py:347)
 *   gp_normalizer_accumulate : the same accumulation from a stats vector (e.g. summed over ranks by an all-reduce)
 *   gp_normalizer_apply      : out = (x - mean) / max(std, eps), or with inverse != 0  out = x * max(std, eps) + mean
 *   gp_node_features         : out[r] = [ x[r, feature_start:feature_end] | one_hot(x[r, node_type_col], num_types) ]
 * --------------------------------------------------------------------------------------------- */
int32_t gp_normalizer_blocks(void);
int gp_normalizer_stats(const float* x, int64_t rows, int32_t size, int64_t ld, float* partials, void* stream);
int gp_normalizer_update(const float* partials, int32_t size, int64_t rows, float* stats, float* acc_sum, float* acc_sum_squared,
                         float* acc_count, float* num_accumulations, float max_accumulations, void* stream);
int gp_normalizer_accumulate(const float* stats, int32_t size, float* acc_sum, float* acc_sum_squared, float* acc_count,
                             float* num_accumulations, float max_accumulations, void* stream);
int gp_normalizer_apply(const float* x, int64_t rows, int32_t size, int64_t ld, const float* acc_sum, const float* acc_sum_squared,
                        const float* acc_count, float std_epsilon, int32_t inverse, float* out, int64_t ld_out, void* stream);
int gp_node_features(const float* x, int64_t rows, int64_t ld, int32_t feature_start, int32_t feature_end, int32_t node_type_col,
                     int32_t num_types, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Variant flags of the path (graphphysics/models/layers.py:213-249 GatedMLP with SiLU / GELU, 410-491 RoPE on
 * q / k, 637-697 gated attention, 989-1149 GraphNetBlock with use_rope / use_gate): row-wise fp32 kernels around gp_gemm.
 *   kind: 1 ReLU, 2 SiLU, 3 GELU (exact).  gp_act_bwd multiplies d in place by act'(z).  gp_glu_*: g = act(a1) * a2 on
 *   [rows, cols] blocks with row stride ld.  gp_sigmoid_mul_*: out = v * sigmoid(logits).  gp_add_outer: logits[r, c] +=
 *   phi[r] * vec[c].  gp_rope_rel: relative rotary embedding of gathered sender rows by pos[src] - pos[dst]; gp_rope_nodes:
 *   rotary embedding of per-node q / k in the (N, head_dim, heads) layout, in place; `inverse` applies the transpose (the
 *   backward).  gp_concat_rows / gp_split_cols: [a | b | c] per row and its transpose (column block copy / accumulate).
 * --------------------------------------------------------------------------------------------- */
int gp_act_fwd(const float* z, int64_t n, int32_t kind, gp_bf16* out_bf16, float* out_f32, void* stream);
int gp_act_bwd(const float* z, int64_t n, int32_t kind, float* d, void* stream);
int gp_glu_fwd(const float* a1, const float* a2, int32_t ld, int64_t rows, int32_t cols, int32_t kind, gp_bf16* out_bf16, float* out_f32,
               void* stream);
int gp_glu_bwd(const float* a1, const float* a2, int32_t ld, const float* dg, int64_t rows, int32_t cols, int32_t kind, float* da1,
               float* da2, int32_t ld_d, void* stream);
int gp_sigmoid_mul_fwd(const float* logits, const float* v, int64_t n, float* out, void* stream);
int gp_sigmoid_mul_bwd(const float* logits, const float* v, const float* dout, int64_t n, float* dlogits, float* dv, void* stream);
int gp_add_outer(float* logits, const float* phi, const float* vec, int64_t rows, int32_t cols, void* stream);
int gp_rope_rel(const float* x, const float* pos, int32_t ld_pos, const int32_t* src, const int32_t* dst, int64_t num_edges, int32_t hidden,
                int32_t axes, int32_t pair_count, float base, int32_t inverse, float* out, void* stream);
int gp_rope_nodes(float* t, const float* pos, int32_t ld_pos, int64_t num_nodes, int32_t head_dim, int32_t num_heads, int32_t pos_dim,
                  int32_t m, float base, int32_t inverse, void* stream);
int gp_concat_rows(const float* a, int32_t wa, const float* b, int32_t wb, const float* c, int32_t wc, int64_t rows, float* out, void* stream);
int gp_split_cols(const float* d, int32_t width, int32_t col0, int32_t w, int64_t rows, float* out, int32_t accumulate, void* stream);
/* out[n] = sum_{p in [rowptr[n], rowptr[n+1])} src[perm ? perm[p] : p]  (fp32 rows, ascending p: the receiver sum of the variant
 * path and the transpose of its row gathers) */
int gp_segsum_rows_f32(const float* src, int32_t ld, const int32_t* perm, const int32_t* rowptr, int64_t num_segments, int32_t hidden,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GP_B200_H */
