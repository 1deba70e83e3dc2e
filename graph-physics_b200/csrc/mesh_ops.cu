// mesh_ops.cu -- graph construction on the device: what the reference's preprocessing pipeline does to a
// mesh on the host before every sample reaches the model (graphphysics/dataset/preprocessing.py:16-23 edge features,
// 92-140 world edges, 143-175 world-position features, 177-238 noise, 410-424 FaceToEdge; graphphysics/utils/
// torch_graph.py:194-210 tetrahedra -> triangles).  Integer / byte work, HBM- and latency-bound; every result is a SET
// or a per-edge value, so the kernels are free to use integer atomics for bucket sizes while the outputs stay
// bit-reproducible and bit-exact against the CPU oracle (oracle/gp_oracle.py face_to_edge / edge_features /
// world_edges):
//
//   gp_cell_edge_candidates   directed candidate pairs of every triangle / tetrahedron edge, both directions
//   gp_coalesce_count/_write  PyG coalesce: unique directed pairs sorted by (row, col).  Bucket by row (counting sort),
//                             mark the first occurrence of every column inside its bucket, scan the per-row unique
//                             counts, and let every representative take the rank of its column among the row's
//                             representatives.  O(sum of squared bucket sizes) -- mesh degrees are small.
//   gp_edge_features          [pos[row] - pos[col], ||pos[col] - pos[row]||_2] in fp32 without contraction (T.Cartesian +
//                             T.Distance, norm=False), also used for the world-position features of DeformingPlate
//   gp_world_pairs_count/_fill  radius search between OBSTACLE and NORMAL nodes on a uniform grid (the reference's
//                             cKDTree.query_pairs + node-type mask): squared distances in fp64 from the fp32 positions
//   gp_add_noise              x[:, c0:c1] += noise * scale on NORMAL nodes (training noise injection)
#include <math.h>

#include "common.cuh"
#include "../../include/gp_b200.h"

namespace {
constexpr int kTB = 256;
inline unsigned nblk(long long n) { return (unsigned)((n + kTB - 1) / kTB); }

// ------------------------------------------------------------------------------------------------ candidates
// cells: vertex-major [verts][n] (PyG `face`, torch_graph.py:197 `cells.T`) or cell-major [n][verts]
__global__ void cell_candidates_kernel(const int64_t* __restrict__ cells, long long n, int verts, int cell_major,
                                       int64_t* __restrict__ row, int64_t* __restrict__ col) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int64_t v[4];
    for (int i = 0; i < verts; ++i) v[i] = cell_major ? cells[c * verts + i] : cells[(long long)i * n + c];
    if (verts == 3) {          // FaceToEdge: (0,1), (1,2), (0,2), then to_undirected
        const int a[3] = {0, 1, 0}, b[3] = {1, 2, 2};
        for (int e = 0; e < 3; ++e) {
            row[c * 6 + 2 * e] = v[a[e]];     col[c * 6 + 2 * e] = v[b[e]];
            row[c * 6 + 2 * e + 1] = v[b[e]]; col[c * 6 + 2 * e + 1] = v[a[e]];
        }
    } else {                   // a tetrahedron contributes its 4 triangles = all 6 of its edges
        const int a[6] = {0, 0, 0, 1, 1, 2}, b[6] = {1, 2, 3, 2, 3, 3};
        for (int e = 0; e < 6; ++e) {
            row[c * 12 + 2 * e] = v[a[e]];     col[c * 12 + 2 * e] = v[b[e]];
            row[c * 12 + 2 * e + 1] = v[b[e]]; col[c * 12 + 2 * e + 1] = v[a[e]];
        }
    }
}

// ------------------------------------------------------------------------------------------------ scan (exclusive, n+1 outputs)
constexpr int kScanBlock = 1024;
__global__ void scan_block_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out, int32_t* __restrict__ sums) {
    __shared__ int32_t sh[kScanBlock];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kScanBlock; off <<= 1) {
        const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = sh[threadIdx.x] - v;
    if (threadIdx.x == kScanBlock - 1) sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void scan_sums_kernel(int32_t* sums, int nb, int32_t* total_out) {      // one block
    __shared__ int32_t carry;
    __shared__ int32_t sh[kScanBlock];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < kScanBlock; off <<= 1) {
            const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry += sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void scan_add_kernel(int32_t* __restrict__ out, int n, const int32_t* __restrict__ sums, const int32_t* __restrict__ total) {
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
    if (i == 0) out[n] = *total;
}
// out[0..n] = exclusive scan of in[0..n); ws: [nb + 1] ints
int exclusive_scan(const int32_t* in, int n, int32_t* out, int32_t* ws, cudaStream_t st) {
    const int nb = n > 0 ? (n + kScanBlock - 1) / kScanBlock : 1;
    scan_block_kernel<<<nb, kScanBlock, 0, st>>>(in, n, out, ws);
    scan_sums_kernel<<<1, kScanBlock, 0, st>>>(ws, nb, ws + nb);
    scan_add_kernel<<<nb, kScanBlock, 0, st>>>(out, n, ws, ws + nb);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void zero_kernel(int32_t* p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}
__global__ void count_rows_kernel(const int64_t* __restrict__ row, long long n, int32_t* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[(int)row[i]], 1);
}
__global__ void fill_rows_kernel(const int64_t* __restrict__ row, long long n, const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ bucket) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int r = (int)row[i];
        bucket[rowptr[r] + atomicAdd(&cursor[r], 1)] = (int32_t)i;
    }
}
// position j of the bucketed list is the representative of its (row, col) if no earlier position of the bucket holds
// the same column; ucount[row] = number of representatives
__global__ void mark_first_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, long long n,
                                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ bucket, int32_t* __restrict__ first,
                                  int32_t* __restrict__ ucount) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int id = bucket[j];
    const int r = (int)row[id];
    const int64_t c = col[id];
    int f = 1;
    for (int k = rowptr[r]; k < (int)j; ++k) f &= (col[bucket[k]] != c);
    first[j] = f;
    if (f) atomicAdd(&ucount[r], 1);
}
__global__ void write_unique_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, long long n,
                                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ bucket, const int32_t* __restrict__ first,
                                    const int32_t* __restrict__ out_rowptr, int64_t* __restrict__ out_row, int64_t* __restrict__ out_col) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !first[j]) return;
    const int id = bucket[j];
    const int r = (int)row[id];
    const int64_t c = col[id];
    int rank = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) rank += (first[k] && col[bucket[k]] < c);
    const long long o = (long long)out_rowptr[r] + rank;
    out_row[o] = r;
    out_col[o] = c;
}

// ------------------------------------------------------------------------------------------------ edge features
// out[e] = [pos[row] - pos[col] (dim values), sqrt(sum_d (pos[col] - pos[row])_d^2)]: fp32, no fused multiply-add, sums in
// index order -- numpy / ATen arithmetic on the host
__global__ void edge_features_kernel(const float* __restrict__ pos, int ld_pos, int dim, const int64_t* __restrict__ row,
                                     const int64_t* __restrict__ col, long long E, float* __restrict__ out, int ld_out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float* pr = pos + (size_t)row[e] * ld_pos;
    const float* pc = pos + (size_t)col[e] * ld_pos;
    float ss = 0.f;
    for (int d = 0; d < dim; ++d) {
        const float a = pr[d], b = pc[d];
        out[(size_t)e * ld_out + d] = __fsub_rn(a, b);
        const float diff = __fsub_rn(b, a);
        const float sq = __fmul_rn(diff, diff);
        ss = d == 0 ? sq : __fadd_rn(ss, sq);
    }
    out[(size_t)e * ld_out + dim] = __fsqrt_rn(ss);
}

// ------------------------------------------------------------------------------------------------ world edges
struct Grid {
    float lo[3];
    float inv_cell;
    int g[3];
};
// one block: bounding box of the positions -> grid with cells of at least `radius` (at most 128 cells per axis)
__global__ void grid_setup_kernel(const float* __restrict__ pos, int ld, int n, float radius, Grid* grid) {
    __shared__ float smin[3][kTB], smax[3][kTB];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int d = 0; d < 3; ++d) {
            const float v = pos[(size_t)i * ld + d];
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    for (int d = 0; d < 3; ++d) { smin[d][threadIdx.x] = mn[d]; smax[d][threadIdx.x] = mx[d]; }
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int d = 0; d < 3; ++d) {
                smin[d][threadIdx.x] = fminf(smin[d][threadIdx.x], smin[d][threadIdx.x + s]);
                smax[d][threadIdx.x] = fmaxf(smax[d][threadIdx.x], smax[d][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float ext = 0.f;
        for (int d = 0; d < 3; ++d) ext = fmaxf(ext, smax[d][0] - smin[d][0]);
        // a little wider than the radius so that rounding in the cell index can never push a pair two cells apart
        const float cell = fmaxf(radius * 1.001f, ext / 127.f);
        grid->inv_cell = 1.f / cell;
        for (int d = 0; d < 3; ++d) {
            grid->lo[d] = smin[d][0];
            grid->g[d] = min(128, (int)((smax[d][0] - smin[d][0]) / cell) + 1);
        }
    }
}
__device__ __forceinline__ void cell_of(const Grid& G, const float* p, int (&c)[3]) {
    for (int d = 0; d < 3; ++d) c[d] = max(0, min(G.g[d] - 1, (int)((p[d] - G.lo[d]) * G.inv_cell)));
}
// key of node i: its cell if it is a NORMAL node, else the extra bucket 128^3 (never searched)
constexpr int kMaxCells = 128 * 128 * 128;
__global__ void cell_keys_kernel(const float* __restrict__ pos, int ld, const float* __restrict__ type, int ld_type, int n, int normal_type,
                                 const Grid* __restrict__ grid, int32_t* __restrict__ key, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = kMaxCells;
    if (type[(size_t)i * ld_type] == (float)normal_type) {
        int c[3];
        cell_of(*grid, pos + (size_t)i * ld, c);
        k = (c[2] * 128 + c[1]) * 128 + c[0];
    }
    key[i] = k;
    atomicAdd(&count[k], 1);
}
__global__ void fill_cells_kernel(const int32_t* __restrict__ key, int n, const int32_t* __restrict__ cellptr, int32_t* __restrict__ cursor,
                                  int32_t* __restrict__ members) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) members[cellptr[key[i]] + atomicAdd(&cursor[key[i]], 1)] = i;
}
// For OBSTACLE node i: NORMAL nodes j within `radius` (squared distance in fp64, <=, like cKDTree.query_pairs).  WRITE = false
// counts them, WRITE = true writes the pairs (i, j) and (j, i) at 2 * (offset[i] + k).  The members of a cell are visited
// in bucket order, which integer atomics filled nondeterministically: the SET of pairs is what is defined, and the
// coalesce that follows sorts it.
template <bool WRITE>
__global__ void world_pairs_kernel(const float* __restrict__ pos, int ld, const float* __restrict__ type, int ld_type, int n, int obstacle_type,
                                   double r2, const Grid* __restrict__ grid, const int32_t* __restrict__ cellptr,
                                   const int32_t* __restrict__ members, int32_t* __restrict__ cnt, const int32_t* __restrict__ offset,
                                   int64_t* __restrict__ out_row, int64_t* __restrict__ out_col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int found = 0;
    if (type[(size_t)i * ld_type] == (float)obstacle_type) {
        const Grid G = *grid;
        const float* pi = pos + (size_t)i * ld;
        int c[3];
        cell_of(G, pi, c);
        const double x = pi[0], y = pi[1], z = pi[2];
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int cx = c[0] + dx, cy = c[1] + dy, cz = c[2] + dz;
                    if (cx < 0 || cy < 0 || cz < 0 || cx >= G.g[0] || cy >= G.g[1] || cz >= G.g[2]) continue;
                    const int k = (cz * 128 + cy) * 128 + cx;
                    for (int m = cellptr[k]; m < cellptr[k + 1]; ++m) {
                        const int j = members[m];
                        const float* pj = pos + (size_t)j * ld;
                        const double ax = x - (double)pj[0], ay = y - (double)pj[1], az = z - (double)pj[2];
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
                        if (d2 <= r2) {
                            if (WRITE) {
                                const long long o = 2ll * ((long long)offset[i] + found);
                                out_row[o] = i; out_col[o] = j;
                                out_row[o + 1] = j; out_col[o + 1] = i;
                            }
                            ++found;
                        }
                    }
                }
    }
    if (!WRITE) cnt[i] = found;
}

// ------------------------------------------------------------------------------------------------ noise
__global__ void add_noise_kernel(float* __restrict__ x, int ld, int rows, int c0, int width, int type_col, int normal_type,
                                 const float* __restrict__ noise, float scale) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)rows * width) return;
    const int r = (int)(g / width), c = (int)(g - (long long)r * width);
    if (x[(size_t)r * ld + type_col] == (float)normal_type) {
        float* p = x + (size_t)r * ld + c0 + c;
        *p = __fadd_rn(*p, __fmul_rn(noise[g], scale));
    }
}
}  // namespace

extern "C" int gp_cell_edge_candidates(const int64_t* cells, int64_t n_cells, int32_t verts_per_cell, int32_t cell_major, int64_t* cand_row,
                                       int64_t* cand_col, void* stream) {
    if (n_cells <= 0) return 0;
    GP_REQUIRE(verts_per_cell == 3 || verts_per_cell == 4, "gp_cell_edge_candidates: cells must be triangles (3) or tetrahedra (4)");
    GP_REQUIRE(cells && cand_row && cand_col, "gp_cell_edge_candidates: null pointer");
    cell_candidates_kernel<<<nblk(n_cells), kTB, 0, static_cast<cudaStream_t>(stream)>>>(cells, n_cells, verts_per_cell, cell_major, cand_row,
                                                                                       cand_col);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// workspace layout (int32): rowptr [N+1] | count / cursor [N+1] | out_rowptr [N+1] | scan scratch [N/1024 + 2] | bucket [n] | first [n]
extern "C" int64_t gp_coalesce_workspace_bytes(int64_t n_cand, int32_t num_nodes) {
    return 4 * (3 * ((int64_t)num_nodes + 1) + (num_nodes / kScanBlock + 3) + 2 * n_cand + 16);
}
namespace {
struct CoalesceWs {
    int32_t *rowptr, *count, *out_rowptr, *scan, *bucket, *first;
};
CoalesceWs carve(void* ws, int64_t n, int32_t N) {
    CoalesceWs w;
    w.rowptr = static_cast<int32_t*>(ws);
    w.count = w.rowptr + (N + 1);
    w.out_rowptr = w.count + (N + 1);
    w.scan = w.out_rowptr + (N + 1);
    w.bucket = w.scan + (N / kScanBlock + 3);
    w.first = w.bucket + n;
    return w;
}
}  // namespace

extern "C" int gp_coalesce_count(const int64_t* cand_row, const int64_t* cand_col, int64_t n_cand, int32_t num_nodes, void* workspace,
                                 int32_t* num_unique, void* stream) {
    GP_REQUIRE(n_cand >= 0 && n_cand < (1ll << 31) && num_nodes > 0, "gp_coalesce_count: bad sizes n=%lld N=%d", (long long)n_cand, num_nodes);
    GP_REQUIRE(workspace && num_unique && (n_cand == 0 || (cand_row && cand_col)), "gp_coalesce_count: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CoalesceWs w = carve(workspace, n_cand, num_nodes);
    zero_kernel<<<nblk(2ll * (num_nodes + 1)), kTB, 0, st>>>(w.rowptr, 2ll * (num_nodes + 1));        // rowptr + count
    if (n_cand > 0) count_rows_kernel<<<nblk(n_cand), kTB, 0, st>>>(cand_row, n_cand, w.count);
    int rc = exclusive_scan(w.count, num_nodes, w.rowptr, w.scan, st);
    if (rc) return rc;
    zero_kernel<<<nblk(num_nodes + 1), kTB, 0, st>>>(w.count, num_nodes + 1);
    if (n_cand > 0) fill_rows_kernel<<<nblk(n_cand), kTB, 0, st>>>(cand_row, n_cand, w.rowptr, w.count, w.bucket);
    zero_kernel<<<nblk(num_nodes + 1), kTB, 0, st>>>(w.count, num_nodes + 1);
    if (n_cand > 0) mark_first_kernel<<<nblk(n_cand), kTB, 0, st>>>(cand_row, cand_col, n_cand, w.rowptr, w.bucket, w.first, w.count);
    rc = exclusive_scan(w.count, num_nodes, w.out_rowptr, w.scan, st);
    if (rc) return rc;
    GP_CHECK_CUDA(cudaMemcpyAsync(num_unique, w.out_rowptr + num_nodes, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int gp_coalesce_write(const int64_t* cand_row, const int64_t* cand_col, int64_t n_cand, int32_t num_nodes, void* workspace,
                                 int64_t* out_row, int64_t* out_col, void* stream) {
    if (n_cand <= 0) return 0;
    GP_REQUIRE(workspace && cand_row && cand_col && out_row && out_col, "gp_coalesce_write: null pointer");
    const CoalesceWs w = carve(workspace, n_cand, num_nodes);
    write_unique_kernel<<<nblk(n_cand), kTB, 0, static_cast<cudaStream_t>(stream)>>>(cand_row, cand_col, n_cand, w.rowptr, w.bucket, w.first,
                                                                                   w.out_rowptr, out_row, out_col);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_edge_features(const float* pos, int32_t ld_pos, int32_t dim, const int64_t* row, const int64_t* col, int64_t num_edges,
                                float* out, int32_t ld_out, void* stream) {
    if (num_edges <= 0) return 0;
    GP_REQUIRE(pos && row && col && out && dim >= 1 && dim <= 3 && ld_out >= dim + 1, "gp_edge_features: bad arguments (dim must be 1..3)");
    edge_features_kernel<<<nblk(num_edges), kTB, 0, static_cast<cudaStream_t>(stream)>>>(pos, ld_pos, dim, row, col, num_edges, out, ld_out);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// workspace (int32): grid [16] | key [N] | cnt [N] | offset [N+1] | scan [N/1024 + 3] | members [N] | cellptr [128^3 + 2] | count [128^3 + 2]
extern "C" int64_t gp_world_pairs_workspace_bytes(int32_t num_nodes) {
    return 4 * (16 + 4 * ((int64_t)num_nodes + 1) + (num_nodes / kScanBlock + 3) + 2 * ((int64_t)kMaxCells + 2) + (kMaxCells / kScanBlock + 3));
}
namespace {
struct WorldWs {
    Grid* grid;
    int32_t *key, *cnt, *offset, *scan, *members, *cellptr, *count, *scan2;
};
WorldWs carve_world(void* ws, int32_t N) {
    WorldWs w;
    int32_t* p = static_cast<int32_t*>(ws);
    w.grid = reinterpret_cast<Grid*>(p);  p += 16;
    w.key = p;      p += N + 1;
    w.cnt = p;      p += N + 1;
    w.offset = p;   p += N + 1;
    w.members = p;  p += N + 1;
    w.scan = p;     p += N / kScanBlock + 3;
    w.cellptr = p;  p += kMaxCells + 2;
    w.count = p;    p += kMaxCells + 2;
    w.scan2 = p;
    return w;
}
}  // namespace

extern "C" int gp_world_pairs_count(const float* pos, int32_t ld_pos, const float* node_type, int32_t ld_type, int32_t num_nodes, double radius,
                                    int32_t normal_type, int32_t obstacle_type, void* workspace, int32_t* num_pairs, void* stream) {
    GP_REQUIRE(pos && node_type && workspace && num_pairs && num_nodes > 0 && radius > 0.0, "gp_world_pairs_count: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const WorldWs w = carve_world(workspace, num_nodes);
    grid_setup_kernel<<<1, kTB, 0, st>>>(pos, ld_pos, num_nodes, (float)radius, w.grid);
    zero_kernel<<<nblk(kMaxCells + 2), kTB, 0, st>>>(w.count, kMaxCells + 2);
    cell_keys_kernel<<<nblk(num_nodes), kTB, 0, st>>>(pos, ld_pos, node_type, ld_type, num_nodes, normal_type, w.grid, w.key, w.count);
    int rc = exclusive_scan(w.count, kMaxCells + 1, w.cellptr, w.scan2, st);
    if (rc) return rc;
    zero_kernel<<<nblk(kMaxCells + 2), kTB, 0, st>>>(w.count, kMaxCells + 2);
    fill_cells_kernel<<<nblk(num_nodes), kTB, 0, st>>>(w.key, num_nodes, w.cellptr, w.count, w.members);
    const double r = radius;
    world_pairs_kernel<false><<<nblk(num_nodes), kTB, 0, st>>>(pos, ld_pos, node_type, ld_type, num_nodes, obstacle_type, r * r, w.grid, w.cellptr,
                                                              w.members, w.cnt, nullptr, nullptr, nullptr);
    rc = exclusive_scan(w.cnt, num_nodes, w.offset, w.scan, st);
    if (rc) return rc;
    GP_CHECK_CUDA(cudaMemcpyAsync(num_pairs, w.offset + num_nodes, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int gp_world_pairs_fill(const float* pos, int32_t ld_pos, const float* node_type, int32_t ld_type, int32_t num_nodes, double radius,
                                   int32_t obstacle_type, void* workspace, int64_t* out_row, int64_t* out_col, void* stream) {
    GP_REQUIRE(pos && node_type && workspace && out_row && out_col, "gp_world_pairs_fill: null pointer");
    const WorldWs w = carve_world(workspace, num_nodes);
    const double r = radius;
    world_pairs_kernel<true><<<nblk(num_nodes), kTB, 0, static_cast<cudaStream_t>(stream)>>>(
        pos, ld_pos, node_type, ld_type, num_nodes, obstacle_type, r * r, w.grid, w.cellptr, w.members, nullptr, w.offset, out_row, out_col);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// k-hop adjacency, one hop (graphphysics/utils/torch_graph.py:14-54: adj_k <- adj_k + adj_k . adj without self loops): every
// entry (i, j) of adj_k contributes itself and (i, l) for every neighbour l of j in adj; a would-be self loop (l == i) is
// written as (i, j) again, so no compaction is needed before gp_coalesce_*.  offsets = exclusive prefix sum over the entries of
// 1 + deg_adj(col); adj is given as CSR over its rows (rowptr, col sorted by row).
namespace {
__global__ void khop_candidates_kernel(const int64_t* __restrict__ rowk, const int64_t* __restrict__ colk, long long nk,
                                       const int64_t* __restrict__ rowptr, const int64_t* __restrict__ adj_col,
                                       const int64_t* __restrict__ offsets, int64_t* __restrict__ cand_row, int64_t* __restrict__ cand_col) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nk; e += (long long)gridDim.x * blockDim.x) {
        const int64_t i = rowk[e], j = colk[e];
        int64_t o = offsets[e];
        cand_row[o] = i;
        cand_col[o] = j;
        ++o;
        for (int64_t q = rowptr[j]; q < rowptr[j + 1]; ++q, ++o) {
            const int64_t l = adj_col[q];
            cand_row[o] = i;
            cand_col[o] = (l == i) ? j : l;
        }
    }
}
}  // namespace

extern "C" int gp_khop_candidates(const int64_t* rowk, const int64_t* colk, int64_t num_entries, const int64_t* adj_rowptr,
                                  const int64_t* adj_col, const int64_t* offsets, int64_t* cand_row, int64_t* cand_col, void* stream) {
    if (num_entries < 0 || (num_entries > 0 && (!rowk || !colk || !adj_rowptr || !adj_col || !offsets || !cand_row || !cand_col))) {
        gp::set_error("gp_khop_candidates: bad arguments");
        return -1;
    }
    if (num_entries == 0) return 0;
    long long blocks = (num_entries + 255) / 256;
    if (blocks > gp::sm_count() * 16) blocks = gp::sm_count() * 16;
    khop_candidates_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rowk, colk, num_entries, adj_rowptr, adj_col, offsets,
                                                                                      cand_row, cand_col);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_add_noise(float* x, int32_t ld, int32_t rows, int32_t col_start, int32_t col_end, int32_t node_type_col, int32_t normal_type,
                            const float* noise, float scale, void* stream) {
    const int width = col_end - col_start;
    if (rows <= 0 || width <= 0) return 0;
    GP_REQUIRE(x && noise && col_start >= 0 && col_end <= ld && node_type_col >= 0 && node_type_col < ld, "gp_add_noise: bad arguments");
    add_noise_kernel<<<nblk((long long)rows * width), kTB, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, rows, col_start, width, node_type_col,
                                                                                                  normal_type, noise, scale);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
