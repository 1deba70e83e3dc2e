// graph_ops.cu -- graph layout and halo kernels around the message-passing path (integer / byte work, HBM-bound).
//
// gp_csr_from_coo: the receiver-sorted edge layout the fused kernels run on, from the reference's COO edge_index
//   (graphphysics/models/layers.py:1016-1018 gathers by edge_index, PyG propagate aggregates at edge_index[1]):
//   a STABLE counting sort by receiver, then a stable counting sort of the sorted list by sender (for the backward
//   and for the attention kernels' row view).  Deterministic and bit-exact against the oracle
//   (oracle/gp_oracle.py::csr_by_receiver): buckets are filled in arbitrary order with integer atomics and then every
//   entry takes the rank of its original index inside its bucket.  All kernels first look at a device flag that an
//   optional comparison against the previous edge_index sets: inside a replayed CUDA graph an unchanged topology
//   costs one 12 MB comparison instead of four radix sorts.
// gp_halo_pack / gp_halo_unpack_add: row gathers / scatters (add) for the halo exchange of the node-partitioned mode.
#include "common.cuh"
#include "../../include/gp_b200.h"

namespace {

// state[0] = 1 once `prev` holds a valid copy; state[1] = 1 if the layout must be rebuilt in this call
__global__ void csr_begin_kernel(int32_t* state, int force) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[1] = (force || state[0] == 0) ? 1 : 0;
}
__global__ void csr_compare_kernel(const int64_t* __restrict__ ei, const int64_t* __restrict__ prev, long long n, int32_t* state) {
    if (state[1]) return;      // already known to differ
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    int diff = 0;
    for (; i < n; i += stride) diff |= (ei[i] != prev[i]);
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(&state[1], 1);
}
__global__ void csr_commit_kernel(const int64_t* __restrict__ ei, int64_t* __restrict__ prev, long long n, int32_t* state) {
    if (!state[1]) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) prev[i] = ei[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = 1;
}

__global__ void zero_i32_kernel(int32_t* p, int n, const int32_t* state) {
    if (state && !state[1]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}
// keys: int64 (edge_index row) or int32 gathered through `via` (the receiver-sorted sender list)
template <typename KeyT>
__global__ void count_kernel(const KeyT* __restrict__ keys, int n, int32_t* __restrict__ count, const int32_t* state) {
    if (state && !state[1]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[(int)keys[i]], 1);
}

// exclusive scan of count[0..n) -> out[0..n], out[n] = total; three launches (block sums, scan of sums, add)
constexpr int kScanBlock = 1024;
__global__ void scan_block_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out, int32_t* __restrict__ sums,
                                  const int32_t* state) {
    if (state && !state[1]) return;
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
    if (i < n) out[i] = sh[threadIdx.x] - v;              // exclusive
    if (threadIdx.x == kScanBlock - 1) sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void scan_sums_kernel(int32_t* sums, int nb, const int32_t* state) {        // one block; nb <= a few thousand
    if (state && !state[1]) return;
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    __shared__ int32_t sh[kScanBlock];
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
}
__global__ void scan_add_kernel(int32_t* __restrict__ out, int n, const int32_t* __restrict__ sums, int total, const int32_t* state) {
    if (state && !state[1]) return;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
    if (i == 0) out[n] = total;
}

// unordered bucket fill: tmp[rowptr[key] + slot] = i, slot from an integer atomic on cursor[key]
template <typename KeyT>
__global__ void fill_kernel(const KeyT* __restrict__ keys, int n, const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                            int32_t* __restrict__ tmp, const int32_t* state) {
    if (state && !state[1]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int k = (int)keys[i];
        tmp[rowptr[k] + atomicAdd(&cursor[k], 1)] = i;
    }
}
// every entry takes the rank of its index inside its bucket: perm[rowptr[k] + #{y in bucket: y < x}] = x  (stable order)
template <typename KeyT>
__global__ void rank_kernel(const KeyT* __restrict__ keys, int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tmp,
                            int32_t* __restrict__ perm, const int32_t* state) {
    if (state && !state[1]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = (int)keys[i];
    const int b = rowptr[k], e = rowptr[k + 1];
    int r = 0;
    for (int j = b; j < e; ++j) r += (tmp[j] < i);
    perm[b + r] = i;
}
__global__ void gather_sorted_kernel(const int64_t* __restrict__ ei, int E, const int32_t* __restrict__ perm, int32_t* __restrict__ src_s,
                                     int32_t* __restrict__ dst_s, const int32_t* state) {
    if (state && !state[1]) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < E) {
        const int i = perm[p];
        src_s[p] = (int32_t)ei[i];
        dst_s[p] = (int32_t)ei[(size_t)E + i];
    }
}
__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int n, int32_t* __restrict__ out,
                                  const int32_t* state) {
    if (state && !state[1]) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = src[idx[p]];
}

template <typename KeyT>
int counting_sort(const KeyT* keys, int n, int num_keys, int32_t* rowptr, int32_t* perm, int32_t* ws, const int32_t* state,
                  cudaStream_t st) {
    // ws: count / cursor [num_keys + 1] | block sums [nb] | tmp [n]
    const int nb = num_keys > 0 ? (num_keys + kScanBlock - 1) / kScanBlock : 1;
    int32_t* count = ws;
    int32_t* sums = ws + (num_keys + 1);
    int32_t* tmp = sums + nb;
    const int tb = 256;
    zero_i32_kernel<<<(num_keys + 1 + tb - 1) / tb, tb, 0, st>>>(count, num_keys + 1, state);
    if (n > 0) count_kernel<KeyT><<<(n + tb - 1) / tb, tb, 0, st>>>(keys, n, count, state);
    scan_block_kernel<<<nb, kScanBlock, 0, st>>>(count, num_keys, rowptr, sums, state);
    scan_sums_kernel<<<1, kScanBlock, 0, st>>>(sums, nb, state);
    scan_add_kernel<<<nb, kScanBlock, 0, st>>>(rowptr, num_keys, sums, n, state);
    zero_i32_kernel<<<(num_keys + 1 + tb - 1) / tb, tb, 0, st>>>(count, num_keys + 1, state);
    if (n > 0) {
        fill_kernel<KeyT><<<(n + tb - 1) / tb, tb, 0, st>>>(keys, n, rowptr, count, tmp, state);
        rank_kernel<KeyT><<<(n + tb - 1) / tb, tb, 0, st>>>(keys, n, rowptr, tmp, perm, state);
    }
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- halo rows
template <typename T>
__global__ void halo_pack_kernel(const T* __restrict__ x, int ld, const int32_t* __restrict__ idx, int n, int cols, T* __restrict__ out) {
    // one 16-byte chunk per thread
    constexpr int V = 16 / sizeof(T);
    const int cpr = cols / V;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * cpr) return;
    const int r = (int)(g / cpr), c = (int)(g - (long long)r * cpr);
    reinterpret_cast<uint4*>(out + (size_t)r * cols)[c] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)idx[r] * ld) + c);
}
// x[idx[r]] += in[r]  (idx rows may repeat: a row sent to several peers collects all of them, in ascending r -- one thread owns
// a (destination row, chunk) pair through the CSR `rowptr` over destinations, so the sum order is fixed and atomic-free)
__global__ void halo_unpack_add_kernel(float* __restrict__ x, int ld, const int32_t* __restrict__ dst_rows, const int32_t* __restrict__ rowptr,
                                       const int32_t* __restrict__ order, int n_dst, int cols, const float* __restrict__ in) {
    const int cpr = cols / 4;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n_dst * cpr) return;
    const int d = (int)(g / cpr), c = (int)(g - (long long)d * cpr);
    float4* px = reinterpret_cast<float4*>(x + (size_t)dst_rows[d] * ld) + c;
    float4 a = *px;
    for (int j = rowptr[d]; j < rowptr[d + 1]; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + (size_t)order[j] * cols) + c);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *px = a;
}
template <typename T>
__global__ void halo_unpack_kernel(T* __restrict__ x, int ld, const int32_t* __restrict__ idx, int n, int cols, const T* __restrict__ in) {
    constexpr int V = 16 / sizeof(T);
    const int cpr = cols / V;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * cpr) return;
    const int r = (int)(g / cpr), c = (int)(g - (long long)r * cpr);
    reinterpret_cast<uint4*>(x + (size_t)idx[r] * ld)[c] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)r * cols) + c);
}
}  // namespace

extern "C" int64_t gp_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
    const int64_t nb = num_nodes > 0 ? (num_nodes + kScanBlock - 1) / kScanBlock : 1;
    return 4 * ((int64_t)num_nodes + 1 + nb + num_edges + 16);
}

extern "C" int gp_csr_from_coo(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes, int32_t* perm_dst, int32_t* src_sorted,
                               int32_t* dst_sorted, int32_t* rowptr_dst, int32_t* perm_src, int32_t* rowptr_src, int32_t* att_col,
                               void* workspace, int64_t* prev_edge_index, int32_t* state, void* stream) {
    GP_REQUIRE(num_edges >= 0 && num_edges < (1ll << 31) && num_nodes >= 0, "gp_csr_from_coo: bad sizes E=%lld N=%d", (long long)num_edges,
               num_nodes);
    GP_REQUIRE(rowptr_dst && rowptr_src && workspace, "gp_csr_from_coo: null output");
    GP_REQUIRE(num_edges == 0 || (edge_index && perm_dst && src_sorted && dst_sorted && perm_src), "gp_csr_from_coo: null edge array");
    GP_REQUIRE(num_edges == 0 || (prev_edge_index == nullptr) == (state == nullptr), "gp_csr_from_coo: prev_edge_index and state go together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int E = (int)num_edges, N = num_nodes, tb = 256;
    int32_t* ws = static_cast<int32_t*>(workspace);
    const int32_t* flag = state;
    if (state) {
        // unchanged topology (a replayed training / roll-out step on a static mesh): everything below returns at once
        csr_begin_kernel<<<1, 32, 0, st>>>(state, 0);
        if (E > 0) csr_compare_kernel<<<296, 256, 0, st>>>(edge_index, prev_edge_index, 2ll * E, state);
    }
    int rc = counting_sort<int64_t>(edge_index + E, E, N, rowptr_dst, perm_dst, ws, flag, st);        // by receiver = edge_index[1]
    if (rc) return rc;
    if (E > 0) gather_sorted_kernel<<<(E + tb - 1) / tb, tb, 0, st>>>(edge_index, E, perm_dst, src_sorted, dst_sorted, flag);
    rc = counting_sort<int32_t>(src_sorted, E, N, rowptr_src, perm_src, ws, flag, st);                  // sorted list, by sender
    if (rc) return rc;
    if (att_col && E > 0) gather_i32_kernel<<<(E + tb - 1) / tb, tb, 0, st>>>(dst_sorted, perm_src, E, att_col, flag);
    if (state && E > 0) csr_commit_kernel<<<296, 256, 0, st>>>(edge_index, prev_edge_index, 2ll * E, state);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_halo_pack(const void* x, int32_t ld, int32_t elem_bytes, const int32_t* idx, int32_t n, int32_t cols, void* out,
                            void* stream) {
    if (n <= 0) return 0;
    GP_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "gp_halo_pack: element size must be 2 (bf16) or 4 (fp32)");
    GP_REQUIRE((cols * elem_bytes) % 16 == 0 && (ld * elem_bytes) % 16 == 0, "gp_halo_pack: rows must be multiples of 16 bytes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n * (cols * elem_bytes / 16);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (elem_bytes == 2)
        halo_pack_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(x), ld, idx, n, cols, static_cast<uint16_t*>(out));
    else
        halo_pack_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), ld, idx, n, cols, static_cast<float*>(out));
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_halo_unpack(void* x, int32_t ld, int32_t elem_bytes, const int32_t* idx, int32_t n, int32_t cols, const void* in,
                              void* stream) {
    if (n <= 0) return 0;
    GP_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "gp_halo_unpack: element size must be 2 (bf16) or 4 (fp32)");
    GP_REQUIRE((cols * elem_bytes) % 16 == 0 && (ld * elem_bytes) % 16 == 0, "gp_halo_unpack: rows must be multiples of 16 bytes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n * (cols * elem_bytes / 16);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (elem_bytes == 2)
        halo_unpack_kernel<uint16_t><<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(x), ld, idx, n, cols, static_cast<const uint16_t*>(in));
    else
        halo_unpack_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(x), ld, idx, n, cols, static_cast<const float*>(in));
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gp_halo_unpack_add(float* x, int32_t ld, const int32_t* dst_rows, const int32_t* rowptr, const int32_t* order,
                                  int32_t n_dst, int32_t cols, const float* in, void* stream) {
    if (n_dst <= 0) return 0;
    GP_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "gp_halo_unpack_add: rows must be multiples of 16 bytes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n_dst * (cols / 4);
    halo_unpack_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ld, dst_rows, rowptr, order, n_dst, cols, in);
    GP_CHECK_CUDA(cudaGetLastError());
    return 0;
}
