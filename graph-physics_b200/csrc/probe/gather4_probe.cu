// gather4_probe.cu -- bring-up test for TMA row gathers (cp.async.bulk.tensor.2d ... tile::gather4): four
// indexed rows of a row-major bf16 matrix per instruction, landing in the SW128 row-tile layout of tc5.cuh.
// A 128-row tile of gathered rows = 32 lanes x 2 column blocks = 64 instructions, no per-thread address
// arithmetic.  The tensor map needs a one-row box, {64, 1}; a {64, 4} box faults (illegal instruction).
// Run: gather4_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "../tc5.cuh"
using namespace tc5;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const void* tmap, int32_t col, int32_t r0, int32_t r1, int32_t r2,
                                            int32_t r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) gather_kernel(const __grid_constant__ CUtensorMap map, const int* idx /*[128]*/, int H,
                                                        uint16_t* img_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    const int nblk = H / 64;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, nblk * 16384);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int l = threadIdx.x;
        const int4 r = *reinterpret_cast<const int4*>(idx + 4 * l);
        for (int b = 0; b < nblk; ++b)
            tma_gather4(smem_u32(smem) + b * 16384 + l * 512, &map, b * 64, r.x, r.y, r.z, r.w, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nblk * 8192; i += blockDim.x) img_out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    const int N = 1000, H = 128;
    std::vector<uint16_t> src((size_t)N * H);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (uint16_t)(i * 13 + 5);
    std::vector<int> idx(128);
    for (int i = 0; i < 128; ++i) idx[i] = (i * 37 + 11) % N;
    uint16_t *d_src, *d_img; int* d_idx;
    CK(cudaMalloc(&d_src, src.size() * 2)); CK(cudaMalloc(&d_img, 2 * 16384)); CK(cudaMalloc(&d_idx, 512));
    CK(cudaMemcpy(d_src, src.data(), src.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_idx, idx.data(), 512, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    int ok_any = 0;
    for (int box_rows : {1}) {
        CUtensorMap m;
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)N}, strides[1] = {(cuuint64_t)H * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
        CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box {64,%d}: encode failed (%d)\n", box_rows, (int)r); continue; }
        CK(cudaMemset(d_img, 0xEE, 2 * 16384));
        gather_kernel<<<1, 128, 40 * 1024>>>(m, d_idx, H, d_img);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("box {64,%d}: kernel failed: %s\n", box_rows, cudaGetErrorString(e)); return 3; }
        std::vector<uint16_t> img(2 * 8192);
        CK(cudaMemcpy(img.data(), d_img, img.size() * 2, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r_ = 0; r_ < 128; ++r_)
            for (int c = 0; c < H; ++c) {
                const int blk = c >> 6, cc = c & 63;
                const size_t off = (size_t)blk * 8192 + (size_t)r_ * 64 + (((cc >> 3) ^ (r_ & 7)) << 3) + (cc & 7);
                if (img[off] != src[(size_t)idx[r_] * H + c]) ++bad;
            }
        printf("box {64,%d}: gathered tile vs SW128 layout: %s (%d bad of %d)\n", box_rows, bad ? "FAIL" : "PASS", bad, 128 * H);
        ok_any |= (bad == 0);
    }
    return ok_any ? 0 : 3;
}
