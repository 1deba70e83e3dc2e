// tma_probe.cu -- bring-up test for the TMA path: a tiled tensor map (box 64x128, SWIZZLE_128B) over a
// row-major bf16 matrix must land in shared memory in exactly the SW128 row-tile layout of tc5.cuh,
// zero-fill rows past the end, and store back clipped.  Run: tma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "../tc5.cuh"
using namespace tc5;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap in_map,
                                                    const __grid_constant__ CUtensorMap out_map, int H, int row0,
                                                    uint16_t* img_out /* raw smem image, H/64 blocks x 16 KB */) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    const int nblk = (H + 63) / 64;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, nblk * 128 * 128);
        for (int b = 0; b < nblk; ++b) tma_load_2d(smem_u32(smem) + b * 16384, &in_map, b * 64, row0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nblk * 8192; i += blockDim.x) img_out[i] = reinterpret_cast<uint16_t*>(smem)[i];
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int b = 0; b < nblk; ++b) tma_store_2d(&out_map, b * 64, row0, smem_u32(smem) + b * 16384);
        tma_store_commit();
        tma_store_wait_all();
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    int fails = 0;
    setvbuf(stdout, nullptr, _IONBF, 0);
    for (int H : {128, 64, 32}) {
        const int rows = 200, row0 = 128;          // second tile is ragged: rows 200..255 do not exist
        std::vector<uint16_t> in((size_t)rows * H), out((size_t)rows * H, 0xFFFF);
        for (size_t i = 0; i < in.size(); ++i) in[i] = (uint16_t)(i * 7 + 3);
        uint16_t *d_in, *d_out, *d_img;
        const int nblk = (H + 63) / 64;
        CK(cudaMalloc(&d_in, in.size() * 2)); CK(cudaMalloc(&d_out, out.size() * 2)); CK(cudaMalloc(&d_img, nblk * 16384));
        CK(cudaMemcpy(d_in, in.data(), in.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_out, 0xFF, out.size() * 2));
        CUtensorMap im, om;
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)H * 2};
        cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};   // H=32: columns 32..63 are out of bounds -> zero fill, 128-B rows kept
        CUresult r1 = encode(&im, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&om, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("H=%d encode failed %d %d\n", H, (int)r1, (int)r2); ++fails; continue; }
        CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
        tma_kernel<<<1, 128, 40 * 1024>>>(im, om, H, row0, d_img);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        std::vector<uint16_t> img((size_t)nblk * 8192);
        CK(cudaMemcpy(img.data(), d_img, img.size() * 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost));
        int bad_img = 0, bad_out = 0;
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < H; ++c) {
                const int blk = c >> 6, cc = c & 63;
                const size_t off = (size_t)blk * 8192 + (size_t)r * 64 + (((cc >> 3) ^ (r & 7)) << 3) + (cc & 7);   // in uint16 units
                const uint16_t expect = (row0 + r < rows) ? in[(size_t)(row0 + r) * H + c] : 0;
                if (img[off] != expect) ++bad_img;
            }
        fflush(stdout);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < H; ++c) {
                const uint16_t expect = (r >= row0) ? in[(size_t)r * H + c] : 0xFFFF;     // only the tile's rows were stored
                if (out[(size_t)r * H + c] != expect) ++bad_out;
            }
        printf("H=%3d  smem image vs SW128 layout: %s (%d bad)   store back: %s (%d bad)\n", H, bad_img ? "FAIL" : "PASS", bad_img,
               bad_out ? "FAIL" : "PASS", bad_out);
        fails += (bad_img != 0) + (bad_out != 0);
    }
    return fails ? 3 : 0;
}
