// tmem_bw_probe.cu -- TMEM <-> register bandwidth per SM (tcgen05.ld / tcgen05.st, 32x32b), for sizing the epilogues:
// cycles to move a 128-lane x 128-column fp32 accumulator (64 KB) with 4 warps (one per lane quarter) and with 8
// warps (two per lane quarter, each half of the columns), x16 / x32 shapes.   Run: tmem_bw_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "../tc5.cuh"
using namespace tc5;

template <int NW, int SHAPE, bool STORE>
__global__ void __launch_bounds__(NW * 32, 1) k(long long* out, int reps, float* sink) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tl = tmem_addr(slot, (warp & 3) * 32, 0);
    constexpr int PARTS = NW / 4;                 // warps sharing a lane quarter split the 128 columns
    const int c0 = (warp >> 2) * (128 / PARTS);
    uint32_t acc = 0;
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = tid + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < 128 / PARTS; c += SHAPE) {
            if (STORE) {
                if (SHAPE == 32) { tmem_st16(tl + c0 + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0])); tmem_st16(tl + c0 + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16])); }
                else tmem_st16(tl + c0 + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            } else {
                if (SHAPE == 32) tmem_ld32(tl + c0 + c, v); else tmem_ld16(tl + c0 + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                tmem_ld_wait();
                acc += v[0] + v[SHAPE - 1];
            }
        }
        if (STORE) tmem_st_wait();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    if (sink) sink[tid] = (float)acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int NW, int SHAPE, bool STORE>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 8);
    const int reps = 200;
    k<NW, SHAPE, STORE><<<1, NW * 32>>>(d, reps, nullptr);
    k<NW, SHAPE, STORE><<<1, NW * 32>>>(d, reps, nullptr);
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-44s %8.0f cycles per 64 KB tile  -> %6.1f B/cycle/SM  (%s)\n", name, (double)h / reps, 65536.0 * reps / h, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<4, 16, false>("ld  4 warps, 32x32b.x16, wait per load");
    run<4, 32, false>("ld  4 warps, 32x32b.x32, wait per load");
    run<8, 16, false>("ld  8 warps, 32x32b.x16, wait per load");
    run<8, 32, false>("ld  8 warps, 32x32b.x32, wait per load");
    run<16, 32, false>("ld 16 warps, 32x32b.x32, wait per load");
    run<4, 16, true>("st  4 warps, 32x32b.x16, one wait per tile");
    run<8, 16, true>("st  8 warps, 32x32b.x16, one wait per tile");
    return 0;
}
