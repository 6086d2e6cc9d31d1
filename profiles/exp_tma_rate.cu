// exp_tma_rate.cu — micro-benchmark: how fast does one SM pull [128 rows x 128 bytes] operand tiles through TMA when every row is its own
// 128-byte segment (the head tiles of the attention kernels: rank-3 map (cols, B, L), box (64, 1, 128), rows B * ld elements apart)?
// Question (profiles/README.md, round 2 second session): the forward v4 kernel loads K/V once per slot (twice the rows of v3) and is slower
// although its slots are balanced - is the TMA row rate a limit?  Data is L2-resident (a few MB walked repeatedly), 4 tiles in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I recommend_b200/csrc -o profiles/exp_tma_rate.bin profiles/exp_tma_rate.cu -lcuda && ./profiles/exp_tma_rate.bin
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ot_common.cuh"
using namespace ot;

static constexpr int TILE = 128 * 128;    // bytes
static constexpr int NBUF = 4;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int mode, int iters, int B, int H, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NBUF * TILE);
  if (threadIdx.x == 0) { for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int b = blockIdx.x % B;
    const long long t0 = clock64();
    for (int i = 0; i < iters + NBUF; ++i) {
      const int s = i % NBUF;
      if (i >= NBUF) mbar_wait(&bars[s], ((i / NBUF) - 1) & 1);
      if (i < iters) {
        mbar_arrive_expect_tx(&bars[s], TILE);
        const int h = i % H, l0 = ((i / H) % 4) * 128;
        if (mode == 0) tma_load_3d(smem + s * TILE, &tm, &bars[s], h * 64, b, l0);          // (cols, B, L): rows B * ld apart
        else tma_load_2d(smem + s * TILE, &tm, &bars[s], h * 64, (b * 4 + (i / H) % 4) * 128);   // plain 2-D: rows ld apart
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int B = 148, H = 4, L = 512, ld = 256;
  __nv_bfloat16* buf;
  cudaMalloc(&buf, (size_t)L * B * ld * 2);
  cudaMemset(buf, 0, (size_t)L * B * ld * 2);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * TILE + 2048);
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tm;
    if (mode == 0) {
      cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)B, (cuuint64_t)L};
      cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * B};
      cuuint32_t box[3] = {64, 1, 128}, es[3] = {1, 1, 1};
      CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode 3d failed %d\n", (int)r); return 1; }
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)L * B};
      cuuint64_t str[1] = {(cuuint64_t)ld * 2};
      cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
      CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode 2d failed %d\n", (int)r); return 1; }
    }
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 128, NBUF * TILE + 2048>>>(tm, mode, iters, B, H, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
    const double clk = s / 148 / iters;
    printf("mode %d (%s): %.1f clk per 16 KB tile = %.1f B/clk/SM, %.2f clk per 128-byte row\n", mode,
           mode == 0 ? "rank-3 head tile, rows B*ld apart" : "rank-2 tile, rows ld apart", clk, TILE / clk, clk / 128);
  }
  return 0;
}
