// exp_mma_rate.cu — micro-benchmark: tensor-pipe time of the tcgen05.mma sequences the attention kernels issue per 128 x 128
// (query tile, key tile) step, with NO element-wise work around them.  Round-2 question (profiles/README.md): the phase timers of
// ot_attn_bwd_fused show the S/dP products taking ~2300 clk and the dV/dK/dQ products ~2250 clk per step against 512 + 768 clk of
// arithmetic - is that the tensor pipe itself (operand layout / shared-memory operand bandwidth), or the code around it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I recommend_b200/csrc -o /tmp/exp_mma_rate profiles/exp_mma_rate.cu && /tmp/exp_mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ot_common.cuh"
using namespace ot;

// A from TMEM (K-major by definition), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

static constexpr int TILE = 128 * 64 * 2;   // [128 x 64] bf16, 128-byte swizzle
static constexpr int PT = 2 * 128 * 128;    // [128 x 128] bf16, two slabs
static constexpr int SMEM = 4 * TILE + 2 * PT + 1024 + 256;

__global__ void __launch_bounds__(128, 1) k(int mode, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem, *sK = smem + TILE, *sdO = smem + 2 * TILE, *sV = smem + 3 * TILE, *sP = smem + 4 * TILE, *sdS = sP + PT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + PT);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  for (int i = threadIdx.x; i < (4 * TILE + 2 * PT) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);
    constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t idesc_qt = make_idesc_bf16(128, 64, 1, 1);   // dQ with A = dS^T tile used MN-major
    const uint64_t kQ = make_smem_desc<128>(smem_u32(sQ), 16), kK = make_smem_desc<128>(smem_u32(sK), 16);
    const uint64_t kdO = make_smem_desc<128>(smem_u32(sdO), 16), kV = make_smem_desc<128>(smem_u32(sV), 16);
    const uint64_t mQ = make_smem_desc<128>(smem_u32(sQ), 128 * 128), mdO = make_smem_desc<128>(smem_u32(sdO), 128 * 128);
    const uint64_t mK = make_smem_desc<128>(smem_u32(sK), 128 * 128), mV = make_smem_desc<128>(smem_u32(sV), 128 * 128);
    const uint64_t P_mn = make_smem_desc<128>(smem_u32(sP), 128 * 128), dS_mn = make_smem_desc<128>(smem_u32(sdS), 128 * 128);
    const uint64_t P_k0 = make_smem_desc<128>(smem_u32(sP), 16), P_k1 = make_smem_desc<128>(smem_u32(sP) + 128 * 128, 16);
    const uint64_t dS_k0 = make_smem_desc<128>(smem_u32(sdS), 16), dS_k1 = make_smem_desc<128>(smem_u32(sdS) + 128 * 128, 16);
    const uint32_t T_S = tb, T_DP = tb + 128, T_DV = tb + 256, T_DK = tb + 320, T_DQ = tb + 384, T_P = tb + 448;
    auto S = [&]() { for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(T_S, kQ + 2 * kk, kK + 2 * kk, idesc_s, kk != 0); };
    auto DP = [&]() { for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(T_DP, kdO + 2 * kk, kV + 2 * kk, idesc_s, kk != 0); };
    auto DV = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(T_DV, P_mn + 128 * kk, mdO + 128 * kk, idesc_t, 1); };
    auto DK = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(T_DK, dS_mn + 128 * kk, mQ + 128 * kk, idesc_t, 1); };
    auto DQ = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(T_DQ, (kk < 4 ? dS_k0 : dS_k1) + 2 * (kk & 3), mK + 128 * kk, idesc_q, kk != 0); };
    auto DQT = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(T_DQ, dS_mn + 128 * kk, mK + 128 * kk, idesc_qt, kk != 0); };
    auto PV = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(T_DV, (kk < 4 ? P_k0 : P_k1) + 2 * (kk & 3), mV + 128 * kk, idesc_q, 1); };
    auto PV_TS = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(T_DV, T_P + 8 * kk, mV + 128 * kk, idesc_q, 1); };
    auto DV_TS = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(T_DV, T_P + 8 * kk, mdO + 128 * kk, idesc_q, 1); };
    auto DK_TS = [&]() { for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(T_DK, T_DP + 8 * kk, mQ + 128 * kk, idesc_q, 1); };
    // N = 256 products for reference: S over two key tiles at once (B = K tiles 0 and 1 are adjacent: sK, sdO) is not expressible with
    // one descriptor here, so use the GEMM-like shape M 128 x N 256 x K 16 with B = [256 rows x 64] made of sQ..sK (adjacent tiles)
    constexpr uint32_t idesc_256 = make_idesc_bf16(128, 256, 0, 0);
    auto S256 = [&]() { for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(T_S, kdO + 2 * kk, kQ + 2 * kk, idesc_256, kk != 0); };
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      switch (mode) {
        case 0: S(); break;
        case 1: S(); DP(); break;
        case 2: DV(); break;
        case 3: DQ(); break;
        case 4: DV(); DK(); DQ(); break;
        case 5: S(); DP(); DV(); DK(); DQ(); break;
        case 6: PV(); break;
        case 7: PV_TS(); break;
        case 8: DV_TS(); break;
        case 9: S(); DP(); DV_TS(); DK_TS(); DQT(); break;
        case 10: S(); PV_TS(); break;
        case 11: S(); PV(); break;
        case 12: DQT(); break;
        case 13: S256(); break;
        case 14: DK(); break;
        default: break;
      }
      umma_commit(&bars[i & 1]);
      if (i >= 1) mbar_wait(&bars[(i - 1) & 1], ((i - 1) >> 1) & 1);   // one step of slack: the pipe never drains
    }
    mbar_wait(&bars[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const char* names[] = {"S (4 x 128x128x16 SS, K-major A/B)", "S + dP (8)", "dV (8 x 128x64x16, A and B MN-major)", "dQ (8 x 128x64x16, A K-major, B MN-major)",
                         "dV + dK + dQ (24)", "bwd today: S dP dV dK dQ all SS (32)", "PV fwd today (8, A = P smem K-major)", "PV with A = P in TMEM (8)",
                         "dV with A = P^T in TMEM (8)", "bwd TS form: S^T dP^T SS, dV dK TS, dQ SS with A MN-major (32)", "fwd TS form: S + PV(TS)",
                         "fwd today: S + PV (SS)", "dQ with A = dS^T MN-major (8)", "128x256x16 SS x4 (GEMM-like)", "dK (8, A and B MN-major)"};
  const double ideal[] = {256, 512, 256, 256, 768, 1280, 256, 256, 256, 1280, 512, 512, 256, 512, 256};
  const int iters = 2000;
  for (int mode = 0; mode < 15; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 128, SMEM>>>(mode, iters, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
    printf("mode %2d  %7.1f clk/step  (arithmetic %4.0f)  %s\n", mode, s / 148 / iters, ideal[mode], names[mode]);
  }
  return 0;
}
