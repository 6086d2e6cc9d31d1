// ot_attn_bwd_fused.cu — single-pass backward of the pruned causal attention for head_dim 64
// (tape.gradient of OT/model.py:101-114), sm_100a tcgen05, warp-specialised.
//
// One work item = (key tile j, head, sample); the CTA walks the query tiles i that can see key tile j.
// Per (i, j) step, the scores are recomputed ONCE (the two-kernel version in ot_attn_bwd.cu recomputes
// them twice) and all five products run on the tensor pipe:
//     S  = Q_i K_j^T          dP = dO_i V_j^T                       (TMEM [0,128), [128,256))
//     P  = exp2(S*c - lse),   dS = P o (dP - delta) * scale         (8 element-wise warps -> bf16 smem tiles)
//     dV_j += P^T dO_i        dK_j += dS^T Q_i      dQp = dS K_j    (TMEM [256,320), [320,384), [384,448))
// dV_j / dK_j stay in TMEM for the whole item; the dQ partial of every step leaves through vectorised
// bf16x2 reductions (REDG.ADD.BF16x8) into a zero-initialised dQ.
//
// Warp roles: warps 0-7 element-wise (thread = one score row x 64 columns), warp 8 = control (TMA + MMA
// issue by one elected lane).  The control warp issues S/dP of step i+1 as soon as the element-wise warps
// have pulled S/dP of step i out of TMEM, so the tensor pipe works under the exp/convert phase.
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnBwdFusedKParams {
  int B, H, Lq, Lk, n_qt, n_kt, total_items;
  float scale, scale_log2;
  const float* lse;    // [B,H,Lq]
  const float* delta;  // [B,H,Lq]
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
};

static constexpr float kLog2eF = 1.4426950408889634f;
static constexpr int FB_THREADS = 288;
static constexpr int FB_DH = 64;

struct AttnBwdFusedCfg {
  using T = AttnTile<FB_DH, 128>;
  static constexpr int SMEM_BYTES = T::TILE_BYTES * 6 + 2 * PT_BYTES + 256;
  static constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320, T_DQ = 384;
};

__device__ __forceinline__ void red_add_bf16x8(__nv_bfloat16* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(FB_THREADS, 1)
ot_attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                         const __grid_constant__ AttnBwdFusedKParams p) {
  constexpr int DH = FB_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  using Cfg = AttnBwdFusedCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + T::TILE_BYTES;
  uint8_t* sQ = sV + T::TILE_BYTES;          // [2]
  uint8_t* sdO = sQ + 2 * T::TILE_BYTES;     // [2]
  uint8_t* sP = sdO + 2 * T::TILE_BYTES;
  uint8_t* sdS = sP + PT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + PT_BYTES);
  uint64_t* bar_kv = bars;          // K_j, V_j landed                 (per item)
  uint64_t* bar_q = bars + 1;       // [2] Q_i, dO_i landed            (per stage use)
  uint64_t* bar_s = bars + 3;       // S, dP MMAs complete             (per step)
  uint64_t* bar_sread = bars + 4;   // E warps pulled S, dP out of TMEM (per step, 8 arrivals)
  uint64_t* bar_pds = bars + 5;     // P, dS tiles written             (per step, 8 arrivals)
  uint64_t* bar_d = bars + 6;       // dV, dK, dQp MMAs complete       (per step)
  uint64_t* bar_dqfree = bars + 7;  // E warps pulled dQp out of TMEM  (per step, 8 arrivals)
  uint64_t* bar_accfree = bars + 8; // E warps pulled dV, dK out       (per item, 8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(bar_kv, 1);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_sread, 8);
    mbar_init(bar_pds, 8);
    mbar_init(bar_d, 1);
    mbar_init(bar_dqfree, 8);
    mbar_init(bar_accfree, 8);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int off = p.Lk - p.Lq;

  if (warp == 8) {
    // ============================== control warp: TMA + MMA issue ==============================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_t = make_idesc_bf16(128, DH, 1, 1);   // P^T dO, dS^T Q: both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, DH, 0, 1);   // dS K: A K-major, B MN-major
      uint32_t g = 0;               // global step counter of this CTA
      uint32_t q_uses[2] = {0, 0};
      uint32_t n_items = 0;
      auto issue_s_dp = [&](int st) {
        const uint32_t aQ = smem_u32(sQ + st * T::TILE_BYTES), aK = smem_u32(sK);
        const uint32_t adO = smem_u32(sdO + st * T::TILE_BYTES), aV = smem_u32(sV);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base + Cfg::T_S, tile_desc_kmajor<DH, SWB>(aQ, kk), tile_desc_kmajor<DH, SWB>(aK, kk), idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base + Cfg::T_DP, tile_desc_kmajor<DH, SWB>(adO, kk), tile_desc_kmajor<DH, SWB>(aV, kk), idesc_s, kk != 0);
        umma_commit(bar_s);
      };
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++n_items) {
        const int kt = item % p.n_kt;
        const int bh = item / p.n_kt;
        const int h = bh % p.H;
        const int b = bh / p.H;
        const int k0 = kt * 128;
        const int i_min = (k0 - off) < 0 ? 0 : (k0 - off) / 128;
        const int n_i = p.n_qt - i_min;
        // K/V of the previous item are no longer read once its last MMAs completed (bar_d of step g-1)
        if (g > 0) mbar_wait(bar_d, (g - 1) & 1);
        mbar_arrive_expect_tx(bar_kv, 2 * T::TILE_BYTES);
        load_head_tile<DH, SWB>(sK, &tmK, bar_kv, h, b, k0);
        load_head_tile<DH, SWB>(sV, &tmV, bar_kv, h, b, k0);
        mbar_arrive_expect_tx(&bar_q[0], 2 * T::TILE_BYTES);
        load_head_tile<DH, SWB>(sQ, &tmQ, &bar_q[0], h, b, i_min * 128);
        load_head_tile<DH, SWB>(sdO, &tmdO, &bar_q[0], h, b, i_min * 128);
        mbar_wait(bar_kv, n_items & 1);
        mbar_wait(&bar_q[0], q_uses[0] & 1);
        q_uses[0]++;
        if (g > 0) mbar_wait(bar_sread, (g - 1) & 1);   // S/dP columns free
        tc_fence_after();
        issue_s_dp(0);
        for (int ii = 0; ii < n_i; ++ii, ++g) {
          const int st = ii & 1;
          if (ii + 1 < n_i) {
            const int ns = st ^ 1;
            // stage ns was read by the MMAs of step ii-1
            if (ii >= 1) mbar_wait(bar_d, (g - 1) & 1);
            mbar_arrive_expect_tx(&bar_q[ns], 2 * T::TILE_BYTES);
            load_head_tile<DH, SWB>(sQ + ns * T::TILE_BYTES, &tmQ, &bar_q[ns], h, b, (i_min + ii + 1) * 128);
            load_head_tile<DH, SWB>(sdO + ns * T::TILE_BYTES, &tmdO, &bar_q[ns], h, b, (i_min + ii + 1) * 128);
            mbar_wait(bar_sread, g & 1);                // E pulled S/dP of this step
            mbar_wait(&bar_q[ns], q_uses[ns] & 1);
            q_uses[ns]++;
            tc_fence_after();
            issue_s_dp(ns);                             // S/dP of step ii+1 run under the exp phase of step ii
          }
          mbar_wait(bar_pds, g & 1);                    // P, dS of this step are in smem
          if (g > 0) mbar_wait(bar_dqfree, (g - 1) & 1);  // dQp columns free
          if (ii == 0 && n_items > 0) mbar_wait(bar_accfree, (n_items - 1) & 1);  // dV/dK of the previous item read
          tc_fence_after();
          const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
          const uint32_t aQ = smem_u32(sQ + st * T::TILE_BYTES), adO = smem_u32(sdO + st * T::TILE_BYTES), aK = smem_u32(sK);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // dV[key, e] += sum_q P[q, key] dO[q, e]
            umma_bf16_ss(tmem_base + Cfg::T_DV, ptile_desc_mnmajor(aP, kk), tile_desc_mnmajor<DH, SWB>(adO, kk), idesc_t, (ii | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // dK[key, e] += sum_q dS[q, key] Q[q, e]
            umma_bf16_ss(tmem_base + Cfg::T_DK, ptile_desc_mnmajor(adS, kk), tile_desc_mnmajor<DH, SWB>(aQ, kk), idesc_t, (ii | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // dQp[q, e] = sum_key dS[q, key] K[key, e]
            umma_bf16_ss(tmem_base + Cfg::T_DQ, ptile_desc_kmajor(adS, kk), tile_desc_mnmajor<DH, SWB>(aK, kk), idesc_q, kk != 0);
          umma_commit(bar_d);
        }
      }
    }
  } else {
    // ============================== element-wise warps ==============================
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t g = 0;
    uint32_t n_items = 0;

    auto flush_dq = [&](int q0, int b, int h) {
      // dQ partial of the finished step: columns [half*32, +32) of this thread's query row
      uint32_t v[32];
      tmem_ld_x32(t_row + Cfg::T_DQ + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dqfree);
      if (q0 + row < p.Lq) {
        __nv_bfloat16* dst = p.dq + ((long long)(q0 + row) * p.B + b) * p.lddq + h * DH + half * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          red_add_bf16x8(dst + ch * 8, pack_bf16x2(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1])),
                         pack_bf16x2(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3])),
                         pack_bf16x2(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5])),
                         pack_bf16x2(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7])));
      }
    };

    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++n_items) {
      const int kt = item % p.n_kt;
      const int bh = item / p.n_kt;
      const int h = bh % p.H;
      const int b = bh / p.H;
      const int k0 = kt * 128;
      const int i_min = (k0 - off) < 0 ? 0 : (k0 - off) / 128;
      const int n_i = p.n_qt - i_min;
      for (int ii = 0; ii < n_i; ++ii, ++g) {
        const int q0 = (i_min + ii) * 128;
        const bool row_valid = (q0 + row) < p.Lq;
        const int pq = off + q0 + row;
        float lse2 = 0.0f, delta = 0.0f;
        if (row_valid) {
          const long long si = ((long long)b * p.H + h) * p.Lq + q0 + row;
          lse2 = p.lse[si] * kLog2eF;
          delta = p.delta[si];
        }
        mbar_wait(bar_s, g & 1);
        tc_fence_after();
        uint32_t pk[2][16], dk_[2][16];   // packed bf16 P and dS, 2 x 32 columns
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = half * 64 + c * 32;
          uint32_t vs[32], vd[32];
          tmem_ld_x32(t_row + Cfg::T_S + col0, vs);
          tmem_ld_x32(t_row + Cfg::T_DP + col0, vd);
          tmem_ld_wait();
          if (c == 1) {   // S and dP are out of TMEM: the control warp may issue the next S/dP
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sread);
          }
#pragma unroll
          for (int i2 = 0; i2 < 16; ++i2) {
            float pv0 = exp2f(__uint_as_float(vs[2 * i2]) * p.scale_log2 - lse2);
            float pv1 = exp2f(__uint_as_float(vs[2 * i2 + 1]) * p.scale_log2 - lse2);
            if (!row_valid || (k0 + col0 + 2 * i2 > pq)) pv0 = 0.0f;
            if (!row_valid || (k0 + col0 + 2 * i2 + 1 > pq)) pv1 = 0.0f;
            const float ds0 = pv0 * (__uint_as_float(vd[2 * i2]) - delta) * p.scale;
            const float ds1 = pv1 * (__uint_as_float(vd[2 * i2 + 1]) - delta) * p.scale;
            pk[c][i2] = pack_bf16x2(pv0, pv1);
            dk_[c][i2] = pack_bf16x2(ds0, ds1);
          }
        }
        // P/dS tiles are free once the MMAs of the previous step have read them
        if (ii > 0) {
          mbar_wait(bar_d, (g - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = half * 64 + c * 32;
          uint8_t* slabP = sP + (col0 >> 6) * PT_SLAB_BYTES;
          uint8_t* slabD = sdS + (col0 >> 6) * PT_SLAB_BYTES;
          const int ch0 = (col0 & 63) >> 3;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            *reinterpret_cast<uint4*>(slabP + swz_off<128>(row, ch0 + ch)) =
                make_uint4(pk[c][ch * 4 + 0], pk[c][ch * 4 + 1], pk[c][ch * 4 + 2], pk[c][ch * 4 + 3]);
            *reinterpret_cast<uint4*>(slabD + swz_off<128>(row, ch0 + ch)) =
                make_uint4(dk_[c][ch * 4 + 0], dk_[c][ch * 4 + 1], dk_[c][ch * 4 + 2], dk_[c][ch * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
        if (ii > 0) flush_dq(q0 - 128, b, h);   // dQ partial of the previous step (its MMAs completed above)
      }
      // ---- item epilogue: last dQ partial, then dV / dK of this key tile ----
      mbar_wait(bar_d, (g - 1) & 1);
      tc_fence_after();
      flush_dq((i_min + n_i - 1) * 128, b, h);
      const bool key_valid = (k0 + row) < p.Lk;
      uint32_t v[32];
      tmem_ld_x32(t_row + Cfg::T_DV + half * 32, v);
      tmem_ld_wait();
      if (key_valid) {
        __nv_bfloat16* dst = p.dv + ((long long)(k0 + row) * p.B + b) * p.lddv + h * DH + half * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1]));
          q.y = pack_bf16x2(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3]));
          q.z = pack_bf16x2(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5]));
          q.w = pack_bf16x2(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + ch * 8) = q;
        }
      }
      tmem_ld_x32(t_row + Cfg::T_DK + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree);
      if (key_valid) {
        __nv_bfloat16* dst = p.dk + ((long long)(k0 + row) * p.B + b) * p.lddk + h * DH + half * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1]));
          q.y = pack_bf16x2(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3]));
          q.z = pack_bf16x2(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5]));
          q.w = pack_bf16x2(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + ch * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// delta[b,h,q] = sum_e dO[q,b,h,e] * O[q,b,h,e]   (one thread per (row, head))
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ d_o, long long lddo,
                  float* __restrict__ delta, int B, int H, int Lq, int dh) {
  const long long total = (long long)Lq * B * H;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(idx % H);
    const long long rowi = idx / H;          // = q*B + b
    const int b = (int)(rowi % B);
    const int q = (int)(rowi / B);
    const uint4* po = reinterpret_cast<const uint4*>(o + rowi * ldo + h * dh);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + rowi * lddo + h * dh);
    float acc = 0.0f;
    for (int ch = 0; ch < dh / 8; ++ch) {
      const uint4 a = po[ch], g = pd[ch];
      acc += bf16lo(a.x) * bf16lo(g.x) + bf16hi(a.x) * bf16hi(g.x) + bf16lo(a.y) * bf16lo(g.y) + bf16hi(a.y) * bf16hi(g.y) +
             bf16lo(a.z) * bf16lo(g.z) + bf16hi(a.z) * bf16hi(g.z) + bf16lo(a.w) * bf16lo(g.w) + bf16hi(a.w) * bf16hi(g.w);
    }
    delta[((long long)b * H + h) * Lq + q] = acc;
  }
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

int attn_bwd_fused_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_head_tmap(&tm[0], p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tm[1], p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tm[2], p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  if ((rc = make_head_tmap(&tm[3], p->d_o, cols, p->B, p->Lq, p->lddo, 128))) return rc;
  AttnBwdFusedKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk;
  kp.n_qt = (p->Lq + 127) / 128; kp.n_kt = (p->Lk + 127) / 128;
  kp.total_items = kp.n_kt * kp.H * kp.B;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * kLog2eF;
  kp.lse = p->lse; kp.delta = p->delta;
  kp.dq = (__nv_bfloat16*)p->dq; kp.lddq = p->lddq; kp.dk = (__nv_bfloat16*)p->dk; kp.lddk = p->lddk;
  kp.dv = (__nv_bfloat16*)p->dv; kp.lddv = p->lddv;

  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdFusedCfg::SMEM_BYTES));
    attr_done = true;
  }
  // dQ is accumulated with reductions: zero the [Lq*B, H*dh] block it covers (row by row when strided)
  if (p->lddq == cols) {
    OT_CUDA_CHECK(cudaMemsetAsync(p->dq, 0, (size_t)p->Lq * p->B * cols * 2, st));
  } else {
    OT_CUDA_CHECK(cudaMemset2DAsync(p->dq, (size_t)p->lddq * 2, 0, (size_t)cols * 2, (size_t)p->Lq * p->B, st));
  }
  {
    const long long total = (long long)p->Lq * p->B * p->H;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    attn_delta_kernel<<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)p->o, p->ldo, (const __nv_bfloat16*)p->d_o, p->lddo, p->delta,
                                                   p->B, p->H, p->Lq, p->head_dim);
    OT_CUDA_CHECK(cudaGetLastError());
  }
  const int sms = num_sms();
  const int grid = kp.total_items < sms ? kp.total_items : sms;
  ot_attn_bwd_fused_kernel<<<grid, FB_THREADS, AttnBwdFusedCfg::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
