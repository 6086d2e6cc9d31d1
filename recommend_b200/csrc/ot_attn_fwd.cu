// ot_attn_fwd.cu — causal attention forward with pyramid query pruning, sm_100a tcgen05.
//
// Replaces the einsum / where / softmax / einsum of MixedMHA.call (OT/model.py:101-114): only the
// retained query tail is computed, the [B,H,Lq,Lk] score tensor never exists.
//
// One CTA (128 threads, thread r owns query row r of the tile) walks a static list of work items
// (query tile, head, sample).  Per 128-key block:
//     S  = Q K^T      tcgen05.mma  (M=128, N=128, K=DH)      -> TMEM columns [0,128)
//     online softmax  tcgen05.ld, fp32 exp2, running max / sum in registers
//     P  -> bf16 -> swizzled smem (K-major A operand)
//     PV = P V        tcgen05.mma  (M=128, N=DH,  K=128, V used MN-major as loaded) -> TMEM [128,128+DH)
//     O += PV (registers, rescaled by the running max)
// K/V blocks are TMA-prefetched one block ahead.  256 TMEM columns and <=113 KB smem per CTA let two
// CTAs share an SM, so one CTA's softmax overlaps the other's MMAs.
#include <stdlib.h>
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdKParams {
  int B, H, Lq, Lk, n_qt, total_items;
  float scale, scale_log2;
  __nv_bfloat16* o; long long ldo;
  float* lse;  // [B, H, Lq]
};

template <int DH, int SWB, int KVS>
struct AttnFwdCfg {
  using T = AttnTile<DH, SWB>;
  static constexpr int SMEM_BYTES = T::TILE_BYTES * (1 + 2 * KVS) + PT_BYTES + 256;
  static constexpr int TMEM_COLS = 256;
};

template <int DH, int SWB, int KVS>
__global__ void __launch_bounds__(128, 2)
ot_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnFwdKParams p) {
  using T = AttnTile<DH, SWB>;
  using Cfg = AttnFwdCfg<DH, SWB, KVS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T::TILE_BYTES;                 // [KVS]
  uint8_t* sV = sK + KVS * T::TILE_BYTES;           // [KVS]
  uint8_t* sP = sV + KVS * T::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PT_BYTES);
  uint64_t* bar_q = bars;          // Q tile landed
  uint64_t* bar_kv = bars + 1;     // [KVS] K+V block landed
  uint64_t* bar_s = bars + 3;      // S MMA complete
  uint64_t* bar_pv = bars + 4;     // PV MMA complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    for (int i = 0; i < KVS; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_S = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t t_PV = t_S + 128;

  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);   // V is MN-major

  uint32_t n_items_done = 0;   // phase of bar_q
  uint32_t kv_uses[2] = {0, 0};  // completed uses per K/V stage
  uint32_t blk_count = 0;      // phase of bar_s / bar_pv
  const int off = p.Lk - p.Lq;

  // A CTA takes whole (sample, head) pairs and walks their query tiles back to back, so that the K/V blocks the tiles share are
  // re-read by the same SM within microseconds and come from L2 (round 2: with the tiles of a pair spread over concurrently
  // running CTAs the head_dim-96 forward read 8.5 GB from DRAM for 2.4 GB of unique K/V, profiles/r2_ncu_summary.csv).
  for (int kk = 0;; ++kk) {
    const int bh = blockIdx.x + (kk / p.n_qt) * gridDim.x;
    if (bh >= p.B * p.H) break;
    const int qt = p.n_qt - 1 - (kk % p.n_qt);  // long tiles first
    const int h = bh % p.H;
    const int b = bh / p.H;
    const int q0 = qt * 128;
    const int q_last = min(q0 + 127, p.Lq - 1);
    const int nkv = min((p.Lk + 127) / 128, (off + q_last) / 128 + 1);
    const int pq = off + q0 + tid;  // absolute position of this thread's query

    if (tid == 0) {
      mbar_arrive_expect_tx(bar_q, T::TILE_BYTES);
      load_head_tile<DH, SWB>(sQ, &tmQ, bar_q, h, b, q0);
      mbar_arrive_expect_tx(&bar_kv[0], 2 * T::TILE_BYTES);
      load_head_tile<DH, SWB>(sK, &tmK, &bar_kv[0], h, b, 0);
      load_head_tile<DH, SWB>(sV, &tmV, &bar_kv[0], h, b, 0);
    }

    float m_run = -INFINITY, l_run = 0.0f;
    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.0f;

    for (int j = 0; j < nkv; ++j) {
      const int st = (KVS == 2) ? (j & 1) : 0;
      if (tid == 0) {
        if (KVS == 2 && j + 1 < nkv) {  // prefetch the next block into the other stage (its PV finished last iteration)
          const int ns = st ^ 1;
          mbar_arrive_expect_tx(&bar_kv[ns], 2 * T::TILE_BYTES);
          load_head_tile<DH, SWB>(sK + ns * T::TILE_BYTES, &tmK, &bar_kv[ns], h, b, (j + 1) * 128);
          load_head_tile<DH, SWB>(sV + ns * T::TILE_BYTES, &tmV, &bar_kv[ns], h, b, (j + 1) * 128);
        }
        if (j == 0) mbar_wait(bar_q, n_items_done & 1);
        mbar_wait(&bar_kv[st], kv_uses[st] & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK + st * T::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base, tile_desc_kmajor<DH, SWB>(aQ, kk), tile_desc_kmajor<DH, SWB>(aK, kk), idesc_s, kk != 0);
        umma_commit(bar_s);
      }
      kv_uses[st]++;
      mbar_wait(bar_s, blk_count & 1);
      tc_fence_after();

      // ---- online softmax over the 128 scores of this row ----
      // blocks strictly below the diagonal of the whole tile need no mask (warp-uniform fast path)
      const bool fast = (j * 128 + 127 <= off + q0);
      const int lim = pq - j * 128;        // column c of this block is visible to this row iff c <= lim
      float m_new = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_S + c * 32, v);
        tmem_ld_wait();
        if (fast) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m_new = fmaxf(m_new, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m_new = fmaxf(m_new, (c * 32 + i <= lim) ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);
      const float mb = m_new * p.scale_log2;
      float rowsum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_S + c * 32, v);
        tmem_ld_wait();
        float pr[32];
        if (fast) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mb));
            rowsum += pr[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mb));
            pr[i] = (c * 32 + i <= lim) ? e : 0.0f;
            rowsum += pr[i];
          }
        }
        ptile_store32(sP, tid, c * 32, pr);
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      if (alpha != 1.0f) {   // the running maximum moved: rescale the accumulator
#pragma unroll
        for (int i = 0; i < DH; ++i) o_acc[i] *= alpha;
      }

      fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to tcgen05.mma
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t aP = smem_u32(sP), aV = smem_u32(sV + st * T::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tmem_base + 128, ptile_desc_kmajor(aP, kk), tile_desc_mnmajor<DH, SWB>(aV, kk), idesc_pv, kk != 0);
        umma_commit(bar_pv);
        if (KVS == 1 && j + 1 < nkv) {
          // single K/V stage: the next block can only be fetched once this block's MMAs have read it
          mbar_wait(bar_pv, blk_count & 1);
          mbar_arrive_expect_tx(&bar_kv[0], 2 * T::TILE_BYTES);
          load_head_tile<DH, SWB>(sK, &tmK, &bar_kv[0], h, b, (j + 1) * 128);
          load_head_tile<DH, SWB>(sV, &tmV, &bar_kv[0], h, b, (j + 1) * 128);
        }
      }
      mbar_wait(bar_pv, blk_count & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(t_PV + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(v[i]);
      }
      ++blk_count;
      tc_fence_before();
    }

    // ---- write O (bf16) and the log-sum-exp of the scaled scores ----
    if (q0 + tid < p.Lq) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* orow = p.o + ((long long)(q0 + tid) * p.B + b) * p.ldo + h * DH;
#pragma unroll
      for (int ch = 0; ch < DH / 8; ++ch) {
        uint4 q;
        q.x = pack_bf16x2(o_acc[ch * 8 + 0] * inv, o_acc[ch * 8 + 1] * inv);
        q.y = pack_bf16x2(o_acc[ch * 8 + 2] * inv, o_acc[ch * 8 + 3] * inv);
        q.z = pack_bf16x2(o_acc[ch * 8 + 4] * inv, o_acc[ch * 8 + 5] * inv);
        q.w = pack_bf16x2(o_acc[ch * 8 + 6] * inv, o_acc[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = q;
      }
      p.lse[((long long)b * p.H + h) * p.Lq + q0 + tid] = m_run * p.scale + logf(l_run);
    }
    ++n_items_done;
    __syncthreads();  // every thread is past its last TMEM / smem read before the next item's loads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int DH, int SWB, int KVS>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnFwdKParams& kp, cudaStream_t st) {
  using Cfg = AttnFwdCfg<DH, SWB, KVS>;
  static bool attr_done = false;
  auto kern = ot_attn_fwd_kernel<DH, SWB, KVS>;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int max_ctas = 2 * num_sms();
  const int n_bh = kp.B * kp.H;
  const int grid = n_bh < max_ctas ? n_bh : max_ctas;
  kern<<<grid, 128, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

// rank-3 map over a token-major buffer: dims (cols, B, L), box (slab cols, 1, 128 rows)
int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)B, (uint64_t)L};
  uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)B};
  uint32_t box[3] = {(uint32_t)(swb / 2), 1, 128};
  return make_tmap_bf16(tm, base, 3, dims, str, box, swb);
}

int attn_fwd_ws_impl(const ot_attn_params* p, cudaStream_t st);   // ot_attn_fwd_ws.cu (head_dim 64, round-1 structure)
int attn_fwd_v2_impl(const ot_attn_params* p, cudaStream_t st);   // ot_attn_fwd_v2.cu (head_dim 64, round-2 structure)
int attn_fwd_v3_impl(const ot_attn_params* p, cudaStream_t st);   // ot_attn_fwd_v3.cu (head_dim 64, P in tensor memory)
int attn_fwd_v4_impl(const ot_attn_params* p, cudaStream_t st);   // ot_attn_fwd_v4.cu (head_dim 64, two independent tile streams per CTA)
int attn_fwd_v5_impl(const ot_attn_params* p, cudaStream_t st);   // ot_attn_fwd_v5.cu (head_dim 64, v3 pipeline with a balanced step schedule)

int attn_fwd_impl(const ot_attn_params* p, cudaStream_t st) {
  if (!p || !p->q || !p->k || !p->v || !p->o || !p->lse) OT_FAIL(OT_ERR_INVALID_ARG, "ot_attn_fwd: null pointer");
  if (p->Lq <= 0 || p->Lk < p->Lq || p->B <= 0 || p->H <= 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_attn_fwd: bad sizes B=%d H=%d Lq=%d Lk=%d", p->B, p->H, p->Lq, p->Lk);
  if (p->head_dim != 32 && p->head_dim != 64 && p->head_dim != 96)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_fwd: head_dim=%d (32, 64 and 96 are supported)", p->head_dim);
  if ((p->ldq % 8) || (p->ldk % 8) || (p->ldv % 8) || (p->ldo % 8)) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_fwd: leading dimensions must be multiples of 8");
  const int swb = (p->head_dim == 64 && p->swizzle != 64) ? 128 : 64;
  if (p->head_dim == 64 && swb == 128 && p->swizzle != 128) {   // swizzle=128 forces the simple kernel
    // OT_ATTN_FWD_IMPL (read once) forces one structure for A/B runs: 1 = round-1 warp-specialised kernel, 2 = v2 (P through shared
    // memory), 3 = v3 (P in tensor memory, TS-form P V, tiles 2p / 2p+1 of a (sample, head) in lockstep over a shared K/V stream), 4 = v4
    // (two independent tile streams per CTA), 5 = v5 (v3 with a balanced step schedule).  Unset, by the number of query tiles of a
    // (sample, head), measured on one box (profiles/README.md): one tile -> v4 (v3 / v5 would idle half of the CTA); two tiles -> v3 (the
    // schedules coincide, v5's per-step bookkeeping costs 7-12 %); three or more -> v5 (0.66 vs 0.76 ms at three tiles, 0.94 vs 0.96 at four).
    static const int impl = [] { const char* e = getenv("OT_ATTN_FWD_IMPL"); return e ? atoi(e) : 0; }();
    if (impl == 1) return attn_fwd_ws_impl(p, st);
    if (impl == 2) return attn_fwd_v2_impl(p, st);
    if (impl == 3) return attn_fwd_v3_impl(p, st);
    if (impl == 4) return attn_fwd_v4_impl(p, st);
    if (impl == 5) return attn_fwd_v5_impl(p, st);
    if (p->Lq <= 128) return attn_fwd_v4_impl(p, st);
    return p->Lq <= 256 ? attn_fwd_v3_impl(p, st) : attn_fwd_v5_impl(p, st);
  }
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, swb))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, swb))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, swb))) return rc;
  AttnFwdKParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.total_items = kp.n_qt * p->H * p->B;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.o = (__nv_bfloat16*)p->o; kp.ldo = p->ldo; kp.lse = p->lse;
  if (p->head_dim == 64) {
    if (swb == 128) return launch_attn_fwd<64, 128, 2>(tq, tk, tv, kp, st);
    return launch_attn_fwd<64, 64, 2>(tq, tk, tv, kp, st);
  }
  if (p->head_dim == 32) return launch_attn_fwd<32, 64, 2>(tq, tk, tv, kp, st);   // one 64-byte-swizzle slab per tile
  return launch_attn_fwd<96, 64, 1>(tq, tk, tv, kp, st);
}

}  // namespace ot
