// ot_attn_cached.cu — inference attention with a cross-candidate cache of the sequence-side K/V
// (north_star item 5; PAPER:144-151; the reference's own branch OT/model.py:95-98 is unusable).
//
// One user, C candidates.  The S tokens (behaviour sequences) precede the NS tokens and the mask is causal
// (OT/model.py:109-110, 235), so their K/V are the same for every candidate: they are computed once per user
// and stay in one [Ls, 2d] buffer.  Per candidate only the Tn surviving NS tokens have their own K/V, and only
// the last Tq of them query.  A 128-row query tile packs BC = 128/Tn candidates x Tq tokens:
//     phase 1  S = Q Kshared^T over the Ls shared keys   — one K/V stream for all BC candidates of the tile
//     phase 2  S = Q Kown^T over the BC*Tn own keys       — block-diagonal (same candidate) and causal mask
// Online softmax across both phases, PV on tcgen05 exactly as in ot_attn_fwd.cu.
// Tile row r <-> (token l = r / BC, candidate cc = r % BC); own-key column c <-> (l' = c / BC, cc' = c % BC).
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnCachedKParams {
  int C, H, Tq, Tn, Ls, BC, n_ct, total_items, n_sblk;
  float scale, scale_log2;
  __nv_bfloat16* o; long long ldo;
};

template <int DH, int SWB>
struct AttnCachedCfg {
  using T = AttnTile<DH, SWB>;
  static constexpr int SMEM_BYTES = T::TILE_BYTES * 3 + PT_BYTES + 256;   // Q + one K/V stage + P
};

// rows (l, cc) of a token-major [T*C, ld] buffer: dims (cols, C, T), box (slab cols, BC, T)
template <int DH, int SWB>
__device__ __forceinline__ void load_cand_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int h, int c0) {
  using T = AttnTile<DH, SWB>;
#pragma unroll
  for (int s = 0; s < T::NSLAB; ++s) tma_load_3d(dst + s * T::SLAB_BYTES, tm, bar, h * DH + s * T::SLABC, c0, 0);
}

template <int DH, int SWB>
__global__ void __launch_bounds__(128, 2)
ot_attn_cached_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKo,
                      const __grid_constant__ CUtensorMap tmVo, const __grid_constant__ CUtensorMap tmKs,
                      const __grid_constant__ CUtensorMap tmVs, const __grid_constant__ AttnCachedKParams p) {
  using T = AttnTile<DH, SWB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T::TILE_BYTES;
  uint8_t* sV = sK + T::TILE_BYTES;
  uint8_t* sP = sV + T::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + PT_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* bar_kv = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_pv = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  // The candidate-packed boxes cover only BC*T rows of a 128-row tile: clear the tiles once so that rows TMA never
  // writes hold finite values (a masked probability times a NaN left over in smem would still be NaN; rows left
  // over from an earlier, larger box are real finite data and are masked).
  for (int i = tid; i < (3 * T::TILE_BYTES) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKo); tma_prefetch_desc(&tmVo); tma_prefetch_desc(&tmKs); tma_prefetch_desc(&tmVs);
    mbar_init(bar_q, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_S = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t t_PV = t_S + 128;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);

  const int BC = p.BC;
  const int l_row = tid / BC;            // query token of this thread's row
  const int cc_row = tid - l_row * BC;   // candidate slot of this thread's row
  const int pq = (p.Tn - p.Tq) + l_row;  // position of the query among the own (NS) keys
  const uint32_t q_bytes = (uint32_t)(BC * p.Tq) * SWB * T::NSLAB;
  const uint32_t own_bytes = (uint32_t)(BC * p.Tn) * SWB * T::NSLAB;

  // visibility of the 128 columns of the candidates' own key block for this thread's row, one bit per column (the row
  // and therefore the pattern is the same for every item): column c holds key token c / BC of candidate slot c % BC
  uint32_t own_mask[4] = {0u, 0u, 0u, 0u};
  for (int c = 0; c < 128; ++c) {
    const int lk = c / BC;
    const int cck = c - lk * BC;
    if ((cck == cc_row) && (lk <= pq) && (lk < p.Tn)) own_mask[c >> 5] |= 1u << (c & 31);
  }
  uint32_t n_items_done = 0, kv_count = 0, blk_count = 0;
  const int nblk = p.n_sblk + 1;

  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
    const int ct = item % p.n_ct;
    const int h = item / p.n_ct;
    const int c0 = ct * BC;
    const bool row_valid = (l_row < p.Tq) && (c0 + cc_row < p.C);

    if (tid == 0) {
      mbar_arrive_expect_tx(bar_q, q_bytes);
      load_cand_tile<DH, SWB>(sQ, &tmQ, bar_q, h, c0);
    }
    float m_run = -INFINITY, l_run = 0.0f;
    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.0f;

    for (int j = 0; j < nblk; ++j) {
      const bool shared_blk = j < p.n_sblk;
      if (tid == 0) {
        if (shared_blk) {
          mbar_arrive_expect_tx(bar_kv, 2 * T::TILE_BYTES);
          load_head_tile<DH, SWB>(sK, &tmKs, bar_kv, h, 0, j * 128);
          load_head_tile<DH, SWB>(sV, &tmVs, bar_kv, h, 0, j * 128);
        } else {
          mbar_arrive_expect_tx(bar_kv, 2 * own_bytes);
          load_cand_tile<DH, SWB>(sK, &tmKo, bar_kv, h, c0);
          load_cand_tile<DH, SWB>(sV, &tmVo, bar_kv, h, c0);
        }
        if (j == 0) mbar_wait(bar_q, n_items_done & 1);
        mbar_wait(bar_kv, kv_count & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base, tile_desc_kmajor<DH, SWB>(aQ, kk), tile_desc_kmajor<DH, SWB>(aK, kk), idesc_s, kk != 0);
        umma_commit(bar_s);
      }
      ++kv_count;
      mbar_wait(bar_s, blk_count & 1);
      tc_fence_after();

      auto allowed = [&](int c) -> bool {
        if (shared_blk) return (j * 128 + c) < p.Ls;                // every NS query follows every S key
        return (own_mask[c >> 5] >> (c & 31)) & 1u;                 // same candidate, causal among its NS tokens
      };
      // full blocks of cached sequence keys are visible to every query: no mask code (the common case: all but the last
      // shared block and the block of the candidates' own keys)
      const bool fast = shared_blk && (j * 128 + 127 < p.Ls);
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
          for (int i = 0; i < 32; ++i)
            if (allowed(c * 32 + i)) m_new = fmaxf(m_new, __uint_as_float(v[i]));
        }
      }
      // a row may have no allowed key in a block (e.g. an invalid padding row): keep the running state finite
      const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.0f : ex2_approx((m_run - m_use) * p.scale_log2);
      const float mb = m_use * p.scale_log2;
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
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mb));
            pr[i] = e;
            rowsum += e;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = allowed(c * 32 + i) ? ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -mb)) : 0.0f;
            pr[i] = e;
            rowsum += e;
          }
        }
        ptile_store32(sP, tid, c * 32, pr);
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
#pragma unroll
      for (int i = 0; i < DH; ++i) o_acc[i] *= alpha;

      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tmem_base + 128, ptile_desc_kmajor(aP, kk), tile_desc_mnmajor<DH, SWB>(aV, kk), idesc_pv, kk != 0);
        umma_commit(bar_pv);
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
    if (row_valid) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* orow = p.o + ((long long)l_row * p.C + c0 + cc_row) * p.ldo + h * DH;
#pragma unroll
      for (int ch = 0; ch < DH / 8; ++ch) {
        uint4 q;
        q.x = pack_bf16x2(o_acc[ch * 8 + 0] * inv, o_acc[ch * 8 + 1] * inv);
        q.y = pack_bf16x2(o_acc[ch * 8 + 2] * inv, o_acc[ch * 8 + 3] * inv);
        q.z = pack_bf16x2(o_acc[ch * 8 + 4] * inv, o_acc[ch * 8 + 5] * inv);
        q.w = pack_bf16x2(o_acc[ch * 8 + 6] * inv, o_acc[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = q;
      }
    }
    ++n_items_done;
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

static int make_cand_tmap(CUtensorMap* tm, const void* base, int cols, int C, int Tt, long long ld, int swb, int BC) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)C, (uint64_t)Tt};
  uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)C};
  uint32_t box[3] = {(uint32_t)(swb / 2), (uint32_t)BC, (uint32_t)Tt};
  return make_tmap_bf16(tm, base, 3, dims, str, box, swb);
}

template <int DH, int SWB>
static int launch_attn_cached(const CUtensorMap* tm, const AttnCachedKParams& kp, cudaStream_t st) {
  using Cfg = AttnCachedCfg<DH, SWB>;
  static bool attr_done = false;
  auto kern = ot_attn_cached_kernel<DH, SWB>;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int max_ctas = 2 * num_sms();
  const int grid = kp.total_items < max_ctas ? kp.total_items : max_ctas;
  kern<<<grid, 128, Cfg::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int attn_cached_impl(const ot_attn_cached_params* p, cudaStream_t st) {
  if (!p || !p->q || !p->k_own || !p->v_own || !p->o) OT_FAIL(OT_ERR_INVALID_ARG, "ot_attn_ns_cached_fwd: null pointer");
  if (p->Ls > 0 && (!p->k_shared || !p->v_shared)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_attn_ns_cached_fwd: shared K/V missing");
  if (p->C <= 0 || p->H <= 0 || p->Tq <= 0 || p->Tn < p->Tq || p->Tn > 128 || p->Ls < 0)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_ns_cached_fwd: C=%d H=%d Tq=%d Tn=%d Ls=%d (need 1 <= Tq <= Tn <= 128)", p->C, p->H, p->Tq, p->Tn, p->Ls);
  if (p->head_dim != 32 && p->head_dim != 64 && p->head_dim != 96)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_ns_cached_fwd: head_dim=%d (32, 64 and 96 are supported)", p->head_dim);
  if ((p->ldq % 8) || (p->ld_own_k % 8) || (p->ld_own_v % 8) || (p->ldo % 8) || (p->Ls > 0 && ((p->ld_shared_k % 8) || (p->ld_shared_v % 8))))
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_ns_cached_fwd: leading dimensions must be multiples of 8");
  const int swb = p->head_dim == 64 ? 128 : 64;
  const int cols = p->H * p->head_dim;
  int BC = 128 / p->Tn;
  if (BC > p->C) BC = p->C;
  if (BC > 256) BC = 256;
  CUtensorMap tm[5];
  int rc;
  if ((rc = make_cand_tmap(&tm[0], p->q, cols, p->C, p->Tq, p->ldq, swb, BC))) return rc;
  if ((rc = make_cand_tmap(&tm[1], p->k_own, cols, p->C, p->Tn, p->ld_own_k, swb, BC))) return rc;
  if ((rc = make_cand_tmap(&tm[2], p->v_own, cols, p->C, p->Tn, p->ld_own_v, swb, BC))) return rc;
  if (p->Ls > 0) {
    if ((rc = make_head_tmap(&tm[3], p->k_shared, cols, 1, p->Ls, p->ld_shared_k, swb))) return rc;
    if ((rc = make_head_tmap(&tm[4], p->v_shared, cols, 1, p->Ls, p->ld_shared_v, swb))) return rc;
  } else {
    tm[3] = tm[1]; tm[4] = tm[2];
  }
  AttnCachedKParams kp;
  kp.C = p->C; kp.H = p->H; kp.Tq = p->Tq; kp.Tn = p->Tn; kp.Ls = p->Ls; kp.BC = BC;
  kp.n_ct = (p->C + BC - 1) / BC; kp.total_items = kp.n_ct * p->H; kp.n_sblk = (p->Ls + 127) / 128;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.o = (__nv_bfloat16*)p->o; kp.ldo = p->ldo;
  if (p->head_dim == 64) return launch_attn_cached<64, 128>(tm, kp, st);
  if (p->head_dim == 32) return launch_attn_cached<32, 64>(tm, kp, st);
  return launch_attn_cached<96, 64>(tm, kp, st);
}

}  // namespace ot
