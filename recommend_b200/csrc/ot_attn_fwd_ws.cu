// ot_attn_fwd_ws.cu — warp-specialised causal attention forward for head_dim 64 (training and full-sequence
// inference path).  Same arithmetic as ot_attn_fwd.cu (OT/model.py:101-114 for the retained query tail), different
// machine mapping:
//   warp 17    loader : walks the (query tile, head, sample) x key-block work list, publishes a step ring in smem,
//                       keeps K/V three blocks ahead and the next item's Q one item ahead through full/empty rings
//   warp 16    MMA    : S = Q K^T of block j+1 is issued into the second S buffer while block j is in softmax;
//                       PV = P V is issued as soon as P(j) is in smem (V used MN-major as loaded)
//   warps 0-15 softmax: thread = (query row, 32 of the 128 key columns): one TMEM read of S, row max exchanged
//                       through smem between the four column quarters, exp2 on the MUFU pipe, P -> bf16 -> swizzled
//                       smem, 16 of the 64 output columns accumulated in registers with the online rescale
// One CTA per SM (TMEM: two S buffers + one PV buffer = 320 columns).
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdWsKParams {
  int B, H, Lq, Lk, n_qt, total_items;
  float scale, scale_log2;
  __nv_bfloat16* o; long long ldo;
  float* lse;  // [B, H, Lq]
  int* sched;  // dynamic work counter (zeroed by the launcher) or NULL = static round-robin
};

static constexpr int FW_THREADS = 576;   // 16 softmax warps + MMA warp + loader warp
static constexpr int FW_DH = 64;

struct AttnFwdWsCfg {
  using T = AttnTile<FW_DH, 128>;
  static constexpr int Q_BUFS = 2;
  static constexpr int KV_STAGES = 3;
  static constexpr int INFO_SLOTS = 8;
  static constexpr int TILES_BYTES = T::TILE_BYTES * (Q_BUFS + 2 * KV_STAGES) + PT_BYTES;
  static constexpr int XCH_BYTES = 2 * 4 * 128 * 4 + 4 * 128 * 4;   // row-max exchange (double-buffered) + row-sum exchange
  static constexpr int SMEM_BYTES = TILES_BYTES + XCH_BYTES + INFO_SLOTS * 32 + 256;
  static constexpr uint32_t T_S0 = 0, T_S1 = 128, T_PV = 256;
};

struct __align__(16) FwdStepInfo {
  int q0, j, b, h;
  int flags;   // bit0 first key block of its item, bit1 last key block of its item, bit2 last step of this CTA, bit3 Q buffer
  int pad0, pad1, pad2;
};
enum { FS_FIRST = 1, FS_LAST = 2, FS_END = 4, FS_QBUF = 8 };

__global__ void __launch_bounds__(FW_THREADS, 1)
ot_attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnFwdWsKParams p) {
  constexpr int DH = FW_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  using Cfg = AttnFwdWsCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                        // [Q_BUFS]
  uint8_t* sK = sQ + Cfg::Q_BUFS * T::TILE_BYTES;            // [KV_STAGES]
  uint8_t* sV = sK + Cfg::KV_STAGES * T::TILE_BYTES;         // [KV_STAGES]
  uint8_t* sP = sV + Cfg::KV_STAGES * T::TILE_BYTES;
  float* s_max = reinterpret_cast<float*>(smem + Cfg::TILES_BYTES);      // [2][4][128]
  float* s_sum = s_max + 2 * 4 * 128;                                     // [4][128]
  FwdStepInfo* info = reinterpret_cast<FwdStepInfo*>(smem + Cfg::TILES_BYTES + Cfg::XCH_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILES_BYTES + Cfg::XCH_BYTES + Cfg::INFO_SLOTS * 32);
  uint64_t* bar_q = bars;             // [2] Q tile landed                        (loader -> MMA)
  uint64_t* bar_qfree = bars + 2;     // [2] last MMA that reads the Q buffer done (MMA commit -> loader)
  uint64_t* bar_kv = bars + 4;        // [3] K/V block landed (+ step info)        (loader -> MMA)
  uint64_t* bar_kvfree = bars + 7;    // [3] PV of the block done                  (MMA commit -> loader)
  uint64_t* bar_s = bars + 10;        // [2] S buffer complete                     (MMA commit -> softmax)
  uint64_t* bar_sfree = bars + 12;    // [2] softmax warps pulled the S buffer out (16 arrivals -> MMA)
  uint64_t* bar_p = bars + 14;        // P tile written                            (16 arrivals -> MMA)
  uint64_t* bar_pv = bars + 15;       // PV complete                               (MMA commit -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_q[i], 1); mbar_init(&bar_qfree[i], 1); mbar_init(&bar_s[i], 1); mbar_init(&bar_sfree[i], 16); }
    for (int i = 0; i < 3; ++i) { mbar_init(&bar_kv[i], 1); mbar_init(&bar_kvfree[i], 1); }
    mbar_init(bar_p, 16);
    mbar_init(bar_pv, 1);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int off = p.Lk - p.Lq;

  if (warp == 17) {
    // ============================== loader warp ==============================
    if (elect_one()) {
      uint32_t t = 0, item_idx = 0;
      // A CTA takes whole (sample, head) pairs and walks their query tiles back to back: the K/V blocks a pair's
      // tiles share are then re-read by the same SM within microseconds and come from L2, not DRAM (profiles/README.md).
      // After its first pair (blockIdx.x) a CTA draws pairs from a device-wide counter (p.sched), so a CTA that starts
      // late takes fewer pairs instead of finishing last; the next pair is fetched while this one is being loaded.
      const int n_bh = p.B * p.H;
      int bh = blockIdx.x;
      while (bh >= 0) {
      int next_bh = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : bh + (int)gridDim.x;
      if (next_bh >= n_bh) next_bh = -1;
      for (int qt = p.n_qt - 1; qt >= 0; --qt, ++item_idx) {   // long tiles first
        const int h = bh % p.H;
        const int b = bh / p.H;
        const int q0 = qt * 128;
        const int q_last = min(q0 + 127, p.Lq - 1);
        const int nkv = min((p.Lk + 127) / 128, (off + q_last) / 128 + 1);
        const int qb = item_idx & 1;
        const bool last_item = (next_bh < 0) && qt == 0;
        if (item_idx >= 2) mbar_wait(&bar_qfree[qb], ((item_idx >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bar_q[qb], T::TILE_BYTES);
        load_head_tile<DH, SWB>(sQ + qb * T::TILE_BYTES, &tmQ, &bar_q[qb], h, b, q0);
        for (int j = 0; j < nkv; ++j, ++t) {
          const int st = t % 3;
          if (t >= 3) mbar_wait(&bar_kvfree[st], ((t / 3) - 1) & 1);
          FwdStepInfo si;
          si.q0 = q0; si.j = j; si.b = b; si.h = h;
          si.flags = (j == 0 ? FS_FIRST : 0) | (j == nkv - 1 ? FS_LAST : 0) | ((last_item && j == nkv - 1) ? FS_END : 0) | (qb ? FS_QBUF : 0);
          si.pad0 = si.pad1 = si.pad2 = 0;
          info[t & (Cfg::INFO_SLOTS - 1)] = si;   // published by the release-arrive on bar_kv below
          mbar_arrive_expect_tx(&bar_kv[st], 2 * T::TILE_BYTES);
          load_head_tile<DH, SWB>(sK + st * T::TILE_BYTES, &tmK, &bar_kv[st], h, b, j * 128);
          load_head_tile<DH, SWB>(sV + st * T::TILE_BYTES, &tmV, &bar_kv[st], h, b, j * 128);
        }
      }
      bh = next_bh;
      }
    }
  } else if (warp == 16) {
    // ============================== MMA warp ==============================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);   // V is MN-major
      const uint64_t tileK = make_smem_desc<SWB>(0, 16);
      const uint64_t tileMN = make_smem_desc<SWB>(0, T::SLAB_BYTES);
      const uint64_t dP0 = make_smem_desc<128>(smem_u32(sP), 16);
      const uint64_t dP1 = make_smem_desc<128>(smem_u32(sP) + PT_SLAB_BYTES, 16);
      auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
      uint32_t n_items = 0;
      auto issue_s = [&](uint32_t t) {   // S(t) = Q K(t)^T into S buffer t&1
        const int st = t % 3;
        mbar_wait(&bar_kv[st], (t / 3) & 1);                        // also publishes info[t]
        const FwdStepInfo si = info[t & (Cfg::INFO_SLOTS - 1)];
        const int qb = (si.flags & FS_QBUF) ? 1 : 0;
        if (si.flags & FS_FIRST) {
          mbar_wait(&bar_q[qb], (n_items >> 1) & 1);
          ++n_items;
        }
        if (t >= 2) mbar_wait(&bar_sfree[t & 1], ((t >> 1) - 1) & 1);   // softmax warps pulled S(t-2) out
        tc_fence_after();
        const uint64_t aQ = tileK + addr14(smem_u32(sQ + qb * T::TILE_BYTES)), aK = tileK + addr14(smem_u32(sK + st * T::TILE_BYTES));
        const uint32_t d = tmem_base + ((t & 1) ? Cfg::T_S1 : Cfg::T_S0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(d, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
        umma_commit(&bar_s[t & 1]);
      };
      issue_s(0);
      uint32_t g = 0;
      bool end = false;
      while (!end) {
        const FwdStepInfo si = info[g & (Cfg::INFO_SLOTS - 1)];
        end = (si.flags & FS_END) != 0;
        if (!end) issue_s(g + 1);                                   // next block's scores run under this block's softmax
        mbar_wait(bar_p, g & 1);                                    // P(g) is in smem, PV(g-1) has been read out
        tc_fence_after();
        const int st = g % 3;
        const uint64_t mV = tileMN + addr14(smem_u32(sV + st * T::TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tmem_base + Cfg::T_PV, (kk < 4 ? dP0 : dP1) + 2 * (kk & 3), mV + 128 * kk, idesc_pv, kk != 0);
        umma_commit(bar_pv);
        umma_commit(&bar_kvfree[st]);
        if (si.flags & FS_LAST) umma_commit(&bar_qfree[(si.flags & FS_QBUF) ? 1 : 0]);
        ++g;
      }
    }
  } else {
    // ============================== softmax warps (0-15) ==============================
    const int quarter = warp >> 2;                 // which 32 of the 128 key columns / which 16 of the 64 output columns
    const int row = (warp & 3) * 32 + lane;        // tile row == TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t g = 0;
    bool end = false;
    float m_run = -INFINITY, l_part = 0.0f;
    float o_acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o_acc[i] = 0.0f;

    while (!end) {
      mbar_wait(&bar_s[g & 1], (g >> 1) & 1);
      tc_fence_after();
      const FwdStepInfo si = info[g & (Cfg::INFO_SLOTS - 1)];
      end = (si.flags & FS_END) != 0;
      const bool first = si.flags & FS_FIRST, last = si.flags & FS_LAST;
      const int q0 = si.q0;
      uint32_t v[32];
      tmem_ld_x32(t_row + ((g & 1) ? Cfg::T_S1 : Cfg::T_S0) + quarter * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_sfree[g & 1]);

      // The causal mask is decided per warp: its 32 rows x 32 columns of the tile are either entirely visible (no mask
      // code at all - every block strictly below the diagonal, and most blocks of a diagonal tile), entirely hidden
      // (nothing to compute) or cut by the diagonal (per-element selects).  Column i of this thread is visible iff
      // i <= lim; lim grows by one per row, so the warp's extremes sit in lanes 0 and 31.
      const int lim = (off + q0 + row) - si.j * 128 - quarter * 32;
      const int lim_first = lim - lane, lim_last = lim_first + 31;
      const bool fast = lim_first >= 31;
      const bool none = lim_last < 0;
      float mx = -INFINITY;
      if (fast) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      } else if (!none) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (i <= lim) ? __uint_as_float(v[i]) : -INFINITY);
      }
      float* xm = s_max + (g & 1) * 512;
      xm[quarter * 128 + row] = mx;
      // the four warps that share these 32 rows (one per column quarter) exchange their maxima; the other twelve
      // softmax warps work on other rows and need not wait here
      named_bar_sync(1 + (warp & 3), 128);
      const float m_new = fmaxf(fmaxf(m_run, fmaxf(xm[row], xm[128 + row])), fmaxf(xm[256 + row], xm[384 + row]));
      // key 0 is visible to every query, so after the first block m_new is finite for every real row
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);
      const float mb = m_new * p.scale_log2;
      float rowsum = 0.0f;
      uint32_t pk[16];
      if (fast) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_mixed(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -mb), 2 * i);
          const float e1 = ex2_mixed(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -mb), 2 * i + 1);
          rowsum += e0 + e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
      } else if (none) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e0 = ex2_mixed(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -mb), 2 * i);
          float e1 = ex2_mixed(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -mb), 2 * i + 1);
          e0 = (2 * i <= lim) ? e0 : 0.0f;
          e1 = (2 * i + 1 <= lim) ? e1 : 0.0f;
          rowsum += e0 + e1;
          pk[i] = pack_bf16x2(e0, e1);
        }
      }
      l_part = l_part * alpha + rowsum;
      m_run = m_new;

      // PV of the previous block: add it before rescaling for this block; its MMA also was the last reader of P
      if (!first) {
        mbar_wait(bar_pv, (g - 1) & 1);
        tc_fence_after();
        uint32_t w[16];
        tmem_ld_x16(t_row + Cfg::T_PV + quarter * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[i] = (o_acc[i] + __uint_as_float(w[i])) * alpha;
      }
      {
        uint8_t* slab = sP + (quarter >> 1) * PT_SLAB_BYTES;
        uint32_t row_v = row;
        asm volatile("" : "+r"(row_v));
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(slab + swz_off<128>(row_v, (quarter & 1) * 4 + ch)) =
              make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);

      if (last) {
        // ---- finish the item: last PV, normalise, write O and the log-sum-exp ----
        mbar_wait(bar_pv, g & 1);
        tc_fence_after();
        uint32_t w[16];
        tmem_ld_x16(t_row + Cfg::T_PV + quarter * 16, w);
        tmem_ld_wait();
        s_sum[quarter * 128 + row] = l_part;
        named_bar_sync(1 + (warp & 3), 128);
        const float l = (s_sum[row] + s_sum[128 + row]) + (s_sum[256 + row] + s_sum[384 + row]);
        if (q0 + row < p.Lq) {
          const float inv = 1.0f / l;
          __nv_bfloat16* orow = p.o + ((long long)(q0 + row) * p.B + si.b) * p.ldo + si.h * DH + quarter * 16;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint4 q;
            q.x = pack_bf16x2((o_acc[ch * 8 + 0] + __uint_as_float(w[ch * 8 + 0])) * inv, (o_acc[ch * 8 + 1] + __uint_as_float(w[ch * 8 + 1])) * inv);
            q.y = pack_bf16x2((o_acc[ch * 8 + 2] + __uint_as_float(w[ch * 8 + 2])) * inv, (o_acc[ch * 8 + 3] + __uint_as_float(w[ch * 8 + 3])) * inv);
            q.z = pack_bf16x2((o_acc[ch * 8 + 4] + __uint_as_float(w[ch * 8 + 4])) * inv, (o_acc[ch * 8 + 5] + __uint_as_float(w[ch * 8 + 5])) * inv);
            q.w = pack_bf16x2((o_acc[ch * 8 + 6] + __uint_as_float(w[ch * 8 + 6])) * inv, (o_acc[ch * 8 + 7] + __uint_as_float(w[ch * 8 + 7])) * inv);
            *reinterpret_cast<uint4*>(orow + ch * 8) = q;
          }
          if (quarter == 0) p.lse[((long long)si.b * p.H + si.h) * p.Lq + q0 + row] = m_run * p.scale + logf(l);
        }
        m_run = -INFINITY;
        l_part = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[i] = 0.0f;
      }
      ++g;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

int attn_fwd_ws_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  AttnFwdWsKParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.total_items = kp.n_qt * p->H * p->B;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.o = (__nv_bfloat16*)p->o; kp.ldo = p->ldo; kp.lse = p->lse;
  kp.sched = sched_slot(st);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnFwdWsCfg::SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms();
  const int n_bh = p->B * p->H;
  const int grid = n_bh < sms ? n_bh : sms;
  ot_attn_fwd_ws_kernel<<<grid, FW_THREADS, AttnFwdWsCfg::SMEM_BYTES, st>>>(tq, tk, tv, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
