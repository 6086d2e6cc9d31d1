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
// Warp roles: warps 0-7 element-wise (thread = one score row x 64 columns); warp 8 issues every tcgen05.mma
// (descriptors hoisted, S/dP of step i+1 issued as soon as the element-wise warps have pulled S/dP of step i out
// of TMEM, so the tensor pipe works under the exp/convert phase); warp 9 walks the work list, publishes a step
// ring in smem and keeps TMA two steps (Q, dO) and one item (K, V) ahead through full/empty mbarrier rings.
#include <stdlib.h>
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
  int dbg;   // profiling experiments only (env OT_DEBUG_ATTN_BWD): bit0 = skip the dQ reductions, bit1 = phase timers
  unsigned long long* dbg_buf;
  int* sched;   // dynamic work counter (zeroed by the launcher) or NULL = static round-robin
};

static constexpr float kLog2eF = 1.4426950408889634f;
static constexpr int FB_THREADS = 448;   // 8 element-wise warps + MMA warp + loader warp + 4 output warps
static constexpr int FB_DH = 64;

struct AttnBwdFusedCfg {
  using T = AttnTile<FB_DH, 128>;
  static constexpr int KV_BUFS = 2;     // K_j/V_j of the next item are fetched while this item runs
  static constexpr int Q_STAGES = 3;    // Q_i/dO_i are fetched two steps ahead (TMA latency stays off the MMA chain)
  static constexpr int INFO_SLOTS = 8;
  static constexpr int TILES_BYTES = T::TILE_BYTES * (2 * KV_BUFS + 2 * Q_STAGES) + 2 * PT_BYTES;
  static constexpr int SMEM_BYTES = TILES_BYTES + INFO_SLOTS * 32 + 256;
  static constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320, T_DQ = 384;
};

__device__ __forceinline__ void red_add_bf16x8(__nv_bfloat16* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One entry of the step ring the loader warp publishes for the MMA warp and the element-wise warps.
struct __align__(16) StepInfo {
  int q0, k0, b, h;
  int flags;        // bit0 first step of its item, bit1 last step of its item, bit2 last step of this CTA, bit3 K/V buffer
  int next_q0;      // query tile of the NEXT step (row statistics are fetched one step ahead), -1 if none
  int next_bh;      // b*H + h of the next step
  int pad;
};
enum { SI_FIRST = 1, SI_LAST = 2, SI_END = 4, SI_KVBUF = 8 };

// Walks the (item, query tile) steps of one CTA in launch order (loader warp only).
struct StepCursor {
  int item, next_item, ii, n_i, i_min, h, b, k0, item_idx;
  bool valid;
  // After its first item (blockIdx.x) a CTA draws items from a device-wide counter (p.sched): a CTA that starts late
  // takes fewer items instead of finishing last.  The following item is fetched when the current one is set up.
  __device__ __forceinline__ void fetch_next(const AttnBwdFusedKParams& p) {
    next_item = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : item + (int)gridDim.x;
  }
  __device__ __forceinline__ void load_item(const AttnBwdFusedKParams& p) {
    valid = item < p.total_items;
    if (!valid) return;
    const int kt = item % p.n_kt;   // early key tiles are seen by the most query tiles
    const int bh = item / p.n_kt;
    h = bh % p.H;
    b = bh / p.H;
    k0 = kt * 128;
    const int off = p.Lk - p.Lq;
    i_min = (k0 - off) < 0 ? 0 : (k0 - off) / 128;
    n_i = p.n_qt - i_min;
    ii = 0;
  }
  // `walker`: the loader's cursor, which goes on to later items (and therefore draws from the counter); everybody else
  // only looks at the CTA's first item
  __device__ __forceinline__ void init(const AttnBwdFusedKParams& p, bool walker) {
    item = blockIdx.x;
    next_item = p.total_items;
    item_idx = 0;
    load_item(p);
    if (walker && valid) fetch_next(p);
  }
  __device__ __forceinline__ void next(const AttnBwdFusedKParams& p) {
    if (++ii == n_i) {
      item = next_item;
      ++item_idx;
      load_item(p);
      if (valid) fetch_next(p);
    }
  }
  __device__ __forceinline__ int q0() const { return (i_min + ii) * 128; }
};

__global__ void __launch_bounds__(FB_THREADS, 1)
ot_attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                         const __grid_constant__ AttnBwdFusedKParams p) {
  constexpr int DH = FB_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  using Cfg = AttnBwdFusedCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                                      // [KV_BUFS]
  uint8_t* sV = sK + Cfg::KV_BUFS * T::TILE_BYTES;         // [KV_BUFS]
  uint8_t* sQ = sV + Cfg::KV_BUFS * T::TILE_BYTES;         // [Q_STAGES]
  uint8_t* sdO = sQ + Cfg::Q_STAGES * T::TILE_BYTES;       // [Q_STAGES]
  uint8_t* sP = sdO + Cfg::Q_STAGES * T::TILE_BYTES;
  uint8_t* sdS = sP + PT_BYTES;
  StepInfo* info = reinterpret_cast<StepInfo*>(smem + Cfg::TILES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::TILES_BYTES + Cfg::INFO_SLOTS * 32);
  uint64_t* bar_kv = bars;            // [2] K_j, V_j landed                     (loader -> MMA)
  uint64_t* bar_kvfree = bars + 2;    // [2] MMAs of the item's last step done   (MMA commit -> loader)
  uint64_t* bar_q = bars + 4;         // [Q_STAGES] Q_i, dO_i landed (+ step info) (loader -> MMA)
  uint64_t* bar_qfree = bars + 7;     // [Q_STAGES] stage released                 (4 output-warp arrivals -> loader)
  uint64_t* bar_s = bars + 10;        // S, dP MMAs complete                     (per step)
  uint64_t* bar_sread = bars + 11;    // E warps pulled S, dP out of TMEM        (per step, 8 arrivals)
  uint64_t* bar_pds = bars + 12;      // P, dS tiles written                     (per step, 8 arrivals)
  uint64_t* bar_d = bars + 13;        // dV, dK, dQp MMAs complete               (per step)
  uint64_t* bar_dqfree = bars + 14;   // output warps pulled dQp out of TMEM     (per step, 4 arrivals)
  uint64_t* bar_accfree = bars + 15;  // output warps pulled dV, dK out          (per item, 4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_kv[i], 1); mbar_init(&bar_kvfree[i], 1); }
    for (int i = 0; i < Cfg::Q_STAGES; ++i) { mbar_init(&bar_q[i], 1); mbar_init(&bar_qfree[i], 4); }
    mbar_init(bar_s, 1);
    mbar_init(bar_sread, 8);
    mbar_init(bar_pds, 8);
    mbar_init(bar_d, 1);
    mbar_init(bar_dqfree, 4);
    mbar_init(bar_accfree, 4);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int off = p.Lk - p.Lq;

  if (warp == 9) {
    // ============================== loader warp: step ring + TMA ==============================
    if (elect_one()) {
      StepCursor c;
      c.init(p, true);
      uint32_t t = 0;
      while (c.valid) {
        const int st = t % Cfg::Q_STAGES;
        const int kb = c.item_idx & 1;
        const bool first = c.ii == 0, last = c.ii == c.n_i - 1;
        if (t >= Cfg::Q_STAGES) mbar_wait(&bar_qfree[st], ((t / Cfg::Q_STAGES) - 1) & 1);   // stage's previous reader done
        if (first && c.item_idx >= 2) mbar_wait(&bar_kvfree[kb], ((c.item_idx >> 1) - 1) & 1);   // buffer's previous item done
        StepCursor n = c;
        n.next(p);
        StepInfo si;
        si.q0 = c.q0(); si.k0 = c.k0; si.b = c.b; si.h = c.h;
        si.flags = (first ? SI_FIRST : 0) | (last ? SI_LAST : 0) | (n.valid ? 0 : SI_END) | (kb ? SI_KVBUF : 0);
        si.next_q0 = n.valid ? n.q0() : -1;
        si.next_bh = n.valid ? n.b * p.H + n.h : 0;
        si.pad = 0;
        info[t & (Cfg::INFO_SLOTS - 1)] = si;     // published by the release-arrive on bar_q below
        if (first) {
          mbar_arrive_expect_tx(&bar_kv[kb], 2 * T::TILE_BYTES);
          load_head_tile<DH, SWB>(sK + kb * T::TILE_BYTES, &tmK, &bar_kv[kb], c.h, c.b, c.k0);
          load_head_tile<DH, SWB>(sV + kb * T::TILE_BYTES, &tmV, &bar_kv[kb], c.h, c.b, c.k0);
        }
        mbar_arrive_expect_tx(&bar_q[st], 2 * T::TILE_BYTES);
        load_head_tile<DH, SWB>(sQ + st * T::TILE_BYTES, &tmQ, &bar_q[st], c.h, c.b, si.q0);
        load_head_tile<DH, SWB>(sdO + st * T::TILE_BYTES, &tmdO, &bar_q[st], c.h, c.b, si.q0);
        c = n;
        ++t;
      }
    }
  } else if (warp == 8) {
    // ============================== MMA warp ==============================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_t = make_idesc_bf16(128, DH, 1, 1);   // P^T dO, dS^T Q: both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, DH, 0, 1);   // dS K: A K-major, B MN-major
      // descriptor bases (the per-k advance is a compile-time constant added to the 14-bit address field)
      const uint64_t dP_mn = make_smem_desc<128>(smem_u32(sP), PT_SLAB_BYTES);
      const uint64_t dS_mn = make_smem_desc<128>(smem_u32(sdS), PT_SLAB_BYTES);
      const uint64_t dS_k0 = make_smem_desc<128>(smem_u32(sdS), 16);
      const uint64_t dS_k1 = make_smem_desc<128>(smem_u32(sdS) + PT_SLAB_BYTES, 16);
      const uint64_t tileK = make_smem_desc<SWB>(0, 16);                // K-major [128 x 64] tile, address added below
      const uint64_t tileMN = make_smem_desc<SWB>(0, T::SLAB_BYTES);    // MN-major use of the same tile
      auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
      uint32_t g = 0;
      uint32_t n_items = 0;      // items whose first step has been issued
      bool end = false;

      auto issue_sdp = [&](uint32_t t) {   // S = Q K^T, dP = dO V^T of step t
        const int st = t % Cfg::Q_STAGES;
        mbar_wait(&bar_q[st], (t / Cfg::Q_STAGES) & 1);           // also publishes info[t]
        const StepInfo si = info[t & (Cfg::INFO_SLOTS - 1)];
        const int kb = (si.flags & SI_KVBUF) ? 1 : 0;
        if (si.flags & SI_FIRST) {
          const uint32_t idx = n_items;                           // per-CTA item index of this item
          mbar_wait(&bar_kv[kb], (idx >> 1) & 1);
          ++n_items;
        }
        if (t >= 1) mbar_wait(bar_sread, (t - 1) & 1);            // S/dP columns free
        tc_fence_after();
        const uint64_t aQ = tileK + addr14(smem_u32(sQ + st * T::TILE_BYTES)), aK = tileK + addr14(smem_u32(sK + kb * T::TILE_BYTES));
        const uint64_t adO = tileK + addr14(smem_u32(sdO + st * T::TILE_BYTES)), aV = tileK + addr14(smem_u32(sV + kb * T::TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base + Cfg::T_S, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_bf16_ss(tmem_base + Cfg::T_DP, adO + 2 * kk, aV + 2 * kk, idesc_s, kk != 0);
        umma_commit(bar_s);
      };

      issue_sdp(0);
      uint32_t items_main = 0;   // items whose first main step has been issued
      while (!end) {
        const StepInfo si = info[g & (Cfg::INFO_SLOTS - 1)];      // visible: bar_q of step g was waited in issue_sdp(g)
        end = (si.flags & SI_END) != 0;
        if (!end) issue_sdp(g + 1);                               // runs under the exp phase of step g
        mbar_wait(bar_pds, g & 1);                                // P, dS of this step are in smem
        if (g > 0) mbar_wait(bar_dqfree, (g - 1) & 1);            // dQp columns free
        if (si.flags & SI_FIRST) {
          if (items_main > 0) mbar_wait(bar_accfree, (items_main - 1) & 1);   // previous dV/dK read out
          ++items_main;
        }
        tc_fence_after();
        const int st = g % Cfg::Q_STAGES;
        const int kb = (si.flags & SI_KVBUF) ? 1 : 0;
        const uint32_t acc = (si.flags & SI_FIRST) ? 0u : 1u;
        const uint64_t mQ = tileMN + addr14(smem_u32(sQ + st * T::TILE_BYTES)), mdO = tileMN + addr14(smem_u32(sdO + st * T::TILE_BYTES));
        const uint64_t mK = tileMN + addr14(smem_u32(sK + kb * T::TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dV[key, e] += sum_q P[q, key] dO[q, e]
          umma_bf16_ss(tmem_base + Cfg::T_DV, dP_mn + 128 * kk, mdO + 128 * kk, idesc_t, (kk != 0) ? 1u : acc);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dK[key, e] += sum_q dS[q, key] Q[q, e]
          umma_bf16_ss(tmem_base + Cfg::T_DK, dS_mn + 128 * kk, mQ + 128 * kk, idesc_t, (kk != 0) ? 1u : acc);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // dQp[q, e] = sum_key dS[q, key] K[key, e]
          umma_bf16_ss(tmem_base + Cfg::T_DQ, (kk < 4 ? dS_k0 : dS_k1) + 2 * (kk & 3), mK + 128 * kk, idesc_q, kk != 0);
        umma_commit(bar_d);
        if (si.flags & SI_LAST) umma_commit(&bar_kvfree[kb]);
        ++g;
      }
    }
  } else if (warp >= 10) {
    // ============================== output warps (10-13) ==============================
    // Every finished [128 x 64] accumulator tile (dQ partial of each step; dV and dK at the end of an item) leaves
    // through these four warps: TMEM -> registers -> bf16 -> swizzled smem staging -> global in full 128-byte rows
    // (row-per-thread stores cost 32 memory transactions per instruction and slowed the whole SM down).  The staging
    // buffer is the Q tile of the step's own pipeline stage, which is dead once bar_d completes; the stage goes back
    // to the loader only when these warps are done with it.  The element-wise warps never wait for their own MMAs.
    const int lgrp = warp & 3;
    const int row = lgrp * 32 + lane;
    const int ot = ((warp - 10) << 5) | lane;         // 0..127 among the output warps
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16);
    uint32_t g = 0;
    bool end = false;
    while (!end) {
      mbar_wait(bar_d, g & 1);                        // dV, dK, dQp MMAs of step g complete
      tc_fence_after();
      const StepInfo si = info[g & (Cfg::INFO_SLOTS - 1)];
      // the Q tile of this step's stage is dead now (every MMA that read it has completed): it is the staging buffer
      uint8_t* sStg = sQ + (g % Cfg::Q_STAGES) * T::TILE_BYTES;
      end = (si.flags & SI_END) != 0;
      const bool last_of_item = si.flags & SI_LAST;
      const int n_pass = last_of_item ? 3 : 1;        // dQp | dV | dK
#pragma unroll 1
      for (int pass = 0; pass < n_pass; ++pass) {
        const uint32_t tcol = pass == 0 ? Cfg::T_DQ : (pass == 1 ? Cfg::T_DV : Cfg::T_DK);
        uint32_t v0[32], v1[32];
        tmem_ld_x32(t_row + tcol, v0);
        tmem_ld_x32(t_row + tcol + 32, v1);
        tmem_ld_wait();
        if (pass == 0 || pass == 2) {                 // dQp / (dV, dK) columns are free again
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pass == 0 ? bar_dqfree : bar_accfree);
        }
        named_bar_sync(2, 128);                       // previous readers of the staging tile are done
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          *reinterpret_cast<uint4*>(sStg + swz_off<128>(row, ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(v0[ch * 8 + 0]), __uint_as_float(v0[ch * 8 + 1])),
                         pack_bf16x2(__uint_as_float(v0[ch * 8 + 2]), __uint_as_float(v0[ch * 8 + 3])),
                         pack_bf16x2(__uint_as_float(v0[ch * 8 + 4]), __uint_as_float(v0[ch * 8 + 5])),
                         pack_bf16x2(__uint_as_float(v0[ch * 8 + 6]), __uint_as_float(v0[ch * 8 + 7])));
          *reinterpret_cast<uint4*>(sStg + swz_off<128>(row, 4 + ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(v1[ch * 8 + 0]), __uint_as_float(v1[ch * 8 + 1])),
                         pack_bf16x2(__uint_as_float(v1[ch * 8 + 2]), __uint_as_float(v1[ch * 8 + 3])),
                         pack_bf16x2(__uint_as_float(v1[ch * 8 + 4]), __uint_as_float(v1[ch * 8 + 5])),
                         pack_bf16x2(__uint_as_float(v1[ch * 8 + 6]), __uint_as_float(v1[ch * 8 + 7])));
        }
        named_bar_sync(2, 128);
        // tile row r <-> query (pass 0) or key (pass 1/2) position base+r; 8 lanes x 16 B cover one output row
        const int base = pass == 0 ? si.q0 : si.k0;
        const int limit = pass == 0 ? p.Lq : p.Lk;
        __nv_bfloat16* gptr = pass == 0 ? p.dq : (pass == 1 ? p.dv : p.dk);
        const long long gld = pass == 0 ? p.lddq : (pass == 1 ? p.lddv : p.lddk);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 16 + (ot >> 3);
          const int ch = ot & 7;
          if (base + r < limit) {
            const uint4 q = *reinterpret_cast<const uint4*>(sStg + swz_off<128>(r, ch));
            __nv_bfloat16* dst = gptr + ((long long)(base + r) * p.B + si.b) * gld + si.h * DH + ch * 8;
            if (pass == 0) {
              if (!(p.dbg & 1)) red_add_bf16x8(dst, q.x, q.y, q.z, q.w);
            } else {
              *reinterpret_cast<uint4*>(dst) = q;
            }
          }
        }
      }
      // hand the Q/dO stage (whose Q tile served as staging) back to the loader
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_qfree[g % Cfg::Q_STAGES]);
      ++g;
    }
  } else {
    // ============================== element-wise warps (0-7) ==============================
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t g = 0;
    // row statistics of the current step, fetched one step ahead so that their latency hides under the math
    float lse_raw = 0.0f, delta = 0.0f;
    {
      StepCursor c0;
      c0.init(p, false);
      if (c0.q0() + row < p.Lq) {
        const long long si0 = ((long long)c0.b * p.H + c0.h) * p.Lq + c0.q0() + row;
        lse_raw = p.lse[si0];
        delta = p.delta[si0];
      }
    }
    bool end = false;
    unsigned long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // phase timers (dbg bit1)
    long long tc0 = clock64();
#define OT_TICK(i) do { if (p.dbg & 2) { const long long t1_ = clock64(); tph[i] += (unsigned long long)(t1_ - tc0); tc0 = t1_; } } while (0)
    while (!end) {
      mbar_wait(bar_s, g & 1);
      OT_TICK(0);   // wait for S/dP
      tc_fence_after();
      const StepInfo si = info[g & (Cfg::INFO_SLOTS - 1)];
      end = (si.flags & SI_END) != 0;
      const int q0 = si.q0, k0 = si.k0;
      const bool row_valid = (q0 + row) < p.Lq;
      const int pq = off + q0 + row;
      const float lse2 = lse_raw * kLog2eF;
      const float dlt = delta;
      if (si.next_q0 >= 0 && si.next_q0 + row < p.Lq) {   // next step's row statistics (used one step later)
        const long long sn = (long long)si.next_bh * p.Lq + si.next_q0 + row;
        lse_raw = p.lse[sn];
        delta = p.delta[sn];
      }
      uint32_t pk[32], dsk[32];   // packed bf16 P and dS of this thread's 64 columns
      // The mask is decided per warp (32 rows x 64 columns of the tile): entirely visible with only valid rows -> no
      // mask code; entirely hidden -> P = dS = 0 without touching TMEM; cut by the diagonal (or the last query rows)
      // -> per-element selects.  Column j of the tile is visible to this row iff j <= lim.
      const int lim = row_valid ? (pq - k0) : -1;
      const int lim_first = off + q0 + (row - lane) - k0;                     // lim of the warp's first row (if it exists)
      const bool fast = (q0 + (row - lane) + 31 < p.Lq) && (lim_first >= half * 64 + 63);
      const bool none = (lim_first + 31 < half * 64) || (q0 + (row - lane) >= p.Lq);
      const float dlt_s = dlt * p.scale;
      if (none) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { pk[i] = 0u; dsk[i] = 0u; }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sread);      // nothing to read: as far as this warp goes S / dP may be reused
      } else {
        uint32_t vs[2][8], vd[2][8];   // double-buffered TMEM reads: chunk c+1 is in flight while chunk c is processed
        tmem_ld_x8(t_row + Cfg::T_S + half * 64, vs[0]);
        tmem_ld_x8(t_row + Cfg::T_DP + half * 64, vd[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cb = c & 1;
          if (c < 7) {
            tmem_ld_x8(t_row + Cfg::T_S + half * 64 + (c + 1) * 8, vs[cb ^ 1]);
            tmem_ld_x8(t_row + Cfg::T_DP + half * 64 + (c + 1) * 8, vd[cb ^ 1]);
          }
          const int j0 = half * 64 + c * 8;
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            float pv0 = ex2_mixed(fmaf(__uint_as_float(vs[cb][2 * i2]), p.scale_log2, -lse2), 2 * i2);
            float pv1 = ex2_mixed(fmaf(__uint_as_float(vs[cb][2 * i2 + 1]), p.scale_log2, -lse2), 2 * i2 + 1);
            if (!fast) {
              pv0 = (j0 + 2 * i2 <= lim) ? pv0 : 0.0f;
              pv1 = (j0 + 2 * i2 + 1 <= lim) ? pv1 : 0.0f;
            }
            const float ds0 = pv0 * fmaf(__uint_as_float(vd[cb][2 * i2]), p.scale, -dlt_s);
            const float ds1 = pv1 * fmaf(__uint_as_float(vd[cb][2 * i2 + 1]), p.scale, -dlt_s);
            pk[c * 4 + i2] = pack_bf16x2(pv0, pv1);
            dsk[c * 4 + i2] = pack_bf16x2(ds0, ds1);
          }
          if (c < 7) tmem_ld_wait();
          if (c == 6) {   // every S/dP column of this thread is in registers: the MMA warp may issue the next S/dP
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sread);
          }
        }
      }
      OT_TICK(1);   // TMEM loads + exp math
      // P/dS tiles are free once the MMAs of the previous step (possibly the previous item's last) have read them
      if (g > 0) {
        mbar_wait(bar_d, (g - 1) & 1);
        tc_fence_after();
      }
      OT_TICK(2);   // wait for the previous step's MMAs
      {
        // this thread's 64 columns are one 128-byte row of slab `half` of the P / dS tiles
        uint8_t* slabP = sP + half * PT_SLAB_BYTES;
        uint8_t* slabD = sdS + half * PT_SLAB_BYTES;
        // recompute the 8 swizzled offsets here (2 ALU ops each): hoisted out of the step loop they get spilled to
        // local memory, which is an L2 round trip in a kernel that leaves L1 almost no capacity
        uint32_t row_v = row;
        asm volatile("" : "+r"(row_v));
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t o = swz_off<128>(row_v, ch);
          *reinterpret_cast<uint4*>(slabP + o) = make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
          *reinterpret_cast<uint4*>(slabD + o) = make_uint4(dsk[ch * 4 + 0], dsk[ch * 4 + 1], dsk[ch * 4 + 2], dsk[ch * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pds);
      OT_TICK(3);   // P/dS stores

      ++g;
    }
    if ((p.dbg & 2) && p.dbg_buf != nullptr && warp == 0 && lane == 0 && blockIdx.x < 8) {
      for (int i = 0; i < 7; ++i) p.dbg_buf[blockIdx.x * 8 + i] = tph[i];
      p.dbg_buf[blockIdx.x * 8 + 7] = g;
    }
#undef OT_TICK
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// delta[b,h,q] = sum_e dO[q,b,h,e] * O[q,b,h,e].  dh/8 consecutive lanes share one (row, head): every lane reads one 16-byte
// chunk of O and of dO (fully coalesced rows) and the partial dot products meet in a shuffle reduction.
template <int LPH>   // lanes per (row, head) = dh / 8: 8 (dh 64), 4 (dh 32), 16 (dh 128)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ d_o, long long lddo,
                  float* __restrict__ delta, int B, int H, int Lq, __nv_bfloat16* __restrict__ dq, long long lddq) {
  const long long total = (long long)Lq * B * H * LPH;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx0 = (long long)blockIdx.x * blockDim.x; idx0 < total; idx0 += stride) {   // warp-uniform trip count
    const long long idx = idx0 + threadIdx.x;
    float acc = 0.0f;
    long long item = idx / LPH;
    const int ch = (int)(idx % LPH);
    const bool live = idx < total;
    if (!live) item = 0;
    const int h = (int)(item % H);
    const long long rowi = item / H;          // = q*B + b
    if (live) {
      const uint4 a = reinterpret_cast<const uint4*>(o + rowi * ldo + h * (LPH * 8))[ch];
      const uint4 g = reinterpret_cast<const uint4*>(d_o + rowi * lddo + h * (LPH * 8))[ch];
      acc = bf16lo(a.x) * bf16lo(g.x) + bf16hi(a.x) * bf16hi(g.x) + bf16lo(a.y) * bf16lo(g.y) + bf16hi(a.y) * bf16hi(g.y) +
            bf16lo(a.z) * bf16lo(g.z) + bf16hi(a.z) * bf16hi(g.z) + bf16lo(a.w) * bf16lo(g.w) + bf16hi(a.w) * bf16hi(g.w);
      // dQ is accumulated with reductions by the main kernel: this pass, which walks the same (row, head, chunk) cells, zeroes it
      // (a separate memset was one more launch, and a 2-D one - strided dQ - ran far below the copy rate)
      reinterpret_cast<uint4*>(dq + rowi * lddq + h * (LPH * 8))[ch] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int off = LPH / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (live && ch == 0) {
      const int b = (int)(rowi % B);
      const int q = (int)(rowi / B);
      delta[((long long)b * H + h) * Lq + q] = acc;
    }
  }
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

// Shared by the head_dim-64 backward kernels: delta = rowsum(dO o O), and dQ zeroed in the same pass.
int attn_bwd_prologue(const ot_attn_params* p, cudaStream_t st) {
  {
    const int lph = p->head_dim / 8;
    const long long total = (long long)p->Lq * p->B * p->H * lph;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
#define OT_LAUNCH_DELTA(N) attn_delta_kernel<N><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)p->o, p->ldo, (const __nv_bfloat16*)p->d_o, \
                                                                             p->lddo, p->delta, p->B, p->H, p->Lq, (__nv_bfloat16*)p->dq, p->lddq)
    if (lph == 8) OT_LAUNCH_DELTA(8); else if (lph == 4) OT_LAUNCH_DELTA(4); else if (lph == 16) OT_LAUNCH_DELTA(16);
    else OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_attn_bwd: head_dim=%d", p->head_dim);
#undef OT_LAUNCH_DELTA
    OT_CUDA_CHECK(cudaGetLastError());
  }
  return OT_OK;
}

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
  kp.sched = sched_slot(st);
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * kLog2eF;
  kp.lse = p->lse; kp.delta = p->delta;
  kp.dq = (__nv_bfloat16*)p->dq; kp.lddq = p->lddq; kp.dk = (__nv_bfloat16*)p->dk; kp.lddk = p->lddk;
  kp.dv = (__nv_bfloat16*)p->dv; kp.lddv = p->lddv;
  kp.dbg = 0;
  kp.dbg_buf = nullptr;
#ifdef OT_ATTN_BWD_DEBUG   // profiling builds only (-DOT_ATTN_BWD_DEBUG): the product launcher neither allocates nor synchronises
  {
    const char* e = getenv("OT_DEBUG_ATTN_BWD");
    kp.dbg = e ? atoi(e) : 0;
    if (kp.dbg & 2) {
      static unsigned long long* buf = nullptr;
      if (!buf) cudaMalloc(&buf, 64 * sizeof(unsigned long long));
      kp.dbg_buf = buf;
    }
  }
#endif

  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdFusedCfg::SMEM_BYTES));
    attr_done = true;
  }
  if ((rc = attn_bwd_prologue(p, st))) return rc;
  const int sms = num_sms();
  const int grid = kp.total_items < sms ? kp.total_items : sms;
  ot_attn_bwd_fused_kernel<<<grid, FB_THREADS, AttnBwdFusedCfg::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], kp);
  OT_CUDA_CHECK(cudaGetLastError());
#ifdef OT_ATTN_BWD_DEBUG
  if (kp.dbg & 2) {   // debugging aid only: synchronises and prints the element-wise warps' phase timers
    unsigned long long h[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, kp.dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* names[7] = {"wait S/dP", "tmem+exp", "wait prev MMAs", "P/dS stores", "-", "-", "-"};
    for (int c = 0; c < 2; ++c) {
      fprintf(stderr, "[attn_bwd cta %d] steps=%llu cycles/step:", c, h[c * 8 + 7]);
      for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)h[c * 8 + i] / (double)(h[c * 8 + 7] ? h[c * 8 + 7] : 1));
      fprintf(stderr, "\n");
    }
  }
#endif
  return OT_OK;
}

}  // namespace ot
