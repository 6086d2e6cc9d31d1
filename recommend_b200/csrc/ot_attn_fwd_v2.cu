// ot_attn_fwd_v2.cu — causal attention forward for head_dim 64, round-2 structure (OT/model.py:101-114 for the retained query
// tail; same arithmetic as ot_attn_fwd_ws.cu, different machine mapping).  What changed and why (profiles/README.md, round 2):
//   * ONE THREAD PER QUERY ROW.  The round-1 kernel spread a row over four warps and exchanged row maxima through shared memory
//     behind a named barrier in every key block; ncu showed it latency-bound (XU pipe 31 %, issue 48 %, tensor 18 %).  Here a
//     softmax thread walks the 128 scores of its row in TMEM twice (maximum, then exponentials; TMEM reads run at > 400 B/clk/SM,
//     profiles/exp_tmem_ld_rate.cu; keeping all 128 in registers spilled) with the next chunk's load under the current chunk's
//     arithmetic and four-way split maximum / sum accumulators - no exchange, no barrier, no 128-deep dependent chain.
//   * TWO QUERY TILES PER CTA IN LOCKSTEP.  Tiles 2p and 2p+1 of one (sample, head) share every K/V block they both need: the
//     block is loaded once, two softmax warpgroups (one per tile) work out of phase, the MMA warp interleaves S = Q K^T of the
//     next block with P V of the current one, so the tensor pipe and the two warpgroups cover each other's latencies.
//   * O STAYS IN TMEM with a lazy rescale: P V accumulates in place; a row rescales its accumulator only when its running maximum
//     has grown by more than 2^8 since the reference it is using (rare after the first block), instead of pulling 64 output columns
//     through the registers in every block.
//   warps 0-3  softmax of tile A (TMEM lanes 0-127 of S_A / O_A)      warps 4-7  softmax of tile B
//   warp 8     MMA issuer                                             warp 9     loader (TMA), publishes the step ring
// TMEM: S_A 0-127, S_B 128-255, O_A 256-319, O_B 320-383.   SMEM: Q 2 items x 2 tiles, K/V 3 stages, P 2 tiles (bf16, swizzled).
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdV2KParams {
  int B, H, Lq, Lk, n_qt, n_pairs;      // n_pairs = ceil(n_qt / 2) tile pairs per (sample, head)
  float scale, scale_log2;
  __nv_bfloat16* o; long long ldo;
  float* lse;  // [B, H, Lq]
  int* sched;
};

static constexpr int F2_THREADS = 320;
static constexpr int F2_DH = 64;
static constexpr int F2_TILE = 128 * F2_DH * 2;          // 16 KB
static constexpr int F2_KV_STAGES = 3;
static constexpr int F2_INFO_SLOTS = 8;
static constexpr int F2_OFF_K = 4 * F2_TILE;                              // Q: [item buffer][tile] = 4 tiles
static constexpr int F2_OFF_V = F2_OFF_K + F2_KV_STAGES * F2_TILE;
static constexpr int F2_OFF_P = F2_OFF_V + F2_KV_STAGES * F2_TILE;
static constexpr int F2_OFF_INFO = F2_OFF_P + 2 * PT_BYTES;
static constexpr int F2_OFF_BARS = F2_OFF_INFO + F2_INFO_SLOTS * 32;
static constexpr int F2_SMEM_BYTES = F2_OFF_BARS + 512;
static_assert(F2_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t F2_T_S = 0, F2_T_O = 256;      // + tile * 128 / + tile * 64
static constexpr float F2_TAU = 8.0f;                    // lazy rescale threshold, log2 units
#ifndef OT_V2_ABLATE
#define OT_V2_ABLATE 0     // timing experiments only (profiles/README.md): 1 no exponentials, 2 no maximum pass, 3 no P stores, 4 neither exp nor sums
#endif

struct __align__(16) F2StepInfo {
  int q0, j, b, h;       // q0 = first query row of tile A
  int flags;
  int pad0, pad1, pad2;
};
enum { F2_FIRST = 1, F2_A = 2, F2_B = 4, F2_LAST_A = 8, F2_LAST_B = 16, F2_END = 32, F2_QBUF = 64 };

__global__ void __launch_bounds__(F2_THREADS, 1)
ot_attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnFwdV2KParams p) {
  constexpr int DH = F2_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + F2_OFF_K;
  uint8_t* sV = smem + F2_OFF_V;
  uint8_t* sP = smem + F2_OFF_P;
  F2StepInfo* info = reinterpret_cast<F2StepInfo*>(smem + F2_OFF_INFO);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2_OFF_BARS);
  uint64_t* bar_q = bars;             // [2] Q tiles of an item landed                        (loader -> MMA)
  uint64_t* bar_qfree = bars + 2;     // [2] every S product of the item is complete          (MMA commit -> loader)
  uint64_t* bar_kv = bars + 4;        // [3] K/V block landed (+ step info)                   (loader -> MMA, softmax)
  uint64_t* bar_kvfree = bars + 7;    // [3] last P V of the step is complete                 (MMA commit -> loader)
  uint64_t* bar_s = bars + 10;        // [2] S of the tile complete                           (MMA commit -> softmax)
  uint64_t* bar_p = bars + 12;        // [2] P written, S read out, O rescaled                (4 arrivals -> MMA)
  uint64_t* bar_o = bars + 14;        // [2] P V of the tile complete                         (MMA commit -> softmax)
  uint64_t* bar_ofree = bars + 16;    // [2] O of a finished tile has been read out           (4 arrivals -> MMA)
  uint64_t* bar_ifull = bars + 18;    // [8] step info published                              (loader -> MMA, softmax)
  uint64_t* bar_ifree = bars + 26;    // [8] step info read by the MMA issuer and the 8 softmax warps (9 arrivals -> loader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q[i], 1); mbar_init(&bar_qfree[i], 1); mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 4);
      mbar_init(&bar_o[i], 1); mbar_init(&bar_ofree[i], 4);
    }
    for (int i = 0; i < F2_KV_STAGES; ++i) { mbar_init(&bar_kv[i], 1); mbar_init(&bar_kvfree[i], 1); }
    for (int i = 0; i < F2_INFO_SLOTS; ++i) { mbar_init(&bar_ifull[i], 1); mbar_init(&bar_ifree[i], 9); }
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int off = p.Lk - p.Lq;

  if (warp == 9) {
    // ============================== loader ==============================
    if (elect_one()) {
      uint32_t t = 0, item_idx = 0;
      const int n_bh = p.B * p.H;
      int bh = blockIdx.x;
      while (bh >= 0) {
        int next_bh = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : bh + (int)gridDim.x;
        if (next_bh >= n_bh) next_bh = -1;
        const int h = bh % p.H, b = bh / p.H;
        for (int pp = p.n_pairs - 1; pp >= 0; --pp, ++item_idx) {      // long pairs first
          const int qtA = 2 * pp;
          const bool hasB = (qtA + 1) < p.n_qt;
          const int q0 = qtA * 128;
          const int lastA = min(q0 + 127, p.Lq - 1);
          const int nkvA = min((p.Lk + 127) / 128, (off + lastA) / 128 + 1);
          const int lastB = min(q0 + 255, p.Lq - 1);
          const int nkvB = hasB ? min((p.Lk + 127) / 128, (off + lastB) / 128 + 1) : 0;
          const int nkv = max(nkvA, nkvB);
          const int qb = item_idx & 1;
          const bool last_item = (next_bh < 0) && pp == 0;
          if (item_idx >= 2) mbar_wait(&bar_qfree[qb], ((item_idx >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&bar_q[qb], (hasB ? 2 : 1) * F2_TILE);
          load_head_tile<DH, SWB>(sQ + (qb * 2) * F2_TILE, &tmQ, &bar_q[qb], h, b, q0);
          if (hasB) load_head_tile<DH, SWB>(sQ + (qb * 2 + 1) * F2_TILE, &tmQ, &bar_q[qb], h, b, q0 + 128);
          for (int j = 0; j < nkv; ++j, ++t) {
            const int st = t % F2_KV_STAGES;
            // The step ring has its own full / free barriers: a warpgroup that is idle in a step (an item whose second tile does
            // not exist) takes no part in the K/V hand-shake of that step and could otherwise be lapped on the K/V barrier.
            const int is = t & (F2_INFO_SLOTS - 1);
            if (t >= F2_INFO_SLOTS) mbar_wait(&bar_ifree[is], ((t / F2_INFO_SLOTS) - 1) & 1);
            F2StepInfo si;
            si.q0 = q0; si.j = j; si.b = b; si.h = h;
            si.flags = (j == 0 ? F2_FIRST : 0) | (j < nkvA ? F2_A : 0) | (j < nkvB ? F2_B : 0) | (j == nkvA - 1 ? F2_LAST_A : 0) |
                       ((hasB && j == nkvB - 1) ? F2_LAST_B : 0) | ((last_item && j == nkv - 1) ? F2_END : 0) | (qb ? F2_QBUF : 0);
            si.pad0 = si.pad1 = si.pad2 = 0;
            info[is] = si;
            mbar_arrive(&bar_ifull[is]);               // release: publishes the slot
            if (t >= F2_KV_STAGES) mbar_wait(&bar_kvfree[st], ((t / F2_KV_STAGES) - 1) & 1);
            mbar_arrive_expect_tx(&bar_kv[st], 2 * F2_TILE);
            load_head_tile<DH, SWB>(sK + st * F2_TILE, &tmK, &bar_kv[st], h, b, j * 128);
            load_head_tile<DH, SWB>(sV + st * F2_TILE, &tmV, &bar_kv[st], h, b, j * 128);
          }
        }
        bh = next_bh;
      }
    }
  } else if (warp == 8) {
    // ============================== MMA issuer ==============================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);     // V is MN-major
      const uint64_t tileK = make_smem_desc<SWB>(0, 16);
      const uint64_t tileMN = make_smem_desc<SWB>(0, T::SLAB_BYTES);
      auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
      uint32_t n_items = 0;
      uint32_t cnt[2] = {0, 0};          // steps issued so far per tile slot (phase counters of bar_p / bar_ofree)
      uint32_t fin[2] = {0, 0};          // finished tiles per slot (phase counter of bar_ofree)
      // S of step t for tile x: waits for the K/V stage (and, on an item's first step, for its Q tiles)
      auto wait_step = [&](uint32_t t) -> F2StepInfo {
        const int st = t % F2_KV_STAGES;
        const int is = t & (F2_INFO_SLOTS - 1);
        mbar_wait(&bar_ifull[is], (t / F2_INFO_SLOTS) & 1);
        const F2StepInfo si = info[is];
        mbar_arrive(&bar_ifree[is]);
        mbar_wait(&bar_kv[st], (t / F2_KV_STAGES) & 1);
        if (si.flags & F2_FIRST) {
          mbar_wait(&bar_q[(si.flags & F2_QBUF) ? 1 : 0], (n_items >> 1) & 1);
          ++n_items;
        }
        return si;
      };
      auto issue_s = [&](const F2StepInfo& si, uint32_t t, int x) {
        const int st = t % F2_KV_STAGES;
        const int qb = (si.flags & F2_QBUF) ? 1 : 0;
        tc_fence_after();
        const uint64_t aQ = tileK + addr14(smem_u32(sQ + (qb * 2 + x) * F2_TILE)), aK = tileK + addr14(smem_u32(sK + st * F2_TILE));
        const uint32_t d = tmem_base + F2_T_S + x * 128;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(d, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
        umma_commit(&bar_s[x]);
      };
      auto issue_pv = [&](const F2StepInfo& si, uint32_t t, int x) {
        const int st = t % F2_KV_STAGES;
        mbar_wait(&bar_p[x], cnt[x] & 1);                               // P(t) of tile x is in smem, O rescaled if needed
        if (si.flags & F2_FIRST) {                                       // first block of a tile overwrites O: the previous tile's O must be out
          if (fin[x] > 0) mbar_wait(&bar_ofree[x], (fin[x] - 1) & 1);
        }
        tc_fence_after();
        const uint64_t dP0 = make_smem_desc<128>(smem_u32(sP + x * PT_BYTES), 16);
        const uint64_t dP1 = make_smem_desc<128>(smem_u32(sP + x * PT_BYTES) + PT_SLAB_BYTES, 16);
        const uint64_t mV = tileMN + addr14(smem_u32(sV + st * F2_TILE));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(tmem_base + F2_T_O + x * DH, (kk < 4 ? dP0 : dP1) + 2 * (kk & 3), mV + 128 * kk, idesc_pv,
                       ((si.flags & F2_FIRST) == 0 || kk != 0) ? 1u : 0u);
        umma_commit(&bar_o[x]);
        ++cnt[x];
        if (si.flags & (x == 0 ? F2_LAST_A : F2_LAST_B)) ++fin[x];
      };
      uint32_t t = 0;
      F2StepInfo cur = wait_step(0);
      if (cur.flags & F2_A) issue_s(cur, 0, 0);
      if (cur.flags & F2_B) issue_s(cur, 0, 1);
      bool end = false;
      while (!end) {
        end = (cur.flags & F2_END) != 0;
        F2StepInfo nxt = cur;
        if (!end) nxt = wait_step(t + 1);                               // next K/V block (and its item's Q) have landed
        // tile A: P V of this step, then S of the next step straight behind it; then the same for tile B
        if (cur.flags & F2_A) issue_pv(cur, t, 0);
        if (!end && (nxt.flags & F2_A)) issue_s(nxt, t + 1, 0);
        if (cur.flags & F2_B) issue_pv(cur, t, 1);
        if (!end && (nxt.flags & F2_B)) issue_s(nxt, t + 1, 1);
        umma_commit(&bar_kvfree[t % F2_KV_STAGES]);                     // every MMA that reads stage t has been issued before this commit
        const bool item_done = end || (nxt.flags & F2_FIRST);
        if (item_done) umma_commit(&bar_qfree[(cur.flags & F2_QBUF) ? 1 : 0]);
        cur = nxt;
        ++t;
      }
    }
  } else {
    // ============================== softmax warpgroups (warps 0-3: tile A, 4-7: tile B) ==============================
    const int x = warp >> 2;                       // tile slot
    const int wrow = (warp & 3) * 32;              // first tile row of this warp == first TMEM lane
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const uint32_t t_s = t_lane + F2_T_S + x * 128, t_o = t_lane + F2_T_O + x * DH;
    const uint32_t act_flag = x == 0 ? F2_A : F2_B, last_flag = x == 0 ? F2_LAST_A : F2_LAST_B;
    uint8_t* myP = sP + x * PT_BYTES;
    uint32_t g = 0;          // global step counter (every step of the CTA)
    uint32_t n = 0;          // steps of this tile slot so far (phase counter of bar_s / bar_o)
    bool end = false;
    float m_ref = -INFINITY, l_run = 0.0f;

    while (!end) {
      const int is = g & (F2_INFO_SLOTS - 1);
      mbar_wait(&bar_ifull[is], (g / F2_INFO_SLOTS) & 1);
      const F2StepInfo si = info[is];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ifree[is]);
      end = (si.flags & F2_END) != 0;
      ++g;
      if (!(si.flags & act_flag)) continue;
      const bool first = si.flags & F2_FIRST, last = si.flags & last_flag;
      const int q0 = si.q0 + x * 128;
      const bool warp_valid = (q0 + wrow) < p.Lq;   // a warp whose 32 rows lie past the end of the query tail only keeps the barriers moving
      mbar_wait(&bar_s[x], n & 1);
      tc_fence_after();
      // column i of this block is visible to this row iff i <= lim (causal mask aligned to the sequence tail, OT/model.py:64,109)
      const int lim = (off + q0 + row) - si.j * 128;
      const int lim_lo = (off + q0 + wrow) - si.j * 128;           // lane 0; lane 31 has lim_lo + 31
      if (warp_valid) {
        // ---- pass 1: row maximum (four independent partial maxima: a single running maximum is a 128-deep dependent chain).
        // Chunk c (32 columns) is, for the whole warp, hidden (vis == 0), cut by the diagonal (1) or fully visible (2). ----
        int vis[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) vis[c] = (lim_lo + 31 < c * 32) ? 0 : (lim_lo >= c * 32 + 31) ? 2 : 1;
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#if OT_V2_ABLATE == 2
        mx4[0] = 8.0f;
        if (false)
#endif
        {
          uint32_t va[32], vb[32];
          tmem_ld_x32(t_s, va);                                     // chunk 0 always holds key 0 .. visible to someone
          if (vis[1]) tmem_ld_x32(t_s + 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            // chunks (0, 1) on the first trip, (2, 3) on the second; the loads of the second pair fly under the first pair's math
            const int c0 = half * 2, c1 = half * 2 + 1;
            if (vis[c0] == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(va[i]));
            } else if (vis[c0] == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], (c0 * 32 + i <= lim) ? __uint_as_float(va[i]) : -INFINITY);
            }
            if (half == 0 && vis[2]) tmem_ld_x32(t_s + 64, va);
            if (vis[c1] == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(vb[i]));
            } else if (vis[c1] == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], (c1 * 32 + i <= lim) ? __uint_as_float(vb[i]) : -INFINITY);
            }
            if (half == 0) {
              if (vis[3]) tmem_ld_x32(t_s + 96, vb);
              tmem_ld_wait();
            }
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // ---- lazy rescale: keep the old reference maximum unless the new maximum exceeds it by more than 2^TAU ----
        // (key 0 is visible to every query, so mx is finite in the first block of a tile)
        bool need = false;
        float m_new = m_ref;
        if (first) { m_new = mx; }
        else if ((mx - m_ref) * p.scale_log2 > F2_TAU) { m_new = mx; need = true; }
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&bar_o[x], (n - 1) & 1);                       // P V of the previous block has landed in O
          tc_fence_after();
          const float alpha = need ? ex2_approx((m_ref - m_new) * p.scale_log2) : 1.0f;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {        // 8 columns at a time: the 128 scores of the row are live in registers here
            uint32_t w[8];
            tmem_ld_x8(t_o + c * 8, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_o + c * 8), "r"(w[0]),
                         "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                         : "memory");
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          l_run *= alpha;
        }
        m_ref = m_new;
        // ---- pass 2: p = 2^((s - m_ref) * scale * log2 e), row sum, bf16 P tile ----
        const float mb = m_ref * p.scale_log2;
        float rs4[4] = {0.0f, 0.0f, 0.0f, 0.0f};                    // independent partial row sums (no 128-deep add chain)
        {
          uint32_t vv[2][32];
          tmem_ld_x32(t_s, vv[0]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t (&v)[32] = vv[c & 1];
            if (c < 3 && vis[c + 1]) tmem_ld_x32(t_s + (c + 1) * 32, vv[(c + 1) & 1]);    // next chunk's load under this chunk's math
            uint32_t pk[16];
            if (vis[c] == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            } else {
              // Three phases over the chunk's 32 values - arguments, exponentials, sums / packs - instead of one fused loop: with two
              // softmax warps per scheduler nothing hides the MUFU latency (~20 clk) if a result is consumed two instructions after
              // it was issued, which is how ptxas scheduled the fused form (ncu: the hot stall was `wait` on the MUFU line).
              float e[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = fmaf(__uint_as_float(v[i]), p.scale_log2, -mb);
#if OT_V2_ABLATE != 1 && OT_V2_ABLATE != 4
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = ex2_mixed(e[i], i);
#endif
              if (vis[c] == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) e[i] = (c * 32 + i <= lim) ? e[i] : 0.0f;
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
#if OT_V2_ABLATE != 4
                rs4[i & 3] += e[2 * i] + e[2 * i + 1];
#endif
                pk[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
              }
            }
            uint8_t* slab = myP + (c >> 1) * PT_SLAB_BYTES;
#if OT_V2_ABLATE == 3
            if (pk[0] == 0x12345678u)
#endif
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              *reinterpret_cast<uint4*>(slab + swz_off<128>(row, (c & 1) * 4 + ch)) =
                  make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
            if (c < 3) tmem_ld_wait();
          }
        }
        const float rowsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        l_run += rowsum;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[x]);

      if (last) {
        // ---- finish the tile: last P V, normalise, write O and the log-sum-exp ----
        mbar_wait(&bar_o[x], n & 1);
        tc_fence_after();
        if (warp_valid) {
          const bool row_ok = (q0 + row) < p.Lq;
          const float inv = 1.0f / l_run;
          __nv_bfloat16* orow = p.o + ((long long)(q0 + row) * p.B + si.b) * p.ldo + si.h * DH;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t w[32];
            tmem_ld_x32(t_o + c * 32, w);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                uint4 q;
                q.x = pack_bf16x2(__uint_as_float(w[ch * 8 + 0]) * inv, __uint_as_float(w[ch * 8 + 1]) * inv);
                q.y = pack_bf16x2(__uint_as_float(w[ch * 8 + 2]) * inv, __uint_as_float(w[ch * 8 + 3]) * inv);
                q.z = pack_bf16x2(__uint_as_float(w[ch * 8 + 4]) * inv, __uint_as_float(w[ch * 8 + 5]) * inv);
                q.w = pack_bf16x2(__uint_as_float(w[ch * 8 + 6]) * inv, __uint_as_float(w[ch * 8 + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + c * 32 + ch * 8) = q;
              }
            }
          }
          if (row_ok) p.lse[((long long)si.b * p.H + si.h) * p.Lq + q0 + row] = m_ref * p.scale + logf(l_run);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_ofree[x]);
        m_ref = -INFINITY;
        l_run = 0.0f;
      }
      ++n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

int attn_fwd_v2_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  AttnFwdV2KParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.n_pairs = (kp.n_qt + 1) / 2;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.o = (__nv_bfloat16*)p->o; kp.ldo = p->ldo; kp.lse = p->lse;
  kp.sched = sched_slot(st);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms();
  const int n_bh = p->B * p->H;
  const int grid = n_bh < sms ? n_bh : sms;
  ot_attn_fwd_v2_kernel<<<grid, F2_THREADS, F2_SMEM_BYTES, st>>>(tq, tk, tv, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
