// ot_attn_fwd_v3.cu — causal attention forward for head_dim 64 with the probabilities in TENSOR MEMORY (OT/model.py:101-114 for
// the retained query tail; same arithmetic as ot_attn_fwd_v2.cu).  What ncu said about v2 and the first v3 build
// (profiles/README.md, round 2): the two softmax warps of a scheduler are the critical path, and 60 % of their time was NOT the
// exponentials - 28 % tile epilogue (wait for the last P V, 64-column read-out, row-per-thread global stores), 17 % step
// prologue, 16 % P hand-over, with the XU pipe idle meanwhile.  Structure now:
//   * a softmax thread pulls its whole score row (up to 128 fp32) out of TMEM once, releases the S columns at once (the next S
//     product runs under this block's exponentials) and writes P back to TMEM as packed bf16 pairs (tcgen05.st), chunk by chunk;
//   * P V is issued with A = P in tensor memory (TS form): no P tile in shared memory, no swizzled stores, no proxy fence;
//   * S and P V cover only the key columns some row of the tile can see (N / K extent in steps of 16);
//   * a separate EPILOGUE warpgroup finishes tiles: it waits for the last P V, scales O by 1 / l, stages the bf16 tile in shared
//     memory and stores it with one TMA (full 128-byte rows), writes the log-sum-exp - the softmax warpgroup hands over the row
//     statistics through shared memory and goes straight on to its next tile.
// Warps: 0 loader, 1 MMA issuer (2-3 idle), 4-7 epilogue, 8-11 softmax of tile A, 12-15 softmax of tile B (tiles 2p and 2p+1 of
// one (sample, head) share every K/V block).  Registers re-balanced with setmaxnreg (40 / 72 / 200).
// TMEM: S_A 0, S_B 128, O_A 256, O_B 320, P_A 384, P_B 448 (64 columns hold 128 bf16 probabilities per row).
// SMEM: Q 2 items x 2 tiles, K/V 4 stages, one 16 KB output staging tile, row statistics, step ring, mbarriers.
#include "ot_attn_fwd_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdV3KParams {
  int B, H, Lq, Lk, n_qt, n_pairs;      // n_pairs = ceil(n_qt / 2) tile pairs per (sample, head)
  float scale, scale_log2;
  float* lse;  // [B, H, Lq]
  int* sched;
};

static constexpr int F3_THREADS = 512;
static constexpr int F3_CTRL_REGS = 40, F3_EPI_REGS = 72, F3_SM_REGS = 200;   // 128 x (128-40) + 128 x (128-72) released = 256 x (200-128) taken
static constexpr int F3_DH = 64;
static constexpr int F3_TILE = 128 * F3_DH * 2;          // 16 KB
static constexpr int F3_KV_STAGES = 4;
static constexpr int F3_INFO_SLOTS = 8;
static constexpr int F3_OFF_K = 4 * F3_TILE;                              // Q: [item buffer][tile] = 4 tiles
static constexpr int F3_OFF_V = F3_OFF_K + F3_KV_STAGES * F3_TILE;
static constexpr int F3_OFF_STG = F3_OFF_V + F3_KV_STAGES * F3_TILE;      // output staging tile
static constexpr int F3_OFF_STATS = F3_OFF_STG + F3_TILE;                 // [slot][buffer][128] float2 (l, m)
static constexpr int F3_OFF_INFO = F3_OFF_STATS + 2 * 2 * 128 * 8;
static constexpr int F3_OFF_BARS = F3_OFF_INFO + F3_INFO_SLOTS * 16;
static constexpr int F3_SMEM_BYTES = F3_OFF_BARS + 512;
static_assert(F3_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t F3_T_S = 0, F3_T_O = 256, F3_T_P = 384;     // + tile * 128 / + tile * 64 / + tile * 64
// mbarrier byte offsets inside the barrier block
enum : uint32_t {
  B3_Q = 0,          // [2] Q tiles of an item landed                        (loader -> MMA)
  B3_QFREE = 16,     // [2] every S product of the item is complete          (MMA commit -> loader)
  B3_KV = 32,        // [4] K/V block landed                                 (loader -> MMA)
  B3_KVFREE = 64,    // [4] every product that reads the stage is complete   (MMA commit -> loader)
  B3_S = 96,         // [2] S of the tile complete                           (MMA commit -> softmax)
  B3_SFREE = 112,    // [2] S pulled into registers                          (4 arrivals -> MMA)
  B3_P = 128,        // [2] P written to TMEM, O rescaled                    (4 arrivals -> MMA)
  B3_O = 144,        // [2] P V of the tile complete                         (MMA commit -> softmax, epilogue)
  B3_OFREE = 160,    // [2] O of a finished tile has been read out           (4 epilogue arrivals -> MMA)
  B3_STATS = 176,    // [2] row statistics of a finished tile written        (4 softmax arrivals -> epilogue)
  B3_IFULL = 192,    // [8] step info published                              (loader -> everybody)
  B3_IFREE = 256,    // [8] step info read: MMA issuer + 8 softmax warps + 4 epilogue warps = 13 arrivals -> loader
  B3_TMEM = 320
};
enum { F3_FIRST = 1, F3_A = 2, F3_B = 4, F3_LAST_A = 8, F3_LAST_B = 16, F3_END = 32, F3_QBUF = 64 };
// step info, 16 bytes: x = q0 | j << 16, y = b, z = h | flags << 8, w = nA | nB << 8
//   nA / nB: key columns of this block that some row of tile A / B can see, rounded up to 16 (0: tile idle in this step)

__global__ void __launch_bounds__(F3_THREADS, 1)
ot_attn_fwd_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ AttnFwdV3KParams p) {
  constexpr int DH = F3_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + F3_OFF_BARS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((sbase & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    auto init = [&](uint32_t off, int n, uint32_t count) {
      for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + off + 8 * i), "r"(count));
    };
    init(B3_Q, 2, 1); init(B3_QFREE, 2, 1); init(B3_KV, 4, 1); init(B3_KVFREE, 4, 1); init(B3_S, 2, 1); init(B3_SFREE, 2, 4);
    init(B3_P, 2, 4); init(B3_O, 2, 1); init(B3_OFREE, 2, 4); init(B3_STATS, 2, 4); init(B3_IFULL, 8, 1); init(B3_IFREE, 8, 13);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(reinterpret_cast<uint32_t*>(smem + F3_OFF_BARS + B3_TMEM), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + F3_OFF_BARS + B3_TMEM);
  const int off = p.Lk - p.Lq;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F3_CTRL_REGS));
    if (warp == 0) {
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
            const int nkvA = (off + lastA) / 128 + 1;                    // off + lastA <= Lk - 1
            const int lastB = min(q0 + 255, p.Lq - 1);
            const int nkvB = hasB ? (off + lastB) / 128 + 1 : 0;
            const int nkv = max(nkvA, nkvB);
            const int qb = item_idx & 1;
            const bool last_item = (next_bh < 0) && pp == 0;
            uint64_t* bq = reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_Q) + qb;
            if (item_idx >= 2) mbar_wait_s(bars + B3_QFREE + 8 * qb, ((item_idx >> 1) - 1) & 1);
            mbar_arrive_expect_tx(bq, (hasB ? 2 : 1) * F3_TILE);
            load_head_tile<DH, SWB>(smem + (qb * 2) * F3_TILE, &tmQ, bq, h, b, q0);
            if (hasB) load_head_tile<DH, SWB>(smem + (qb * 2 + 1) * F3_TILE, &tmQ, bq, h, b, q0 + 128);
            for (int j = 0; j < nkv; ++j, ++t) {
              const int st = t % F3_KV_STAGES;
              const int is = t & (F3_INFO_SLOTS - 1);
              if (t >= F3_INFO_SLOTS) mbar_wait_s(bars + B3_IFREE + 8 * is, ((t / F3_INFO_SLOTS) - 1) & 1);
              const int flags = (j == 0 ? F3_FIRST : 0) | (j < nkvA ? F3_A : 0) | (j < nkvB ? F3_B : 0) | (j == nkvA - 1 ? F3_LAST_A : 0) |
                                ((hasB && j == nkvB - 1) ? F3_LAST_B : 0) | ((last_item && j == nkv - 1) ? F3_END : 0) | (qb ? F3_QBUF : 0);
              const int nA = j < nkvA ? min(128, (off + lastA - j * 128 + 16) & ~15) : 0;     // (last visible column + 1) rounded up to 16
              const int nB = j < nkvB ? min(128, (off + lastB - j * 128 + 16) & ~15) : 0;
              *reinterpret_cast<int4*>(smem + F3_OFF_INFO + is * 16) = make_int4(q0 | (j << 16), b, h | (flags << 8), nA | (nB << 8));
              mbar_arrive_s(bars + B3_IFULL + 8 * is);               // release: publishes the slot
              uint64_t* bkv = reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_KV) + st;
              if (t >= F3_KV_STAGES) mbar_wait_s(bars + B3_KVFREE + 8 * st, ((t / F3_KV_STAGES) - 1) & 1);
              mbar_arrive_expect_tx(bkv, 2 * F3_TILE);
              load_head_tile<DH, SWB>(smem + F3_OFF_K + st * F3_TILE, &tmK, bkv, h, b, j * 128);
              load_head_tile<DH, SWB>(smem + F3_OFF_V + st * F3_TILE, &tmV, bkv, h, b, j * 128);
            }
          }
          bh = next_bh;
        }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer ==============================
      if (elect_one()) {
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);     // A = P (tensor memory, K-major), V is MN-major
        const uint64_t tileK = make_smem_desc<SWB>(0, 16);
        const uint64_t tileMN = make_smem_desc<SWB>(0, T::SLAB_BYTES);
        auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
        uint32_t n_items = 0;
        uint32_t cnt[2] = {0, 0};          // active steps handled so far per tile slot (phase counter of SFREE / P)
        uint32_t fin[2] = {0, 0};          // finished tiles per slot (phase counter of OFREE)
        auto wait_step = [&](uint32_t t) -> int4 {
          const int st = t % F3_KV_STAGES;
          const int is = t & (F3_INFO_SLOTS - 1);
          mbar_wait_s(bars + B3_IFULL + 8 * is, (t / F3_INFO_SLOTS) & 1);
          const int4 si = *reinterpret_cast<const int4*>(smem + F3_OFF_INFO + is * 16);
          mbar_arrive_s(bars + B3_IFREE + 8 * is);
          mbar_wait_s(bars + B3_KV + 8 * st, (t / F3_KV_STAGES) & 1);
          if ((si.z >> 8) & F3_FIRST) {
            mbar_wait_s(bars + B3_Q + 8 * (((si.z >> 8) & F3_QBUF) ? 1 : 0), (n_items >> 1) & 1);
            ++n_items;
          }
          return si;
        };
        auto issue_s = [&](const int4& si, uint32_t t, int x) {
          const int st = t % F3_KV_STAGES;
          const int qb = ((si.z >> 8) & F3_QBUF) ? 1 : 0;
          const uint32_t idesc_s = make_idesc_bf16(128, x ? ((si.w >> 8) & 0xff) : (si.w & 0xff), 0, 0);
          tc_fence_after();
          const uint64_t aQ = tileK + addr14(sbase + (qb * 2 + x) * F3_TILE), aK = tileK + addr14(sbase + F3_OFF_K + st * F3_TILE);
          const uint32_t d = tmem_base + F3_T_S + x * 128;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(d, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
          umma_commit(reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_S) + x);
        };
        auto issue_pv = [&](const int4& si, uint32_t t, int x) {
          const int st = t % F3_KV_STAGES;
          const int flags = si.z >> 8;
          mbar_wait_s(bars + B3_P + 8 * x, cnt[x] & 1);                      // P(t) of tile x is in TMEM, O rescaled if needed
          if ((flags & F3_FIRST) && fin[x] > 0) mbar_wait_s(bars + B3_OFREE + 8 * x, (fin[x] - 1) & 1);   // previous tile's O is out
          tc_fence_after();
          const uint64_t mV = tileMN + addr14(sbase + F3_OFF_V + st * F3_TILE);
          const uint32_t aP = tmem_base + F3_T_P + x * 64;
          const int nk = (x ? ((si.w >> 8) & 0xff) : (si.w & 0xff)) >> 4;
          const bool first = flags & F3_FIRST;
#pragma unroll 1
          for (int kk = 0; kk < nk; ++kk)
            umma_bf16_ts(tmem_base + F3_T_O + x * DH, aP + 8 * kk, mV + 128 * kk, idesc_pv, (first && kk == 0) ? 0u : 1u);
          umma_commit(reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_O) + x);
          ++cnt[x];
          if (flags & (x == 0 ? F3_LAST_A : F3_LAST_B)) ++fin[x];
        };
        uint32_t t = 0;
        int4 cur = wait_step(0);
        if ((cur.z >> 8) & F3_A) issue_s(cur, 0, 0);
        if ((cur.z >> 8) & F3_B) issue_s(cur, 0, 1);
        bool end = false;
        while (!end) {
          const int cf = cur.z >> 8;
          end = (cf & F3_END) != 0;
          int4 nxt = cur;
          if (!end) nxt = wait_step(t + 1);                               // next K/V block (and its item's Q) have landed
          const int nf = nxt.z >> 8;
          // S of the next step as soon as this step's scores have left TMEM (it runs under this step's exponentials) ...
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const int act = x == 0 ? F3_A : F3_B;
            if (cf & act) mbar_wait_s(bars + B3_SFREE + 8 * x, cnt[x] & 1);
            if (!end && (nf & act)) issue_s(nxt, t + 1, x);
          }
          // ... then P V of this step for each tile when its probabilities arrive
          if (cf & F3_A) issue_pv(cur, t, 0);
          if (cf & F3_B) issue_pv(cur, t, 1);
          umma_commit(reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_KVFREE) + (t % F3_KV_STAGES));   // every MMA reading stage t is issued
          if (end || (nf & F3_FIRST)) umma_commit(reinterpret_cast<uint64_t*>(smem + F3_OFF_BARS + B3_QFREE) + ((cf & F3_QBUF) ? 1 : 0));
          cur = nxt;
          ++t;
        }
      }
    }
  } else if (warp < 8) {
    // ============================== epilogue warpgroup (warps 4-7): finishes tiles of both slots ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F3_EPI_REGS));
    const int wrow = (warp & 3) * 32;
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const int et = tid - 128;                     // 0..127
    uint32_t g = 0, n[2] = {0, 0}, fin[2] = {0, 0};
    bool end = false;
    bool store_pending = false;
    while (!end) {
      const int is = g & (F3_INFO_SLOTS - 1);
      mbar_wait_s(bars + B3_IFULL + 8 * is, (g / F3_INFO_SLOTS) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + F3_OFF_INFO + is * 16);
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bars + B3_IFREE + 8 * is);
      const int flags = si.z >> 8;
      end = (flags & F3_END) != 0;
      ++g;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (!(flags & (x == 0 ? F3_A : F3_B))) continue;
        ++n[x];
        if (!(flags & (x == 0 ? F3_LAST_A : F3_LAST_B))) continue;
        const int q0 = (si.x & 0xffff) + x * 128;
        const int b = si.y, h = si.z & 0xff;
        const uint32_t buf = fin[x] & 1;
        mbar_wait_s(bars + B3_STATS + 8 * x, fin[x] & 1);           // l, m of the tile's rows are in shared memory
        mbar_wait_s(bars + B3_O + 8 * x, (n[x] - 1) & 1);           // the tile's last P V has completed
        tc_fence_after();
        uint32_t w0[32], w1[32];
        tmem_ld_x32(t_lane + F3_T_O + x * DH, w0);
        tmem_ld_x32(t_lane + F3_T_O + x * DH + 32, w1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bars + B3_OFREE + 8 * x);      // the next tile of this slot may overwrite O
        const float2 lm = *reinterpret_cast<const float2*>(smem + F3_OFF_STATS + ((x * 2 + buf) * 128 + row) * 8);
        const float inv = 1.0f / lm.x;
        if (q0 + row < p.Lq) p.lse[((long long)b * p.H + h) * p.Lq + q0 + row] = lm.y * p.scale + logf(lm.x);
        // the staging tile is free once the previous tile's TMA store has read it
        if (store_pending) {
          if (et == 0) bulk_wait_read0();
          named_bar_sync(3, 128);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          *reinterpret_cast<uint4*>(smem + F3_OFF_STG + swz_off<128>(row, ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(w0[ch * 8 + 0]) * inv, __uint_as_float(w0[ch * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 2]) * inv, __uint_as_float(w0[ch * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 4]) * inv, __uint_as_float(w0[ch * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 6]) * inv, __uint_as_float(w0[ch * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(smem + F3_OFF_STG + swz_off<128>(row, 4 + ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(w1[ch * 8 + 0]) * inv, __uint_as_float(w1[ch * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 2]) * inv, __uint_as_float(w1[ch * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 4]) * inv, __uint_as_float(w1[ch * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 6]) * inv, __uint_as_float(w1[ch * 8 + 7]) * inv));
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (et == 0) {
          tma_store_3d(&tmO, sbase + F3_OFF_STG, h * DH, b, q0);    // rows past Lq are clipped by the tensor map
          bulk_commit();
        }
        store_pending = true;
        ++fin[x];
      }
    }
    if (et == 0) bulk_wait_all();
  } else {
    // ============================== softmax warpgroups (warps 8-11: tile A, 12-15: tile B) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F3_SM_REGS));
    const int x = (warp >> 2) - 2;                 // tile slot
    const int wrow = (warp & 3) * 32;              // first tile row of this warp == first TMEM lane
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const uint32_t t_s = t_lane + F3_T_S + x * 128, t_o = t_lane + F3_T_O + x * DH, t_p = t_lane + F3_T_P + x * 64;
    const int act_flag = x == 0 ? F3_A : F3_B, last_flag = x == 0 ? F3_LAST_A : F3_LAST_B;
    const uint32_t bar_s = bars + B3_S + 8 * x, bar_sfree = bars + B3_SFREE + 8 * x, bar_p = bars + B3_P + 8 * x;
    const uint32_t bar_o = bars + B3_O + 8 * x, bar_stats = bars + B3_STATS + 8 * x;
    uint32_t g = 0;          // global step counter (every step of the CTA)
    uint32_t n = 0;          // steps of this tile slot so far (phase counter of S / O)
    uint32_t fin = 0;        // finished tiles of this slot
    bool end = false;
    float m_ref = -INFINITY, l_run = 0.0f;

    while (!end) {
      const int is = g & (F3_INFO_SLOTS - 1);
      mbar_wait_s(bars + B3_IFULL + 8 * is, (g / F3_INFO_SLOTS) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + F3_OFF_INFO + is * 16);
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bars + B3_IFREE + 8 * is);
      const int flags = si.z >> 8;
      end = (flags & F3_END) != 0;
      ++g;
      if (!(flags & act_flag)) continue;
      const bool first = flags & F3_FIRST, last = flags & last_flag;
      const int j128 = (si.x >> 16) * 128;
      const int q0 = (si.x & 0xffff) + x * 128;
      const int ncols = x ? ((si.w >> 8) & 0xff) : (si.w & 0xff);
      const bool warp_valid = (q0 + wrow) < p.Lq;   // a warp whose 32 rows lie past the end of the query tail only keeps the barriers moving
      // column i of this block is visible to this row iff i <= lim (causal mask aligned to the sequence tail, OT/model.py:64,109)
      const int lim_lo = (off + q0 + wrow) - j128;                  // lane 0; lane 31 has lim_lo + 31
      const int lim = lim_lo + lane;
      // Chunk c (32 columns) is, for the whole warp, hidden (vis == 0), cut by the diagonal (1) or fully visible (2).
      int vis[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) vis[c] = (!warp_valid || lim_lo + 31 < c * 32) ? 0 : (lim_lo >= c * 32 + 31) ? 2 : 1;
      mbar_wait_s(bar_s, n & 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (vis[c]) tmem_ld_x32(t_s + c * 32, s[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bar_sfree);                      // the S columns may be overwritten by the next block's product

      bool waited_o = (n == 0);                                     // nothing to wait for before the very first P of this slot
      if (warp_valid) {
        // ---- row maximum (four independent partial maxima) ----
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (vis[c] == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[c][i]));
          } else if (vis[c] == 1) {
            // cut by the diagonal: the scores this row may not see are replaced by -inf HERE, once - they then drop out of the maximum
            // (a maximum over hidden columns would let the rounding of a row depend on LATER keys: the bit-exact causality the tests
            // pin, T6) and give exactly 0 in the exponentials below, which therefore take the mask-free packed path for this chunk too
            const int lim_c = lim - c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[c][i] = (i <= lim_c) ? s[c][i] : 0xff800000u;
              mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[c][i]));
            }
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // ---- lazy rescale: keep the old reference maximum unless the new maximum exceeds it by more than 2^TAU ----
        // (key 0 is visible to every query, so mx is finite in the first block of a tile)
        bool need = false;
        float m_new = m_ref;
        if (first) { m_new = mx; }
        else if ((mx - m_ref) * p.scale_log2 > F3_TAU) { m_new = mx; need = true; }
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait_s(bar_o, (n - 1) & 1);                         // P V of the previous block has landed in O
          waited_o = true;
          tc_fence_after();
          const float alpha = need ? ex2_approx((m_ref - m_new) * p.scale_log2) : 1.0f;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint32_t w[8];
            tmem_ld_x8(t_o + c * 8, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
            tmem_st_x8(t_o + c * 8, w);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
        m_ref = m_new;
      }
      // ---- p = 2^((s - m_ref) * scale * log2 e), row sums, packed bf16 pairs written back to TMEM chunk by chunk ----
      const float mb = m_ref * p.scale_log2;
      f32x2 rs4[2] = {pk2(0.0f), pk2(0.0f)};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 < ncols) {
          uint32_t pk[16];
          if (vis[c] != 0) {                                          // hidden scores of a cut chunk are -inf by now: 2^(-inf) = 0
            f3_softmax_chunk<false>(s[c], pk, p.scale_log2, mb, 0, rs4);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          }
          if (!waited_o) {                                           // the previous block's P must have been consumed by its P V
            mbar_wait_s(bar_o, (n - 1) & 1);
            waited_o = true;
            tc_fence_after();
          }
          if (warp_valid) tmem_st_x16(t_p + c * 16, pk);
        }
      }
      if (!waited_o) mbar_wait_s(bar_o, (n - 1) & 1);               // (keeps the phase sequence of a warp that stored nothing)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bar_p);
      {
        float r0, r1, r2, r3;
        upk2(rs4[0], r0, r1);
        upk2(rs4[1], r2, r3);
        l_run += (r0 + r1) + (r2 + r3);
      }

      if (last) {
        // hand the row statistics to the epilogue warpgroup and go on with the next tile
        *reinterpret_cast<float2*>(smem + F3_OFF_STATS + ((x * 2 + (fin & 1)) * 128 + row) * 8) = make_float2(l_run, m_ref);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bar_stats);
        ++fin;
        m_ref = -INFINITY;
        l_run = 0.0f;
      }
      ++n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

int attn_fwd_v3_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  if ((rc = make_head_tmap(&to, p->o, cols, p->B, p->Lq, p->ldo, 128))) return rc;
  AttnFwdV3KParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.n_pairs = (kp.n_qt + 1) / 2;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.lse = p->lse;
  kp.sched = sched_slot(st);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_fwd_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F3_SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms();
  const int n_bh = p->B * p->H;
  const int grid = n_bh < sms ? n_bh : sms;
  ot_attn_fwd_v3_kernel<<<grid, F3_THREADS, F3_SMEM_BYTES, st>>>(tq, tk, tv, to, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
