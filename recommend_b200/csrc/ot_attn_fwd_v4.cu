// ot_attn_fwd_v4.cu — causal attention forward for head_dim 64: TWO INDEPENDENT TILE STREAMS per CTA (OT/model.py:101-114 for the
// retained query tail; the softmax arithmetic is ot_attn_fwd_v3.cu's, shared through ot_attn_fwd_common.cuh).
// Why (profiles/README.md, round 2 second session).  v3 runs the query tiles 2p and 2p+1 of a (sample, head) in lockstep over a
// shared K/V stream.  Under the causal mask the later tile always needs one key block more, so slot A idles in 2 of the 8 steps
// of a four-tile (sample, head) and in 4 of 7 when the tile count is odd; one MMA issuer serves both slots in order, so the
// slot that is ahead waits for the other's probabilities before its own next S product is issued (ncu: 11 % of the softmax
// warps' time on that barrier); and a one-tile (sample, head) uses half of the CTA.  Here each slot is a pipeline of its own:
//   loader_x -> (Q tile, K ring, V ring, step ring)_x -> MMA issuer_x -> softmax warpgroup_x,     x = A, B
// with its own barriers; the tiles of a (sample, head) are dealt to the two slots so that their key-block counts balance
// (snake order by length; one-tile shapes alternate whole (sample, head)s), and a shared epilogue warpgroup finishes the tiles of
// both.  The K and V rings are released separately - K after the S product, V after P V - so two stages each are a step and
// a half of prefetch.  Both slots of a CTA walk the same (sample, head) sequence, so its K/V rows come from DRAM once.
// Warps: 0 loader A, 1 MMA issuer A, 2 loader B, 3 MMA issuer B, 4-7 epilogue, 8-11 softmax A, 12-15 softmax B.
// TMEM: S_A 0, S_B 128, O_A 256, O_B 320, P_A 384, P_B 448.
#include "ot_attn_fwd_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdV4KParams {
  int B, H, Lq, Lk, n_qt;
  float scale, scale_log2;
  float* lse;  // [B, H, Lq]
  int* sched;
};

static constexpr int F4_THREADS = 512;
static constexpr int F4_CTRL_REGS = 40, F4_EPI_REGS = 72, F4_SM_REGS = 200;   // 128 x (128-40) + 128 x (128-72) released = 256 x (200-128) taken
static constexpr int F4_DH = 64;
static constexpr int F4_TILE = 128 * F4_DH * 2;          // 16 KB
static constexpr int F4_OFF_Q = 0;                                   // [slot][2]
static constexpr int F4_OFF_K = F4_OFF_Q + 4 * F4_TILE;              // [slot][2]
static constexpr int F4_OFF_V = F4_OFF_K + 4 * F4_TILE;              // [slot][2]
static constexpr int F4_OFF_STG = F4_OFF_V + 4 * F4_TILE;            // output staging tile
static constexpr int F4_OFF_STATS = F4_OFF_STG + F4_TILE;            // [slot][2][128] float2 (l, m)
static constexpr int F4_OFF_REC = F4_OFF_STATS + 2 * 2 * 128 * 8;    // [slot][2] int4: tile-end records for the epilogue
static constexpr int F4_OFF_INFO = F4_OFF_REC + 4 * 16;              // [slot][4] int4
static constexpr int F4_OFF_BH = F4_OFF_INFO + 8 * 16;               // [8] int: the CTA's (sample, head) sequence + [2] int epilogue selector
static constexpr int F4_OFF_BARS = F4_OFF_BH + 64;
static constexpr int F4_SMEM_BYTES = F4_OFF_BARS + 640;
static_assert(F4_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t F4_T_S = 0, F4_T_O = 256, F4_T_P = 384;     // + slot * 128 / 64 / 64
// mbarrier byte offsets; per-slot blocks of F4_SLOT_BARS bytes follow the shared ones
enum : uint32_t {
  C4_BHFULL = 0,     // [8] (sample, head) published                      (loader A -> loader B)
  C4_BHFREE = 64,    // [8] ... and read                                  (loader B -> loader A)
  C4_TMEM = 128,
  C4_SLOT0 = 160,
  F4_SLOT_BARS = 232,
  S4_Q = 0,          // [2] Q tile landed                                 (loader -> MMA)
  S4_QFREE = 16,     // [2] every S product of the tile issued + complete (MMA commit -> loader)
  S4_K = 32,         // [2] K block landed
  S4_KFREE = 48,     // [2] its S product complete                        (MMA commit -> loader)
  S4_V = 64,         // [2] V block landed
  S4_VFREE = 80,     // [2] its P V product complete                      (MMA commit -> loader)
  S4_S = 96,         //     S complete                                    (MMA commit -> softmax)
  S4_SFREE = 104,    //     S pulled into registers                       (4 arrivals -> MMA)
  S4_P = 112,        //     P written to TMEM                             (4 arrivals -> MMA)
  S4_O = 120,        //     P V complete                                  (MMA commit -> softmax, epilogue)
  S4_OFREE = 128,    //     O of a finished tile read out                 (4 epilogue arrivals -> MMA)
  S4_STATS = 136,    //     tile-end record + row statistics written      (4 softmax arrivals -> epilogue)
  S4_IFULL = 144,    // [4] step info published                           (loader -> MMA, softmax)
  S4_IFREE = 176     // [4] ... and read: MMA issuer + 4 softmax warps = 5 arrivals -> loader
};
enum { G4_FIRST = 1, G4_LAST = 2, G4_END = 4 };
// step info, 16 bytes: x = q0 | j << 16, y = b, z = h | flags << 8, w = ncols | qbuf << 8
// tile-end record, 16 bytes: x = q0 (or -1: this slot is finished), y = b, z = h, w = steps of the slot so far

// which slot takes the k-th longest tile of a (sample, head) with n_qt tiles (seq = the pair's index in the CTA's sequence)
__device__ __forceinline__ int f4_slot_of(int k, int n_qt, uint32_t seq) {
  return n_qt == 1 ? static_cast<int>(seq & 1u) : (((k + 1) >> 1) & 1);      // snake: A B B A A B B A ...
}

__global__ void __launch_bounds__(F4_THREADS, 1)
ot_attn_fwd_v4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ AttnFwdV4KParams p) {
  constexpr int DH = F4_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + F4_OFF_BARS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  volatile int* bh_ring = reinterpret_cast<volatile int*>(smem + F4_OFF_BH);
  volatile int* epi_sel = bh_ring + 8;

  if (tid == 0) {
    if ((sbase & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    auto init = [&](uint32_t off, int n, uint32_t count) {
      for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + off + 8 * i), "r"(count));
    };
    init(C4_BHFULL, 8, 1); init(C4_BHFREE, 8, 1);
    for (int x = 0; x < 2; ++x) {
      const uint32_t sb = C4_SLOT0 + x * F4_SLOT_BARS;
      init(sb + S4_Q, 2, 1); init(sb + S4_QFREE, 2, 1); init(sb + S4_K, 2, 1); init(sb + S4_KFREE, 2, 1); init(sb + S4_V, 2, 1); init(sb + S4_VFREE, 2, 1);
      init(sb + S4_S, 1, 1); init(sb + S4_SFREE, 1, 4); init(sb + S4_P, 1, 4); init(sb + S4_O, 1, 1); init(sb + S4_OFREE, 1, 4); init(sb + S4_STATS, 1, 4);
      init(sb + S4_IFULL, 4, 1); init(sb + S4_IFREE, 4, 5);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(reinterpret_cast<uint32_t*>(smem + F4_OFF_BARS + C4_TMEM), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + F4_OFF_BARS + C4_TMEM);
  const int off = p.Lk - p.Lq;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F4_CTRL_REGS));
    const int x = warp >> 1;                                   // slot of this control warp
    const uint32_t sb = bars + C4_SLOT0 + x * F4_SLOT_BARS;
    auto bar_ptr = [&](uint32_t off_in_slot) -> uint64_t* {
      return reinterpret_cast<uint64_t*>(smem + F4_OFF_BARS + C4_SLOT0 + x * F4_SLOT_BARS + off_in_slot);
    };
    if ((warp & 1) == 0) {
      // ============================== loader of slot x ==============================
      if (elect_one()) {
        const int n_bh = p.B * p.H;
        uint32_t t = 0, item = 0;
        int bh_next = blockIdx.x;                              // loader A only: the next (sample, head) to publish
        auto publish = [&](int q0, int j, int b, int h, int flags, int ncols, int qb) {
          const int is = t & 3;
          if (t >= 4) mbar_wait_s(sb + S4_IFREE + 8 * is, ((t >> 2) - 1) & 1);
          *reinterpret_cast<int4*>(smem + F4_OFF_INFO + (x * 4 + is) * 16) = make_int4(q0 | (j << 16), b, h | (flags << 8), ncols | (qb << 8));
          mbar_arrive_s(sb + S4_IFULL + 8 * is);
        };
        for (uint32_t seq = 0;; ++seq) {
          int bh;
          const int rs = seq & 7;
          if (x == 0) {
            // loader A draws the CTA's (sample, head) sequence from the device-wide counter and hands it to loader B
            bh = bh_next;
            if (bh >= 0) {
              bh_next = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : bh + (int)gridDim.x;
              if (bh_next >= n_bh) bh_next = -1;
            }
            if (seq >= 8) mbar_wait_s(bars + C4_BHFREE + 8 * rs, ((seq >> 3) - 1) & 1);
            bh_ring[rs] = bh;
            mbar_arrive_s(bars + C4_BHFULL + 8 * rs);
          } else {
            mbar_wait_s(bars + C4_BHFULL + 8 * rs, (seq >> 3) & 1);
            bh = bh_ring[rs];
            mbar_arrive_s(bars + C4_BHFREE + 8 * rs);
          }
          if (bh < 0) break;
          const int h = bh % p.H, b = bh / p.H;
          for (int k = 0; k < p.n_qt; ++k) {                   // k-th longest tile = tile n_qt - 1 - k
            if (f4_slot_of(k, p.n_qt, seq) != x) continue;
            const int q0 = (p.n_qt - 1 - k) * 128;
            const int lastq = min(q0 + 127, p.Lq - 1);
            const int nkv = (off + lastq) / 128 + 1;           // off + lastq <= Lk - 1
            const int qb = item & 1;
            if (item >= 2) mbar_wait_s(sb + S4_QFREE + 8 * qb, ((item >> 1) - 1) & 1);
            mbar_arrive_expect_tx(bar_ptr(S4_Q + 8 * qb), F4_TILE);
            load_head_tile<DH, SWB>(smem + F4_OFF_Q + (x * 2 + qb) * F4_TILE, &tmQ, bar_ptr(S4_Q + 8 * qb), h, b, q0);
            for (int j = 0; j < nkv; ++j, ++t) {
              const int flags = (j == 0 ? G4_FIRST : 0) | (j == nkv - 1 ? G4_LAST : 0);
              const int ncols = min(128, (off + lastq - j * 128 + 16) & ~15);      // (last visible column + 1) rounded up to 16
              publish(q0, j, b, h, flags, ncols, qb);
              const int st = t & 1;
              if (t >= 2) mbar_wait_s(sb + S4_KFREE + 8 * st, ((t >> 1) - 1) & 1);
              mbar_arrive_expect_tx(bar_ptr(S4_K + 8 * st), F4_TILE);
              load_head_tile<DH, SWB>(smem + F4_OFF_K + (x * 2 + st) * F4_TILE, &tmK, bar_ptr(S4_K + 8 * st), h, b, j * 128);
              if (t >= 2) mbar_wait_s(sb + S4_VFREE + 8 * st, ((t >> 1) - 1) & 1);
              mbar_arrive_expect_tx(bar_ptr(S4_V + 8 * st), F4_TILE);
              load_head_tile<DH, SWB>(smem + F4_OFF_V + (x * 2 + st) * F4_TILE, &tmV, bar_ptr(S4_V + 8 * st), h, b, j * 128);
            }
            ++item;
          }
        }
        publish(0, 0, 0, 0, G4_END, 0, 0);                     // end marker: a step without work
      }
    } else {
      // ============================== MMA issuer of slot x ==============================
      if (elect_one()) {
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);     // A = P (tensor memory, K-major), V is MN-major
        const uint64_t tileK = make_smem_desc<SWB>(0, 16);
        const uint64_t tileMN = make_smem_desc<SWB>(0, T::SLAB_BYTES);
        auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
        const uint32_t t_s = tmem_base + F4_T_S + x * 128, t_o = tmem_base + F4_T_O + x * DH, t_p = tmem_base + F4_T_P + x * 64;
        uint32_t items = 0, fin = 0;
        auto wait_info = [&](uint32_t t) -> int4 {
          const int is = t & 3;
          mbar_wait_s(sb + S4_IFULL + 8 * is, (t >> 2) & 1);
          const int4 si = *reinterpret_cast<const int4*>(smem + F4_OFF_INFO + (x * 4 + is) * 16);
          mbar_arrive_s(sb + S4_IFREE + 8 * is);
          return si;
        };
        auto issue_s = [&](const int4& si, uint32_t t) {      // S = Q K^T of step t over the visible key columns
          const int flags = (si.z >> 8) & 0xff, st = t & 1, qb = (si.w >> 8) & 1;
          mbar_wait_s(sb + S4_K + 8 * st, (t >> 1) & 1);
          if (flags & G4_FIRST) { mbar_wait_s(sb + S4_Q + 8 * qb, (items >> 1) & 1); ++items; }
          const uint32_t idesc_s = make_idesc_bf16(128, si.w & 0xff, 0, 0);
          tc_fence_after();
          const uint64_t aQ = tileK + addr14(sbase + F4_OFF_Q + (x * 2 + qb) * F4_TILE), aK = tileK + addr14(sbase + F4_OFF_K + (x * 2 + st) * F4_TILE);
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(t_s, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
          umma_commit(bar_ptr(S4_S));
          umma_commit(bar_ptr(S4_KFREE + 8 * st));
          if (flags & G4_LAST) umma_commit(bar_ptr(S4_QFREE + 8 * qb));     // the tile's last S product: its Q buffer is free when this completes
        };
        auto info_ready = [&](uint32_t t) { return mbar_test_s(sb + S4_IFULL + 8 * (t & 3), (t >> 2) & 1); };
        uint32_t t = 0;
        int4 cur = wait_info(0);
        if (!((cur.z >> 8) & G4_END)) {
          issue_s(cur, 0);
          while (true) {
            const int cf = (cur.z >> 8) & 0xff;
            // S of the next step as soon as this step's scores have left TMEM (it runs under this step's exponentials) - but only if
            // the loader has already published that step.  The issuer must never BLOCK on the next step's info before this step's
            // P V is out: the loaders are coupled through the (sample, head) ring, so loader A can be held by loader B, B by its own
            // slot, that slot by the epilogue, and the epilogue by the P V of slot A's finished tile - a cycle when that P V waits for
            // loader A (seen with many one-step tiles per CTA: B 4096, Lq <= 24).
            int4 nxt = make_int4(0, 0, 0, 0);
            bool have_nxt = false, s_issued = false;
            // So: poll for whichever comes first, the next step's info (normal case: its S goes out now) or this step's probabilities.
            while (true) {
              if (info_ready(t + 1)) {
                nxt = wait_info(t + 1);
                have_nxt = true;
                if (!((nxt.z >> 8) & G4_END)) {
                  mbar_wait_s(sb + S4_SFREE, t & 1);
                  issue_s(nxt, t + 1);
                  s_issued = true;
                }
                break;
              }
              if (mbar_test_s(sb + S4_P, t & 1)) break;
            }
            // ... then P V of this step when its probabilities arrive
            mbar_wait_s(sb + S4_P, t & 1);
            if ((cf & G4_FIRST) && fin > 0) mbar_wait_s(sb + S4_OFREE, (fin - 1) & 1);     // the previous tile's O is out
            mbar_wait_s(sb + S4_V + 8 * (t & 1), (t >> 1) & 1);
            tc_fence_after();
            {
              const uint64_t mV = tileMN + addr14(sbase + F4_OFF_V + (x * 2 + (t & 1)) * F4_TILE);
              const int nk = (cur.w & 0xff) >> 4;
              const bool first = cf & G4_FIRST;
#pragma unroll 1
              for (int kk = 0; kk < nk; ++kk) umma_bf16_ts(t_o, t_p + 8 * kk, mV + 128 * kk, idesc_pv, (first && kk == 0) ? 0u : 1u);
            }
            umma_commit(bar_ptr(S4_O));
            umma_commit(bar_ptr(S4_VFREE + 8 * (t & 1)));
            if (cf & G4_LAST) ++fin;
            if (!have_nxt) nxt = wait_info(t + 1);
            if ((nxt.z >> 8) & G4_END) break;
            if (!s_issued) {
              mbar_wait_s(sb + S4_SFREE, t & 1);
              issue_s(nxt, t + 1);
            }
            cur = nxt;
            ++t;
          }
        }
      }
    }
  } else if (warp < 8) {
    // ============================== epilogue warpgroup (warps 4-7): finishes the tiles of both slots ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F4_EPI_REGS));
    const int wrow = (warp & 3) * 32;
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const int et = tid - 128;                     // 0..127
    uint32_t fin[2] = {0, 0};
    bool done[2] = {false, false};
    bool store_pending = false;
    uint32_t it = 0;
    while (!(done[0] && done[1])) {
      // one thread finds the slot whose next tile-end record is ready; everybody follows it
      if (et == 0) {
        int sel = -1;
        uint32_t spins = 0;
        while (sel < 0) {
          if (!done[0] && mbar_test_s(bars + C4_SLOT0 + S4_STATS, fin[0] & 1)) sel = 0;
          else if (!done[1] && mbar_test_s(bars + C4_SLOT0 + F4_SLOT_BARS + S4_STATS, fin[1] & 1)) sel = 1;
          else {
            __nanosleep(100);
#if OT_HANG_DEBUG
            if (++spins > (1u << 18)) { ot_hang_note(2u, fin[0], fin[1]); sel = done[0] ? 1 : 0; }
#elif OT_HANG_GUARD
            if (++spins > (1u << 25)) __trap();
#endif
          }
        }
        epi_sel[it & 1] = sel;
      }
      named_bar_sync(3, 128);
      const int x = epi_sel[it & 1];
      ++it;
      const uint32_t sb = bars + C4_SLOT0 + x * F4_SLOT_BARS;
      const uint32_t buf = fin[x] & 1;
      const int4 rec = *reinterpret_cast<const int4*>(smem + F4_OFF_REC + (x * 2 + buf) * 16);
      ++fin[x];
      if (rec.x < 0) { done[x] = true; continue; }
      const int q0 = rec.x, b = rec.y, h = rec.z;
      mbar_wait_s(sb + S4_O, (static_cast<uint32_t>(rec.w) - 1) & 1);     // the tile's last P V has completed
      tc_fence_after();
      uint32_t w0[32], w1[32];
      tmem_ld_x32(t_lane + F4_T_O + x * DH, w0);
      tmem_ld_x32(t_lane + F4_T_O + x * DH + 32, w1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_s(sb + S4_OFREE);                        // the next tile of this slot may overwrite O
      const float2 lm = *reinterpret_cast<const float2*>(smem + F4_OFF_STATS + ((x * 2 + buf) * 128 + row) * 8);
      const float inv = 1.0f / lm.x;
      if (q0 + row < p.Lq) p.lse[((long long)b * p.H + h) * p.Lq + q0 + row] = lm.y * p.scale + logf(lm.x);
      if (store_pending) {                                                // the previous tile's TMA store has read the staging tile
        if (et == 0) bulk_wait_read0();
        named_bar_sync(3, 128);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        *reinterpret_cast<uint4*>(smem + F4_OFF_STG + swz_off<128>(row, ch)) =
            make_uint4(pack_bf16x2(__uint_as_float(w0[ch * 8 + 0]) * inv, __uint_as_float(w0[ch * 8 + 1]) * inv),
                       pack_bf16x2(__uint_as_float(w0[ch * 8 + 2]) * inv, __uint_as_float(w0[ch * 8 + 3]) * inv),
                       pack_bf16x2(__uint_as_float(w0[ch * 8 + 4]) * inv, __uint_as_float(w0[ch * 8 + 5]) * inv),
                       pack_bf16x2(__uint_as_float(w0[ch * 8 + 6]) * inv, __uint_as_float(w0[ch * 8 + 7]) * inv));
        *reinterpret_cast<uint4*>(smem + F4_OFF_STG + swz_off<128>(row, 4 + ch)) =
            make_uint4(pack_bf16x2(__uint_as_float(w1[ch * 8 + 0]) * inv, __uint_as_float(w1[ch * 8 + 1]) * inv),
                       pack_bf16x2(__uint_as_float(w1[ch * 8 + 2]) * inv, __uint_as_float(w1[ch * 8 + 3]) * inv),
                       pack_bf16x2(__uint_as_float(w1[ch * 8 + 4]) * inv, __uint_as_float(w1[ch * 8 + 5]) * inv),
                       pack_bf16x2(__uint_as_float(w1[ch * 8 + 6]) * inv, __uint_as_float(w1[ch * 8 + 7]) * inv));
      }
      fence_proxy_async_smem();
      named_bar_sync(3, 128);
      if (et == 0) {
        tma_store_3d(&tmO, sbase + F4_OFF_STG, h * DH, b, q0);            // rows past Lq are clipped by the tensor map
        bulk_commit();
      }
      store_pending = true;
    }
    if (et == 0) bulk_wait_all();
  } else {
    // ============================== softmax warpgroups (warps 8-11: slot A, 12-15: slot B) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F4_SM_REGS));
    const int x = (warp >> 2) - 2;
    const int wrow = (warp & 3) * 32;              // first tile row of this warp == first TMEM lane
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const uint32_t t_s = t_lane + F4_T_S + x * 128, t_o = t_lane + F4_T_O + x * DH, t_p = t_lane + F4_T_P + x * 64;
    const uint32_t sb = bars + C4_SLOT0 + x * F4_SLOT_BARS;
    const uint32_t bar_s = sb + S4_S, bar_sfree = sb + S4_SFREE, bar_p = sb + S4_P, bar_o = sb + S4_O, bar_stats = sb + S4_STATS;
    uint32_t n = 0;          // steps of this slot so far (phase counter of every per-step barrier)
    uint32_t fin = 0;        // finished tiles of this slot
    float m_ref = -INFINITY, l_run = 0.0f;

    while (true) {
      const int is = n & 3;
      mbar_wait_s(sb + S4_IFULL + 8 * is, (n >> 2) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + F4_OFF_INFO + (x * 4 + is) * 16);
      __syncwarp();
      if (lane == 0) mbar_arrive_s(sb + S4_IFREE + 8 * is);
      const int flags = (si.z >> 8) & 0xff;
      if (flags & G4_END) {
        if (fin > 0) mbar_wait_s(sb + S4_OFREE, (fin - 1) & 1);       // (see the tile-end hand-over below)
        if (warp == 8 + 4 * x && lane == 0) *reinterpret_cast<int4*>(smem + F4_OFF_REC + (x * 2 + (fin & 1)) * 16) = make_int4(-1, 0, 0, 0);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bar_stats);                    // "this slot is finished" record for the epilogue
        break;
      }
      const bool first = flags & G4_FIRST, last = flags & G4_LAST;
      const int j128 = (si.x >> 16) * 128;
      const int q0 = si.x & 0xffff;
      const int ncols = si.w & 0xff;
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
      ++n;
      if (last) {
        // hand the row statistics and the tile's coordinates to the epilogue warpgroup and go on with the next tile.  The
        // epilogue must have taken the previous record first (it arrives on OFREE when it has read that tile's O): a parity
        // wait cannot tell a barrier that is two phases ahead from one that has not moved, and two records in a row happen
        // (a one-block tile, or the end marker right behind the last tile).
        if (fin > 0) mbar_wait_s(sb + S4_OFREE, (fin - 1) & 1);
        *reinterpret_cast<float2*>(smem + F4_OFF_STATS + ((x * 2 + (fin & 1)) * 128 + row) * 8) = make_float2(l_run, m_ref);
        if (warp == 8 + 4 * x && lane == 0)
          *reinterpret_cast<int4*>(smem + F4_OFF_REC + (x * 2 + (fin & 1)) * 16) = make_int4(q0, si.y, si.z & 0xff, static_cast<int>(n));
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bar_stats);
        ++fin;
        m_ref = -INFINITY;
        l_run = 0.0f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);

#if OT_HANG_DEBUG
}  // namespace ot
extern "C" int ot_debug_hang_read(unsigned int* out8 /* [64] */) {     // debugging builds only: the record of the first wait that gave up, then cleared
  unsigned int z[64] = {0};
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out8, ot::ot_hang_rec, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(ot::ot_hang_rec, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
namespace ot {
#endif

int attn_fwd_v4_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  if ((rc = make_head_tmap(&to, p->o, cols, p->B, p->Lq, p->ldo, 128))) return rc;
  AttnFwdV4KParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.lse = p->lse;
  kp.sched = sched_slot(st);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_fwd_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms();
  const int n_bh = p->B * p->H;
  const int grid = n_bh < sms ? n_bh : sms;
  ot_attn_fwd_v4_kernel<<<grid, F4_THREADS, F4_SMEM_BYTES, st>>>(tq, tk, tv, to, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
