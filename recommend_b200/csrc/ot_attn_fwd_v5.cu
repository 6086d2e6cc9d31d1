// ot_attn_fwd_v5.cu — causal attention forward for head_dim 64: the v3 pipeline (P in tensor memory, TS-form P V, whole score row in
// registers, epilogue warpgroup; OT/model.py:101-114 for the retained query tail) with a BALANCED step schedule.  v3 pairs the
// query tiles 2p and 2p+1 of a (sample, head) and walks their key blocks in lockstep; under the causal mask the later tile needs one
// block more, so slot A idles in 2 of the 8 steps of a four-tile (sample, head) and in 4 of 7 when the tile count is odd.  v4 (two
// fully independent streams) balanced the slots but doubled the K/V loads and was slower.  Here the two slots still run in common
// steps under one MMA issuer, but each slot carries its own (tile, key block) position: the tiles of a (sample, head) are handed out
// longest first to whichever slot frees up, a slot starts its next tile in the step after it finished one, and the loader issues ONE
// K/V load when both slots want the same block in a step (the common case while two tiles run side by side) and two otherwise:
// 7 steps and 10 block loads instead of 8 and 8 for four tiles, 5 steps instead of 7 for three.
// Warps / TMEM / barriers as in ot_attn_fwd_v3.cu; Q buffers are per slot (two each), the step info is 32 bytes.
#include "ot_attn_fwd_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnFwdV5KParams {
  int B, H, Lq, Lk, n_qt, n_pairs;      // n_pairs = ceil(n_qt / 2) tile pairs per (sample, head)
  float scale, scale_log2;
  float* lse;  // [B, H, Lq]
  int* sched;
};

static constexpr int F5_THREADS = 512;
static constexpr int F5_CTRL_REGS = 40, F5_EPI_REGS = 72, F5_SM_REGS = 200;   // 128 x (128-40) + 128 x (128-72) released = 256 x (200-128) taken
static constexpr int F5_DH = 64;
static constexpr int F5_TILE = 128 * F5_DH * 2;          // 16 KB
static constexpr int F5_KV_STAGES = 4;
static constexpr int F5_INFO_SLOTS = 8;
static constexpr int F5_OFF_K = 4 * F5_TILE;                              // Q: [item buffer][tile] = 4 tiles
static constexpr int F5_OFF_V = F5_OFF_K + F5_KV_STAGES * F5_TILE;
static constexpr int F5_OFF_STG = F5_OFF_V + F5_KV_STAGES * F5_TILE;      // output staging tile
static constexpr int F5_OFF_STATS = F5_OFF_STG + F5_TILE;                 // [slot][buffer][128] float2 (l, m)
static constexpr int F5_OFF_INFO = F5_OFF_STATS + 2 * 2 * 128 * 8;
static constexpr int F5_OFF_BARS = F5_OFF_INFO + F5_INFO_SLOTS * 32;
static constexpr int F5_SMEM_BYTES = F5_OFF_BARS + 640;
static_assert(F5_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t F5_T_S = 0, F5_T_O = 256, F5_T_P = 384;     // + tile * 128 / + tile * 64 / + tile * 64
// mbarrier byte offsets inside the barrier block
enum : uint32_t {
  B5_Q = 0,          // [4] Q tile of slot x, buffer qb landed (index 2 x + qb)     (loader -> MMA)
  B5_QFREE = 32,     // [4] every S product of the tile is complete                 (MMA commit -> loader)
  B5_KV = 64,        // [4] K/V block landed                                        (loader -> MMA)
  B5_KVFREE = 96,    // [4] every product that reads the stage is complete          (MMA commit -> loader)
  B5_S = 128,        // [2] S of the slot complete                                  (MMA commit -> softmax)
  B5_SFREE = 144,    // [2] S pulled into registers                                 (4 arrivals -> MMA)
  B5_P = 160,        // [2] P written to TMEM, O rescaled                           (4 arrivals -> MMA)
  B5_O = 176,        // [2] P V of the slot complete                                (MMA commit -> softmax, epilogue)
  B5_OFREE = 192,    // [2] O of a finished tile has been read out                  (4 epilogue arrivals -> MMA)
  B5_STATS = 208,    // [2] row statistics of a finished tile written               (4 softmax arrivals -> epilogue)
  B5_IFULL = 224,    // [8] step info published                                     (loader -> everybody)
  B5_IFREE = 288,    // [8] step info read: MMA issuer + 8 softmax warps + 4 epilogue warps = 13 arrivals -> loader
  B5_TMEM = 352
};
// step info, 32 bytes: {A.lo, A.hi, B.lo, B.hi} {b, h, step flags, 0}; per slot lo = q0 | j << 16 | ncols << 24,
// hi = slot flags | K/V stage << 8 | K/V phase << 12 | Q buffer << 13.  ncols: key columns of the block some row of the tile can see,
// rounded up to 16.
enum { G5_ACT = 1, G5_FIRST = 2, G5_LAST = 4 };     // slot flags
enum { G5_END = 1 };                                 // step flags

__global__ void __launch_bounds__(F5_THREADS, 1)
ot_attn_fwd_v5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ AttnFwdV5KParams p) {
  constexpr int DH = F5_DH;
  constexpr int SWB = 128;
  using T = AttnTile<DH, SWB>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + F5_OFF_BARS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((sbase & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    auto init = [&](uint32_t off, int n, uint32_t count) {
      for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + off + 8 * i), "r"(count));
    };
    init(B5_Q, 4, 1); init(B5_QFREE, 4, 1); init(B5_KV, 4, 1); init(B5_KVFREE, 4, 1); init(B5_S, 2, 1); init(B5_SFREE, 2, 4);
    init(B5_P, 2, 4); init(B5_O, 2, 1); init(B5_OFREE, 2, 4); init(B5_STATS, 2, 4); init(B5_IFULL, 8, 1); init(B5_IFREE, 8, 13);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(reinterpret_cast<uint32_t*>(smem + F5_OFF_BARS + B5_TMEM), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + F5_OFF_BARS + B5_TMEM);
  const int off = p.Lk - p.Lq;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F5_CTRL_REGS));
    if (warp == 0) {
      // ============================== loader: balanced step schedule + TMA ==============================
      if (elect_one()) {
        uint32_t t = 0, n_loads = 0;
        uint32_t item[2] = {0, 0};                      // tiles started per slot (Q buffer = item & 1)
        const int n_bh = p.B * p.H;
        int bh = blockIdx.x;
        while (bh >= 0) {
          int next_bh = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : bh + (int)gridDim.x;
          if (next_bh >= n_bh) next_bh = -1;
          const int h = bh % p.H, b = bh / p.H;
          int next_k = 0;                               // tiles are handed out longest first: k-th longest = tile n_qt - 1 - k
          int q0s[2] = {0, 0}, js[2] = {0, 0}, nkvs[2] = {0, 0}, lastq[2] = {0, 0}, qbs[2] = {0, 0};
          bool busy[2] = {false, false};
          while (true) {
            bool fresh[2] = {false, false};
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (!busy[x] && next_k < p.n_qt) {        // this slot starts its next tile in this step
                const int q0 = (p.n_qt - 1 - next_k) * 128;
                ++next_k;
                q0s[x] = q0; js[x] = 0; lastq[x] = min(q0 + 127, p.Lq - 1);
                nkvs[x] = (off + lastq[x]) / 128 + 1;   // off + lastq <= Lk - 1
                qbs[x] = item[x] & 1;
                busy[x] = true; fresh[x] = true;
                const int qi = x * 2 + qbs[x];
                uint64_t* bq = reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_Q) + qi;
                if (item[x] >= 2) mbar_wait_s(bars + B5_QFREE + 8 * qi, ((item[x] >> 1) - 1) & 1);
                mbar_arrive_expect_tx(bq, F5_TILE);
                load_head_tile<DH, SWB>(smem + qi * F5_TILE, &tmQ, bq, h, b, q0);
                ++item[x];
              }
            }
            if (!busy[0] && !busy[1]) break;
            // K/V blocks of this step: one load when both slots want the same block
            const bool shared_kv = busy[0] && busy[1] && js[0] == js[1];
            uint32_t ld[2];
            ld[0] = n_loads;
            ld[1] = (busy[0] && busy[1] && !shared_kv) ? n_loads + 1 : n_loads;
            const uint32_t loads_here = (busy[0] && busy[1] && !shared_kv) ? 2u : 1u;
            int lo[2], hi[2];
            bool fin_step = true;                        // is this the (sample, head)'s last step?
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              lo[x] = 0; hi[x] = 0;
              if (busy[x]) {
                const int ncols = min(128, (off + lastq[x] - js[x] * 128 + 16) & ~15);      // (last visible column + 1) rounded up to 16
                const bool last = js[x] == nkvs[x] - 1;
                lo[x] = q0s[x] | (js[x] << 16) | (ncols << 24);
                hi[x] = G5_ACT | (fresh[x] ? G5_FIRST : 0) | (last ? G5_LAST : 0) | ((ld[x] & 3) << 8) | (((ld[x] >> 2) & 1) << 12) | (qbs[x] << 13);
                if (!last) fin_step = false;
              }
            }
            if (next_k < p.n_qt) fin_step = false;
            const int is = t & (F5_INFO_SLOTS - 1);
            if (t >= F5_INFO_SLOTS) mbar_wait_s(bars + B5_IFREE + 8 * is, ((t / F5_INFO_SLOTS) - 1) & 1);
            int4* ip = reinterpret_cast<int4*>(smem + F5_OFF_INFO + is * 32);
            ip[0] = make_int4(lo[0], hi[0], lo[1], hi[1]);
            ip[1] = make_int4(b, h, (fin_step && next_bh < 0) ? G5_END : 0, 0);
            mbar_arrive_s(bars + B5_IFULL + 8 * is);               // release: publishes the slot
            ++t;
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (!busy[x] || (x == 1 && shared_kv)) continue;
              const uint32_t L = ld[x];
              const int st = L & 3;
              uint64_t* bkv = reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_KV) + st;
              if (L >= F5_KV_STAGES) mbar_wait_s(bars + B5_KVFREE + 8 * st, ((L >> 2) - 1) & 1);
              mbar_arrive_expect_tx(bkv, 2 * F5_TILE);
              load_head_tile<DH, SWB>(smem + F5_OFF_K + st * F5_TILE, &tmK, bkv, h, b, js[x] * 128);
              load_head_tile<DH, SWB>(smem + F5_OFF_V + st * F5_TILE, &tmV, bkv, h, b, js[x] * 128);
            }
            n_loads += loads_here;
#pragma unroll
            for (int x = 0; x < 2; ++x)
              if (busy[x] && ++js[x] == nkvs[x]) busy[x] = false;
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
        uint32_t items[2] = {0, 0};        // tiles started per slot (phase counter of Q)
        uint32_t cnt[2] = {0, 0};          // active steps handled so far per slot (phase counter of SFREE / P)
        uint32_t fin[2] = {0, 0};          // finished tiles per slot (phase counter of OFREE)
        struct Step { int4 a; int sflags; };
        auto read_step = [&](uint32_t t) -> Step {
          const int is = t & (F5_INFO_SLOTS - 1);
          mbar_wait_s(bars + B5_IFULL + 8 * is, (t / F5_INFO_SLOTS) & 1);
          Step st;
          st.a = *reinterpret_cast<const int4*>(smem + F5_OFF_INFO + is * 32);
          st.sflags = reinterpret_cast<const int4*>(smem + F5_OFF_INFO + is * 32)[1].z;
          mbar_arrive_s(bars + B5_IFREE + 8 * is);
          return st;
        };
        auto issue_s = [&](int lo, int hi, int x) {       // S = Q K^T of the slot's block over the visible key columns
          const int stg = (hi >> 8) & 3, qi = x * 2 + ((hi >> 13) & 1);
          mbar_wait_s(bars + B5_KV + 8 * stg, (hi >> 12) & 1);
          if (hi & G5_FIRST) { mbar_wait_s(bars + B5_Q + 8 * qi, (items[x] >> 1) & 1); ++items[x]; }
          const uint32_t idesc_s = make_idesc_bf16(128, (lo >> 24) & 0xff, 0, 0);
          tc_fence_after();
          const uint64_t aQ = tileK + addr14(sbase + qi * F5_TILE), aK = tileK + addr14(sbase + F5_OFF_K + stg * F5_TILE);
          const uint32_t d = tmem_base + F5_T_S + x * 128;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(d, aQ + 2 * kk, aK + 2 * kk, idesc_s, kk != 0);
          umma_commit(reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_S) + x);
          if (hi & G5_LAST) umma_commit(reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_QFREE) + qi);   // the tile's last S product
        };
        auto issue_pv = [&](int lo, int hi, int x) {
          const int stg = (hi >> 8) & 3;
          mbar_wait_s(bars + B5_P + 8 * x, cnt[x] & 1);                      // P of the slot's block is in TMEM, O rescaled if needed
          if ((hi & G5_FIRST) && fin[x] > 0) mbar_wait_s(bars + B5_OFREE + 8 * x, (fin[x] - 1) & 1);   // previous tile's O is out
          tc_fence_after();
          const uint64_t mV = tileMN + addr14(sbase + F5_OFF_V + stg * F5_TILE);
          const uint32_t aP = tmem_base + F5_T_P + x * 64;
          const int nk = ((lo >> 24) & 0xff) >> 4;
          const bool first = hi & G5_FIRST;
#pragma unroll 1
          for (int kk = 0; kk < nk; ++kk)
            umma_bf16_ts(tmem_base + F5_T_O + x * DH, aP + 8 * kk, mV + 128 * kk, idesc_pv, (first && kk == 0) ? 0u : 1u);
          umma_commit(reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_O) + x);
          ++cnt[x];
          if (hi & G5_LAST) ++fin[x];
        };
        uint32_t t = 0;
        Step cur = read_step(0);
        if (cur.a.y & G5_ACT) issue_s(cur.a.x, cur.a.y, 0);
        if (cur.a.w & G5_ACT) issue_s(cur.a.z, cur.a.w, 1);
        bool end = false;
        while (!end) {
          end = (cur.sflags & G5_END) != 0;
          Step nxt = cur;
          if (!end) nxt = read_step(t + 1);
          // S of the next step as soon as this step's scores have left TMEM (it runs under this step's exponentials) ...
          if (cur.a.y & G5_ACT) mbar_wait_s(bars + B5_SFREE, cnt[0] & 1);
          if (!end && (nxt.a.y & G5_ACT)) issue_s(nxt.a.x, nxt.a.y, 0);
          if (cur.a.w & G5_ACT) mbar_wait_s(bars + B5_SFREE + 8, cnt[1] & 1);
          if (!end && (nxt.a.w & G5_ACT)) issue_s(nxt.a.z, nxt.a.w, 1);
          // ... then P V of this step for each slot when its probabilities arrive
          if (cur.a.y & G5_ACT) issue_pv(cur.a.x, cur.a.y, 0);
          if (cur.a.w & G5_ACT) issue_pv(cur.a.z, cur.a.w, 1);
          // every product that reads this step's K/V stage(s) has been issued
          const int sA = (cur.a.y >> 8) & 3, sB = (cur.a.w >> 8) & 3;
          if (cur.a.y & G5_ACT) umma_commit(reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_KVFREE) + sA);
          if ((cur.a.w & G5_ACT) && (!(cur.a.y & G5_ACT) || sB != sA)) umma_commit(reinterpret_cast<uint64_t*>(smem + F5_OFF_BARS + B5_KVFREE) + sB);
          cur = nxt;
          ++t;
        }
      }
    }
  } else if (warp < 8) {
    // ============================== epilogue warpgroup (warps 4-7): finishes tiles of both slots ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F5_EPI_REGS));
    const int wrow = (warp & 3) * 32;
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const int et = tid - 128;                     // 0..127
    uint32_t g = 0, n[2] = {0, 0}, fin[2] = {0, 0};
    bool end = false;
    bool store_pending = false;
    while (!end) {
      const int is = g & (F5_INFO_SLOTS - 1);
      mbar_wait_s(bars + B5_IFULL + 8 * is, (g / F5_INFO_SLOTS) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + F5_OFF_INFO + is * 32);
      const int4 sc = reinterpret_cast<const int4*>(smem + F5_OFF_INFO + is * 32)[1];
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bars + B5_IFREE + 8 * is);
      end = (sc.z & G5_END) != 0;
      ++g;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int lo = x ? si.z : si.x, hi = x ? si.w : si.y;
        if (!(hi & G5_ACT)) continue;
        ++n[x];
        if (!(hi & G5_LAST)) continue;
        const int q0 = lo & 0xffff;
        const int b = sc.x, h = sc.y;
        const uint32_t buf = fin[x] & 1;
        mbar_wait_s(bars + B5_STATS + 8 * x, fin[x] & 1);           // l, m of the tile's rows are in shared memory
        mbar_wait_s(bars + B5_O + 8 * x, (n[x] - 1) & 1);           // the tile's last P V has completed
        tc_fence_after();
        uint32_t w0[32], w1[32];
        tmem_ld_x32(t_lane + F5_T_O + x * DH, w0);
        tmem_ld_x32(t_lane + F5_T_O + x * DH + 32, w1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bars + B5_OFREE + 8 * x);      // the next tile of this slot may overwrite O
        const float2 lm = *reinterpret_cast<const float2*>(smem + F5_OFF_STATS + ((x * 2 + buf) * 128 + row) * 8);
        const float inv = 1.0f / lm.x;
        if (q0 + row < p.Lq) p.lse[((long long)b * p.H + h) * p.Lq + q0 + row] = lm.y * p.scale + logf(lm.x);
        // the staging tile is free once the previous tile's TMA store has read it
        if (store_pending) {
          if (et == 0) bulk_wait_read0();
          named_bar_sync(3, 128);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          *reinterpret_cast<uint4*>(smem + F5_OFF_STG + swz_off<128>(row, ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(w0[ch * 8 + 0]) * inv, __uint_as_float(w0[ch * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 2]) * inv, __uint_as_float(w0[ch * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 4]) * inv, __uint_as_float(w0[ch * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(w0[ch * 8 + 6]) * inv, __uint_as_float(w0[ch * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(smem + F5_OFF_STG + swz_off<128>(row, 4 + ch)) =
              make_uint4(pack_bf16x2(__uint_as_float(w1[ch * 8 + 0]) * inv, __uint_as_float(w1[ch * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 2]) * inv, __uint_as_float(w1[ch * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 4]) * inv, __uint_as_float(w1[ch * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(w1[ch * 8 + 6]) * inv, __uint_as_float(w1[ch * 8 + 7]) * inv));
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (et == 0) {
          tma_store_3d(&tmO, sbase + F5_OFF_STG, h * DH, b, q0);    // rows past Lq are clipped by the tensor map
          bulk_commit();
        }
        store_pending = true;
        ++fin[x];
      }
    }
    if (et == 0) bulk_wait_all();
  } else {
    // ============================== softmax warpgroups (warps 8-11: tile A, 12-15: tile B) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F5_SM_REGS));
    const int x = (warp >> 2) - 2;                 // tile slot
    const int wrow = (warp & 3) * 32;              // first tile row of this warp == first TMEM lane
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const uint32_t t_s = t_lane + F5_T_S + x * 128, t_o = t_lane + F5_T_O + x * DH, t_p = t_lane + F5_T_P + x * 64;
    const uint32_t bar_s = bars + B5_S + 8 * x, bar_sfree = bars + B5_SFREE + 8 * x, bar_p = bars + B5_P + 8 * x;
    const uint32_t bar_o = bars + B5_O + 8 * x, bar_stats = bars + B5_STATS + 8 * x;
    uint32_t g = 0;          // global step counter (every step of the CTA)
    uint32_t n = 0;          // steps of this tile slot so far (phase counter of S / O)
    uint32_t fin = 0;        // finished tiles of this slot
    bool end = false;
    float m_ref = -INFINITY, l_run = 0.0f;

    while (!end) {
      const int is = g & (F5_INFO_SLOTS - 1);
      mbar_wait_s(bars + B5_IFULL + 8 * is, (g / F5_INFO_SLOTS) & 1);
      const int2 sr = reinterpret_cast<const int2*>(smem + F5_OFF_INFO + is * 32)[x];      // this slot's record
      const int sflags = reinterpret_cast<const int4*>(smem + F5_OFF_INFO + is * 32)[1].z;
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bars + B5_IFREE + 8 * is);
      end = (sflags & G5_END) != 0;
      ++g;
      if (!(sr.y & G5_ACT)) continue;
      const bool first = sr.y & G5_FIRST, last = sr.y & G5_LAST;
      const int j128 = ((sr.x >> 16) & 0xff) * 128;
      const int q0 = sr.x & 0xffff;
      const int ncols = (sr.x >> 24) & 0xff;
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
        *reinterpret_cast<float2*>(smem + F5_OFF_STATS + ((x * 2 + (fin & 1)) * 128 + row) * 8) = make_float2(l_run, m_ref);
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

int attn_fwd_v5_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_head_tmap(&tq, p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tk, p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tv, p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  if ((rc = make_head_tmap(&to, p->o, cols, p->B, p->Lq, p->ldo, 128))) return rc;
  AttnFwdV5KParams kp;
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk; kp.n_qt = (p->Lq + 127) / 128;
  kp.n_pairs = (kp.n_qt + 1) / 2;
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.lse = p->lse;
  kp.sched = sched_slot(st);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_fwd_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F5_SMEM_BYTES));
    attr_done = true;
  }
  const int sms = num_sms();
  const int n_bh = p->B * p->H;
  const int grid = n_bh < sms ? n_bh : sms;
  ot_attn_fwd_v5_kernel<<<grid, F5_THREADS, F5_SMEM_BYTES, st>>>(tq, tk, tv, to, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
