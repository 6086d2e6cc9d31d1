// ot_attn_bwd_v2.cu — single-pass backward of the pruned causal attention for head_dim 64 (tape.gradient of OT/model.py:101-114),
// KEY-MAJOR form with the probability tiles in TENSOR MEMORY.  Why (profiles/README.md round 2, profiles/exp_mma_rate.cu): the
// phase timers of ot_attn_bwd_fused showed the tensor pipe taking 4500 clk per (query tile, key tile) step for 1280 clk of
// arithmetic - the kernel is bound by shared-memory bandwidth (ncu: 73 % of the data pipe): five products with both operands in
// shared memory (208 KB per step), P and dS written as two 32 KB bf16 tiles, Q / dO arriving by TMA, dQ staged for its
// reductions - 336 KB per step against 128 B/clk.  Here the scores are computed TRANSPOSED,
//     S^T = K_j Q_i^T,   dP^T = V_j dO_i^T          (TMEM lanes = keys, columns = queries, in two 64-query halves)
//     P^T = exp2(S^T c - lse[q]),   dS^T = P^T o (dP^T - delta[q])          (row statistics come from shared memory)
// so that P^T and dS^T can stay in tensor memory as the A operands (TS form) of
//     dV_j += P^T dO_i,   dK_j += dS^T Q_i          (no P tile in shared memory at all, dS^T written once)
//     dQ_i  = dS K_j                                 (A = the dS^T tile read MN-major from shared memory)
// which leaves 144 KB of operand reads + 32 KB of dS^T + TMA per step.  The 64-query halves are double-buffered in TMEM: the
// element-wise warps work on one half while the tensor pipe runs the products of the other.  The softmax scale is applied to dQ
// and dK on the way out (a power of two for head_dim 64: exact).
//
// Warps: 0 loader (TMA + step ring), 1 MMA issuer, 2-3 row statistics (-lse * log2 e, -delta -> shared memory), 4-11 element-wise
// (thread = one key row x 32 query columns of a half), 12-15 output (dQ by TMA reduce-add, dK / dV by TMA store).
// TMEM: S^T halves 0 / 64, dP^T halves 128 / 192, dV 256, dK 320, dQ 384.
#include <stdlib.h>
#include <string.h>
#include "ot_attn.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

struct AttnBwdV2KParams {
  int B, H, Lq, Lk, n_qt, n_kt, total_items;
  float scale, scale_log2;
  const float* lse;    // [B,H,Lq]
  const float* delta;  // [B,H,Lq]
  int* sched;          // dynamic work counter (zeroed by the launcher) or NULL = static round-robin
};

// NCG = column groups per 64-query half = element-wise warps per TMEM lane quarter (2: thread = key row x 32 queries, 8 warps;
// 4: thread = key row x 16 queries, 16 warps = four per scheduler, which hides the MUFU / FMA latencies the two-warp form exposed).
template <int NCG> struct B2Cfg;
template <> struct B2Cfg<2> { static constexpr int THREADS = 512, CTRL_REGS = 56, OUT_REGS = 88, EW_REGS = 184; };   // 128 regs at launch: 128*72 + 128*40 released = 256*56 taken
template <> struct B2Cfg<4> { static constexpr int THREADS = 768, CTRL_REGS = 40, OUT_REGS = 88, EW_REGS = 88; };    //  80 regs at launch: 128*40 released = 128*8 + 512*8 taken
static constexpr int B2_DH = 64;
static constexpr int B2_TILE = 128 * B2_DH * 2;        // 16 KB
static constexpr int B2_Q_STAGES = 3;
static constexpr int B2_INFO_SLOTS = 8;
static constexpr int B2_LOOKAHEAD = 5;       // steps the L2 prefetches run ahead of the TMA loads
static constexpr int B2_OFF_K = 0;                                   // [2]
static constexpr int B2_OFF_V = B2_OFF_K + 2 * B2_TILE;              // [2]
static constexpr int B2_OFF_Q = B2_OFF_V + 2 * B2_TILE;              // [3]
static constexpr int B2_OFF_DO = B2_OFF_Q + B2_Q_STAGES * B2_TILE;   // [3]
static constexpr int B2_OFF_DS = B2_OFF_DO + B2_Q_STAGES * B2_TILE;  // dS^T [128 keys x 128 queries] bf16, two 64-query slabs
static constexpr int B2_OFF_STG = B2_OFF_DS + PT_BYTES;              // output staging tile
static constexpr int B2_OFF_STATS = B2_OFF_STG + B2_TILE;            // [3][2][128] fp32: lse * log2 e, delta
static constexpr int B2_OFF_INFO = B2_OFF_STATS + B2_Q_STAGES * 2 * 128 * 4;
static constexpr int B2_OFF_BARS = B2_OFF_INFO + B2_INFO_SLOTS * 16;
static constexpr int B2_SMEM_BYTES = B2_OFF_BARS + 512;
static_assert(B2_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t B2_T_S = 0, B2_T_DP = 128, B2_T_DV = 256, B2_T_DK = 320, B2_T_DQ = 384, B2_T_K = 448, B2_T_V = 480;
#ifndef OT_B2_KV_TMEM
#define OT_B2_KV_TMEM 1     // 1: K_j / V_j are copied to tensor memory once per item (tcgen05.cp) and S^T / dP^T run in TS form; 0: SS form
#endif
enum : uint32_t {
  BB_KV = 0,         // [2] K_j, V_j landed                                  (loader -> MMA)
  BB_KVFREE = 16,    // [2] every product of the item is complete            (MMA commit -> loader)
  BB_Q = 32,         // [3] Q_i, dO_i landed and row statistics written      (loader + 2 statistics warps -> MMA, element-wise)
  BB_QFREE = 56,     // [3] dV / dK products of the step's second half done  (MMA commit -> loader)
  BB_S = 80,         // [2] S^T, dP^T of a half complete                     (MMA commit -> element-wise)
  BB_PDS = 96,       // [2] P^T, dS^T of a half written to TENSOR memory     (4 NCG arrivals -> MMA: dV / dK products may go)
  BB_PDSS = 280,     //     dS^T of the step written to SHARED memory        (4 NCG arrivals -> MMA: dQ product may go)
  BB_DQ = 112,       //     dQ of the step complete                          (MMA commit -> output, element-wise)
  BB_DQFREE = 120,   //     dQ pulled out of TMEM                            (4 arrivals -> MMA)
  BB_ACC = 128,      //     dV, dK of the item complete                      (MMA commit -> output)
  BB_ACCFREE = 136,  //     dV, dK pulled out of TMEM                        (4 arrivals -> MMA)
  BB_IFULL = 144,    // [8] step info published                              (loader -> everybody)
  BB_IFREE = 208,    // [8] step info read: MMA 1 + statistics 2 + element-wise 4 NCG + output 4 arrivals -> loader
  BB_TMEM = 272
};
enum { SB_FIRST = 1, SB_LAST = 2, SB_END = 4, SB_KVBUF = 8 };
// step info (one query tile against the item's key tile), 16 bytes: x = q0 | k0 << 16, y = b, z = h | flags << 8 | stage << 16, w = 0

__device__ __forceinline__ void b2_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
#ifndef OT_B2_WAIT_HINT_NS
#define OT_B2_WAIT_HINT_NS 20000
#endif
__device__ __forceinline__ void b2_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (ok) return;
  uint32_t spins = 0;
  do {
#if OT_B2_WAIT_HINT_NS > 0
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)OT_B2_WAIT_HINT_NS) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
#if OT_HANG_GUARD
    if (++spins > (1u << 24)) __trap();
#endif
  } while (!ok);
}
__device__ __forceinline__ void b2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared memory -> tensor memory, 128 lanes x 32 bytes (one K = 16 step of a K-major bf16 operand); executes in issue order with the MMAs
__device__ __forceinline__ void b2_tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void b2_tma_prefetch_l2(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1),
               "r"(c2) : "memory");
}
__device__ __forceinline__ void b2_prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void b2_tma_store(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void b2_tma_reduce_add(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Walks the (item, query tile) steps of one CTA in launch order (loader warp only): item = (key tile, head, sample).
// Items after the first are drawn from a device-wide counter.  The loader runs a second, LOOK-AHEAD cursor a few steps in front of
// the one that issues the TMA loads: it draws the items (and hands their ids to the main cursor through a small ring) and asks for
// L2 prefetches of the tiles, so that the loads proper find their 128-byte rows (each in a different DRAM page: consecutive tokens
// of a sample are B * ld elements apart) in L2.  ncu before: the element-wise warps waited 6 % on the step's Q / dO / statistics.
struct B2ItemRing { int ids[8]; uint32_t w, r; };
struct B2Cursor {
  int item, next_item, ii, n_i, i_min, h, b, k0, item_idx;
  bool valid;
  __device__ __forceinline__ void fetch_next(const AttnBwdV2KParams& p, B2ItemRing& ring, bool producer) {
    if (producer) {
      next_item = p.sched != nullptr ? (int)gridDim.x + atomicAdd(p.sched, 1) : item + (int)gridDim.x;
      ring.ids[ring.w++ & 7] = next_item;
    } else {
      next_item = ring.ids[ring.r++ & 7];
    }
  }
  __device__ __forceinline__ void load_item(const AttnBwdV2KParams& p) {
    valid = item < p.total_items;
    if (!valid) return;
    const int kt = item % p.n_kt;
    const int bh = item / p.n_kt;
    h = bh % p.H;
    b = bh / p.H;
    k0 = kt * 128;
    const int off = p.Lk - p.Lq;
    i_min = (k0 - off) < 0 ? 0 : (k0 - off) / 128;     // first query tile that can see some key of this tile
    n_i = p.n_qt - i_min;
    ii = 0;
  }
  __device__ __forceinline__ void init(const AttnBwdV2KParams& p, B2ItemRing& ring, bool producer) {
    item = blockIdx.x;
    next_item = p.total_items;
    item_idx = 0;
    load_item(p);
    if (valid) fetch_next(p, ring, producer);
  }
  __device__ __forceinline__ void next(const AttnBwdV2KParams& p, B2ItemRing& ring, bool producer) {
    if (++ii == n_i) {
      item = next_item;
      ++item_idx;
      load_item(p);
      if (valid) fetch_next(p, ring, producer);
    }
  }
  __device__ __forceinline__ int q0() const { return (i_min + ii) * 128; }
};

template <int NCG>
__global__ void __launch_bounds__(B2Cfg<NCG>::THREADS, 1)
ot_attn_bwd_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
                      const __grid_constant__ CUtensorMap tmdV, const __grid_constant__ AttnBwdV2KParams p) {
  constexpr int DH = B2_DH;
  constexpr int SWB = 128;
  using Cfg = B2Cfg<NCG>;
  constexpr int CW = 64 / NCG;                  // query columns per element-wise thread and half
  constexpr int N_EW = 4 * NCG;                 // element-wise warps
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + B2_OFF_BARS;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    if ((sbase & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQ); tma_prefetch_desc(&tmdK); tma_prefetch_desc(&tmdV);
    auto init = [&](uint32_t off, int n, uint32_t count) {
      for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + off + 8 * i), "r"(count));
    };
    init(BB_KV, 2, 1); init(BB_KVFREE, 2, 1); init(BB_Q, 3, 3); init(BB_QFREE, 3, 1); init(BB_S, 2, 1); init(BB_PDS, 2, N_EW); init(BB_PDSS, 1, N_EW);
    init(BB_DQ, 1, 1); init(BB_DQFREE, 1, 4); init(BB_ACC, 1, 1); init(BB_ACCFREE, 1, 4); init(BB_IFULL, 8, 1); init(BB_IFREE, 8, 7 + N_EW);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(reinterpret_cast<uint32_t*>(smem + B2_OFF_BARS + BB_TMEM), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + B2_OFF_BARS + BB_TMEM);
  const int off = p.Lk - p.Lq;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::CTRL_REGS));
    if (warp == 0) {
      // ============================== loader: step ring + TMA ==============================
      if (elect_one()) {
        B2ItemRing ring;
        ring.w = ring.r = 0;
        B2Cursor pf, c;
        pf.init(p, ring, true);
        c.init(p, ring, false);
        uint32_t t = 0, t_pf = 0;
        while (c.valid) {
          // look-ahead: L2 prefetch of the tiles (and statistics) of the next steps
          while (pf.valid && t_pf < t + B2_LOOKAHEAD) {
            const int pq0 = pf.q0();
            if (pf.ii == 0) {
              b2_tma_prefetch_l2(&tmK, pf.h * DH, pf.b, pf.k0);
              b2_tma_prefetch_l2(&tmV, pf.h * DH, pf.b, pf.k0);
            }
            b2_tma_prefetch_l2(&tmQ, pf.h * DH, pf.b, pq0);
            b2_tma_prefetch_l2(&tmdO, pf.h * DH, pf.b, pq0);
            const long long sb = ((long long)pf.b * p.H + pf.h) * p.Lq + pq0;
#pragma unroll
            for (int i = 0; i < 128; i += 32) {
              if (pq0 + i < p.Lq) { b2_prefetch_l2(p.lse + sb + i); b2_prefetch_l2(p.delta + sb + i); }
            }
            pf.next(p, ring, true);
            ++t_pf;
          }
          const int st = t % B2_Q_STAGES;
          const int kb = c.item_idx & 1;
          const bool first = c.ii == 0, last = c.ii == c.n_i - 1;
          if (t >= B2_Q_STAGES) b2_wait(bars + BB_QFREE + 8 * st, ((t / B2_Q_STAGES) - 1) & 1);      // stage's previous step done
          if (first && c.item_idx >= 2) b2_wait(bars + BB_KVFREE + 8 * kb, ((c.item_idx >> 1) - 1) & 1);   // buffer's previous item done
          B2Cursor n = c;
          B2ItemRing ring_n = ring;             // peek: the copy's ring position is discarded
          n.next(p, ring_n, false);
          const int is = t & (B2_INFO_SLOTS - 1);
          if (t >= B2_INFO_SLOTS) b2_wait(bars + BB_IFREE + 8 * is, ((t / B2_INFO_SLOTS) - 1) & 1);
          const int flags = (first ? SB_FIRST : 0) | (last ? SB_LAST : 0) | (n.valid ? 0 : SB_END) | (kb ? SB_KVBUF : 0);
          const int q0 = c.q0();
          *reinterpret_cast<int4*>(smem + B2_OFF_INFO + is * 16) = make_int4(q0 | (c.k0 << 16), c.b, c.h | (flags << 8) | (st << 16), 0);
          b2_arrive(bars + BB_IFULL + 8 * is);         // release: publishes the slot (the stage is free: the statistics warps may fill it)
          if (first) {
            uint64_t* bkv = reinterpret_cast<uint64_t*>(smem + B2_OFF_BARS + BB_KV) + kb;
            mbar_arrive_expect_tx(bkv, 2 * B2_TILE);
            load_head_tile<DH, SWB>(smem + B2_OFF_K + kb * B2_TILE, &tmK, bkv, c.h, c.b, c.k0);
            load_head_tile<DH, SWB>(smem + B2_OFF_V + kb * B2_TILE, &tmV, bkv, c.h, c.b, c.k0);
          }
          uint64_t* bq = reinterpret_cast<uint64_t*>(smem + B2_OFF_BARS + BB_Q) + st;
          mbar_arrive_expect_tx(bq, 2 * B2_TILE);
          load_head_tile<DH, SWB>(smem + B2_OFF_Q + st * B2_TILE, &tmQ, bq, c.h, c.b, q0);
          load_head_tile<DH, SWB>(smem + B2_OFF_DO + st * B2_TILE, &tmdO, bq, c.h, c.b, q0);
          c.next(p, ring, false);
          ++t;
        }
      }
    } else if (warp == 1) {
      // ============================== MMA issuer ==============================
      if (elect_one()) {
        constexpr uint32_t idesc_sdp = make_idesc_bf16(128, 64, 0, 0);   // [128 keys x 64 queries] = K (K-major) x Q half (K-major)
        constexpr uint32_t idesc_ts = make_idesc_bf16(128, DH, 0, 1);    // A = P^T / dS^T in TMEM, B = dO / Q half, MN-major
        constexpr uint32_t idesc_dq = make_idesc_bf16(128, DH, 1, 1);    // A = dS^T tile MN-major, B = K_j MN-major
        const uint64_t tileK = make_smem_desc<SWB>(0, 16);
        const uint64_t tileMN = make_smem_desc<SWB>(0, 128 * SWB);
        const uint64_t dS_mn = make_smem_desc<128>(sbase + B2_OFF_DS, PT_SLAB_BYTES);
        auto addr14 = [](uint32_t a) -> uint64_t { return static_cast<uint64_t>((a & 0x3FFFFu) >> 4); };
        uint32_t n_items = 0, items_main = 0;
        // step info + operands of step g
        auto fetch = [&](uint32_t g) -> int4 {
          const int is = g & (B2_INFO_SLOTS - 1);
          b2_wait(bars + BB_IFULL + 8 * is, (g / B2_INFO_SLOTS) & 1);
          const int4 si = *reinterpret_cast<const int4*>(smem + B2_OFF_INFO + is * 16);
          b2_arrive(bars + BB_IFREE + 8 * is);
          const int flags = (si.z >> 8) & 0xff, st = (si.z >> 16) & 0xff;
          b2_wait(bars + BB_Q + 8 * st, (g / B2_Q_STAGES) & 1);
          if (flags & SB_FIRST) {
            b2_wait(bars + BB_KV + 8 * ((flags & SB_KVBUF) ? 1 : 0), (n_items >> 1) & 1);
            ++n_items;
          }
          return si;
        };
        auto issue_sdp = [&](const int4& si, int hh) {   // S^T, dP^T of query half hh into buffer hh
          const int flags = (si.z >> 8) & 0xff, st = (si.z >> 16) & 0xff, kb = (flags & SB_KVBUF) ? 1 : 0;
          tc_fence_after();
          const uint64_t aK = tileK + addr14(sbase + B2_OFF_K + kb * B2_TILE), aV = tileK + addr14(sbase + B2_OFF_V + kb * B2_TILE);
          const uint64_t bQ = tileK + addr14(sbase + B2_OFF_Q + st * B2_TILE + hh * 8192);
          const uint64_t bdO = tileK + addr14(sbase + B2_OFF_DO + st * B2_TILE + hh * 8192);
#if OT_B2_KV_TMEM
          if (hh == 0 && (flags & SB_FIRST)) {      // new item: K_j, V_j into tensor memory behind the previous item's last S^T / dP^T products
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk) {
              b2_tmem_cp_128x256b(tmem_base + B2_T_K + 8 * kk, aK + 2 * kk);
              b2_tmem_cp_128x256b(tmem_base + B2_T_V + 8 * kk, aV + 2 * kk);
            }
          }
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ts(tmem_base + B2_T_S + hh * 64, tmem_base + B2_T_K + 8 * kk, bQ + 2 * kk, idesc_sdp, kk != 0);
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ts(tmem_base + B2_T_DP + hh * 64, tmem_base + B2_T_V + 8 * kk, bdO + 2 * kk, idesc_sdp, kk != 0);
#else
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(tmem_base + B2_T_S + hh * 64, aK + 2 * kk, bQ + 2 * kk, idesc_sdp, kk != 0);
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16_ss(tmem_base + B2_T_DP + hh * 64, aV + 2 * kk, bdO + 2 * kk, idesc_sdp, kk != 0);
#endif
          b2_commit(bars + BB_S + 8 * hh);
        };
        auto issue_main = [&](const int4& si, int hh, uint32_t g) {   // dV += P^T dO, dK += dS^T Q over the 64 queries of half hh
          const int flags = (si.z >> 8) & 0xff, st = (si.z >> 16) & 0xff;
          b2_wait(bars + BB_PDS + 8 * hh, g & 1);
          if (hh == 0 && (flags & SB_FIRST)) {
            if (items_main > 0) b2_wait(bars + BB_ACCFREE, (items_main - 1) & 1);   // previous item's dV / dK are out
            ++items_main;
          }
          tc_fence_after();
          const uint32_t fresh = (hh == 0 && (flags & SB_FIRST)) ? 1u : 0u;
          const uint64_t mdO = tileMN + addr14(sbase + B2_OFF_DO + st * B2_TILE + hh * 8192);
          const uint64_t mQ = tileMN + addr14(sbase + B2_OFF_Q + st * B2_TILE + hh * 8192);
          // 16 queries per K step; a thread's CW columns hold their CW / 2 packed columns at the start of its own column range
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tmem_base + B2_T_DV, tmem_base + B2_T_S + hh * 64 + CW * ((16 * kk) / CW) + 8 * (((16 * kk) % CW) / 16), mdO + 128 * kk,
                         idesc_ts, (fresh && kk == 0) ? 0u : 1u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(tmem_base + B2_T_DK, tmem_base + B2_T_DP + hh * 64 + CW * ((16 * kk) / CW) + 8 * (((16 * kk) % CW) / 16), mQ + 128 * kk,
                         idesc_ts, (fresh && kk == 0) ? 0u : 1u);
        };
        uint32_t g = 0;
        int4 cur = fetch(0);
        issue_sdp(cur, 0);
        issue_sdp(cur, 1);
        bool end = false;
        while (!end) {
          const int flags = (cur.z >> 8) & 0xff, st = (cur.z >> 16) & 0xff, kb = (flags & SB_KVBUF) ? 1 : 0;
          end = (flags & SB_END) != 0;
          issue_main(cur, 0, g);
          int4 nxt = cur;
          if (!end) {
            nxt = fetch(g + 1);
            issue_sdp(nxt, 0);                        // overwrites buffer 0 behind the products that read it (issue order)
          }
          issue_main(cur, 1, g);
          b2_commit(bars + BB_QFREE + 8 * st);        // Q_i / dO_i (and the statistics of the step) are no longer needed
          if (flags & SB_LAST) b2_commit(bars + BB_ACC);   // dV_j, dK_j are final
          if (!end) issue_sdp(nxt, 1);
          // dQ_i = dS K_j over the whole query tile (both halves of dS^T are in shared memory)
          b2_wait(bars + BB_PDSS, g & 1);
          if (g > 0) b2_wait(bars + BB_DQFREE, (g - 1) & 1);
          tc_fence_after();
          {
            const uint64_t mK = tileMN + addr14(sbase + B2_OFF_K + kb * B2_TILE);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_bf16_ss(tmem_base + B2_T_DQ, dS_mn + 128 * kk, mK + 128 * kk, idesc_dq, kk != 0);
          }
          b2_commit(bars + BB_DQ);
          if (flags & SB_LAST) b2_commit(bars + BB_KVFREE + 8 * kb);
          cur = nxt;
          ++g;
        }
      }
    } else {
      // ============================== row statistics (warps 2-3): lse * log2 e and delta of the step's 128 queries ==============================
      const int i = tid - 64;          // 0..63: rows i and i + 64
      uint32_t g = 0;
      bool end = false;
      while (!end) {
        const int is = g & (B2_INFO_SLOTS - 1);
        b2_wait(bars + BB_IFULL + 8 * is, (g / B2_INFO_SLOTS) & 1);
        const int4 si = *reinterpret_cast<const int4*>(smem + B2_OFF_INFO + is * 16);
        __syncwarp();
        if (lane == 0) b2_arrive(bars + BB_IFREE + 8 * is);
        const int flags = (si.z >> 8) & 0xff, st = (si.z >> 16) & 0xff;
        end = (flags & SB_END) != 0;
        const int q0 = si.x & 0xffff, b = si.y, h = si.z & 0xff;
        float* sl = reinterpret_cast<float*>(smem + B2_OFF_STATS + st * 1024);
        const long long base = ((long long)b * p.H + h) * p.Lq + q0;
#pragma unroll
        for (int r = i; r < 128; r += 64) {
          const bool ok = (q0 + r) < p.Lq;       // rows past the query tail: probability 0 (lse = +inf), delta 0
          sl[r] = ok ? p.lse[base + r] * -1.4426950408889634f : -INFINITY;
          sl[128 + r] = ok ? -p.delta[base + r] : 0.0f;
        }
        __syncwarp();
        if (lane == 0) b2_arrive(bars + BB_Q + 8 * st);     // release: publishes the 64 rows this warp wrote
        ++g;
      }
    }
  } else if (warp < 4 + N_EW) {
    // ============================== element-wise warps: thread = key row x CW query columns of a half ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::EW_REGS));
    const int ew = warp - 4;
    const int cg = ew >> 2;                            // column group inside the 64-query half
    const int wrow = (ew & 3) * 32;                    // first key row of this warp == first TMEM lane
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    uint32_t g = 0;
    bool end = false;
    while (!end) {
      const int is = g & (B2_INFO_SLOTS - 1);
      b2_wait(bars + BB_IFULL + 8 * is, (g / B2_INFO_SLOTS) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + B2_OFF_INFO + is * 16);
      __syncwarp();
      if (lane == 0) b2_arrive(bars + BB_IFREE + 8 * is);
      const int flags = (si.z >> 8) & 0xff, st = (si.z >> 16) & 0xff;
      end = (flags & SB_END) != 0;
      const int q0 = si.x & 0xffff, k0 = (si.x >> 16) & 0xffff;
      b2_wait(bars + BB_Q + 8 * st, (g / B2_Q_STAGES) & 1);         // row statistics of the step are in shared memory
      const float* sl = reinterpret_cast<const float*>(smem + B2_OFF_STATS + st * 1024);   // -lse * log2 e [128], -delta [128]
      uint32_t dk[2][CW / 2];                                       // dS^T of both halves, kept for the shared-memory copy at the end of the step
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        // column c (0 .. CW-1) of this thread is query q0 + 64 hh + CW cg + c; key k0 + row sees it iff c >= cmin
        const int cmin_lo = k0 + wrow - off - q0 - 64 * hh - CW * cg;      // lane 0 (lane 31: + 31)
        const int cmin = cmin_lo + lane;
        const bool none = cmin_lo >= CW;                                   // the whole 32 x CW block is hidden
        const bool fast = cmin_lo + 31 <= 0;                               // entirely visible
        const uint32_t t_s = t_lane + B2_T_S + hh * 64 + cg * CW, t_dp = t_lane + B2_T_DP + hh * 64 + cg * CW;
        uint32_t pk[CW / 2];
        b2_wait(bars + BB_S + 8 * hh, g & 1);
        tc_fence_after();
        if (none) {
#pragma unroll
          for (int i = 0; i < CW / 2; ++i) { pk[i] = 0u; dk[hh][i] = 0u; }
        } else {
          uint32_t vs[CW], vd[CW];
          if (CW == 32) {
            tmem_ld_x32(t_s, reinterpret_cast<uint32_t (&)[32]>(vs));
            tmem_ld_x32(t_dp, reinterpret_cast<uint32_t (&)[32]>(vd));
          } else {
            tmem_ld_x16(t_s, reinterpret_cast<uint32_t (&)[16]>(vs));
            tmem_ld_x16(t_dp, reinterpret_cast<uint32_t (&)[16]>(vd));
          }
          tmem_ld_wait();
          const float4* l4 = reinterpret_cast<const float4*>(sl + hh * 64 + cg * CW);
          const float4* d4 = reinterpret_cast<const float4*>(sl + 128 + hh * 64 + cg * CW);
          if (fast) {
            // packed fp32 pairs: 2 FFMA2 + 4 MUFU + 2 FADD2 + 2 FMUL2 + 4 packs per four elements (the scalar form issued 12 + 4 + 4)
            const f32x2 c2 = pk2(p.scale_log2);
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
              const float4 nl = l4[c4];
              const float4 nd = d4[c4];
              float a0, a1, a2, a3;
              upk2(fma2(pk2(__uint_as_float(vs[4 * c4 + 0]), __uint_as_float(vs[4 * c4 + 1])), c2, pk2(nl.x, nl.y)), a0, a1);
              upk2(fma2(pk2(__uint_as_float(vs[4 * c4 + 2]), __uint_as_float(vs[4 * c4 + 3])), c2, pk2(nl.z, nl.w)), a2, a3);
              const float p0 = ex2_approx(a0), p1 = ex2_approx(a1), p2 = ex2_approx(a2), p3 = ex2_approx(a3);
              const f32x2 t01 = add2(pk2(__uint_as_float(vd[4 * c4 + 0]), __uint_as_float(vd[4 * c4 + 1])), pk2(nd.x, nd.y));
              const f32x2 t23 = add2(pk2(__uint_as_float(vd[4 * c4 + 2]), __uint_as_float(vd[4 * c4 + 3])), pk2(nd.z, nd.w));
              float s0, s1, s2, s3;
              upk2(mul2(pk2(p0, p1), t01), s0, s1);
              upk2(mul2(pk2(p2, p3), t23), s2, s3);
              pk[2 * c4] = pack_bf16x2(p0, p1);
              pk[2 * c4 + 1] = pack_bf16x2(p2, p3);
              dk[hh][2 * c4] = pack_bf16x2(s0, s1);
              dk[hh][2 * c4 + 1] = pack_bf16x2(s2, s3);
            }
          } else {
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
              const float4 nl = l4[c4];
              const float4 nd = d4[c4];
              float p0 = ex2_approx(fmaf(__uint_as_float(vs[4 * c4 + 0]), p.scale_log2, nl.x));
              float p1 = ex2_approx(fmaf(__uint_as_float(vs[4 * c4 + 1]), p.scale_log2, nl.y));
              float p2 = ex2_approx(fmaf(__uint_as_float(vs[4 * c4 + 2]), p.scale_log2, nl.z));
              float p3 = ex2_approx(fmaf(__uint_as_float(vs[4 * c4 + 3]), p.scale_log2, nl.w));
              p0 = (4 * c4 + 0 >= cmin) ? p0 : 0.0f;
              p1 = (4 * c4 + 1 >= cmin) ? p1 : 0.0f;
              p2 = (4 * c4 + 2 >= cmin) ? p2 : 0.0f;
              p3 = (4 * c4 + 3 >= cmin) ? p3 : 0.0f;
              const float s0 = p0 * (__uint_as_float(vd[4 * c4 + 0]) + nd.x);
              const float s1 = p1 * (__uint_as_float(vd[4 * c4 + 1]) + nd.y);
              const float s2 = p2 * (__uint_as_float(vd[4 * c4 + 2]) + nd.z);
              const float s3 = p3 * (__uint_as_float(vd[4 * c4 + 3]) + nd.w);
              pk[2 * c4] = pack_bf16x2(p0, p1);
              pk[2 * c4 + 1] = pack_bf16x2(p2, p3);
              dk[hh][2 * c4] = pack_bf16x2(s0, s1);
              dk[hh][2 * c4 + 1] = pack_bf16x2(s2, s3);
            }
          }
        }
        // P^T / dS^T replace the first CW / 2 of this thread's own CW S^T / dP^T columns (read out above): the dV / dK products may go
        if (CW == 32) {
          tmem_st_x16(t_s, reinterpret_cast<uint32_t (&)[16]>(pk));
          tmem_st_x16(t_dp, reinterpret_cast<uint32_t (&)[16]>(dk[hh]));
        } else {
          tmem_st_x8(t_s, reinterpret_cast<uint32_t (&)[8]>(pk));
          tmem_st_x8(t_dp, reinterpret_cast<uint32_t (&)[8]>(dk[hh]));
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) b2_arrive(bars + BB_PDS + 8 * hh);
      }
      // dS^T of both halves also goes to shared memory for the dQ product: once per step, off the critical path of the dV / dK
      // products, and late enough that the previous step's dQ product (the reader of the tile) is long complete
      if (g > 0) b2_wait(bars + BB_DQ, (g - 1) & 1);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint8_t* slab = smem + B2_OFF_DS + hh * PT_SLAB_BYTES;
#pragma unroll
        for (int ch = 0; ch < CW / 8; ++ch)
          *reinterpret_cast<uint4*>(slab + swz_off<128>(row, cg * (CW / 8) + ch)) =
              make_uint4(dk[hh][4 * ch], dk[hh][4 * ch + 1], dk[hh][4 * ch + 2], dk[hh][4 * ch + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) b2_arrive(bars + BB_PDSS);
      ++g;
    }
  } else {
    // ============================== output warps (the last four) ==============================
    if (NCG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::OUT_REGS));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::OUT_REGS));
    const int wrow = (warp & 3) * 32;
    const int row = wrow + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wrow) << 16);
    const int ot = tid - (Cfg::THREADS - 128);
    uint32_t g = 0, n_items = 0;
    bool end = false, pending = false;
    // TMEM accumulator tile (64 fp32 columns of this thread's row) x mul -> bf16 -> swizzled staging tile
    auto stage_tile = [&](uint32_t tcol, float mul, uint32_t free_bar) {
      uint32_t v0[32], v1[32];
      tmem_ld_x32(t_lane + tcol, v0);
      tmem_ld_x32(t_lane + tcol + 32, v1);
      tmem_ld_wait();
      if (free_bar != 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) b2_arrive(free_bar);
      }
      if (pending) {                                   // the previous TMA has finished reading the staging tile
        if (ot == 0) bulk_wait_read0();
        named_bar_sync(4, 128);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        *reinterpret_cast<uint4*>(smem + B2_OFF_STG + swz_off<128>(row, ch)) =
            make_uint4(pack_bf16x2(__uint_as_float(v0[ch * 8 + 0]) * mul, __uint_as_float(v0[ch * 8 + 1]) * mul),
                       pack_bf16x2(__uint_as_float(v0[ch * 8 + 2]) * mul, __uint_as_float(v0[ch * 8 + 3]) * mul),
                       pack_bf16x2(__uint_as_float(v0[ch * 8 + 4]) * mul, __uint_as_float(v0[ch * 8 + 5]) * mul),
                       pack_bf16x2(__uint_as_float(v0[ch * 8 + 6]) * mul, __uint_as_float(v0[ch * 8 + 7]) * mul));
        *reinterpret_cast<uint4*>(smem + B2_OFF_STG + swz_off<128>(row, 4 + ch)) =
            make_uint4(pack_bf16x2(__uint_as_float(v1[ch * 8 + 0]) * mul, __uint_as_float(v1[ch * 8 + 1]) * mul),
                       pack_bf16x2(__uint_as_float(v1[ch * 8 + 2]) * mul, __uint_as_float(v1[ch * 8 + 3]) * mul),
                       pack_bf16x2(__uint_as_float(v1[ch * 8 + 4]) * mul, __uint_as_float(v1[ch * 8 + 5]) * mul),
                       pack_bf16x2(__uint_as_float(v1[ch * 8 + 6]) * mul, __uint_as_float(v1[ch * 8 + 7]) * mul));
      }
      fence_proxy_async_smem();
      named_bar_sync(4, 128);
      pending = true;
    };
    while (!end) {
      const int is = g & (B2_INFO_SLOTS - 1);
      b2_wait(bars + BB_IFULL + 8 * is, (g / B2_INFO_SLOTS) & 1);
      const int4 si = *reinterpret_cast<const int4*>(smem + B2_OFF_INFO + is * 16);
      __syncwarp();
      if (lane == 0) b2_arrive(bars + BB_IFREE + 8 * is);
      const int flags = (si.z >> 8) & 0xff;
      end = (flags & SB_END) != 0;
      const int q0 = si.x & 0xffff, k0 = (si.x >> 16) & 0xffff, b = si.y, h = si.z & 0xff;
      if (flags & SB_LAST) {
        // end of an item: dV, dK first - the next item's first products wait for these columns, the dQ columns are not needed
        // again before the end of the next step
        b2_wait(bars + BB_ACC, n_items & 1);
        tc_fence_after();
        stage_tile(B2_T_DV, 1.0f, 0);
        if (ot == 0) { b2_tma_store(&tmdV, sbase + B2_OFF_STG, h * DH, b, k0); bulk_commit(); }
        stage_tile(B2_T_DK, p.scale, bars + BB_ACCFREE);
        if (ot == 0) { b2_tma_store(&tmdK, sbase + B2_OFF_STG, h * DH, b, k0); bulk_commit(); }
        ++n_items;
      }
      // dQ partial of this step: reduce-add into the zero-initialised dQ (rows past Lq are clipped by the tensor map)
      b2_wait(bars + BB_DQ, g & 1);
      tc_fence_after();
      stage_tile(B2_T_DQ, p.scale, bars + BB_DQFREE);
      if (ot == 0) { b2_tma_reduce_add(&tmdQ, sbase + B2_OFF_STG, h * DH, b, q0); bulk_commit(); }
      ++g;
    }
    if (ot == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_head_tmap(CUtensorMap* tm, const void* base, int cols, int B, int L, long long ld, int swb);
int attn_bwd_prologue(const ot_attn_params* p, cudaStream_t st);    // ot_attn_bwd_fused.cu: zero dQ, delta = rowsum(dO o O)

int attn_bwd_v2_impl(const ot_attn_params* p, cudaStream_t st) {
  const int cols = p->H * p->head_dim;
  CUtensorMap tm[7];
  int rc;
  if ((rc = make_head_tmap(&tm[0], p->q, cols, p->B, p->Lq, p->ldq, 128))) return rc;
  if ((rc = make_head_tmap(&tm[1], p->k, cols, p->B, p->Lk, p->ldk, 128))) return rc;
  if ((rc = make_head_tmap(&tm[2], p->v, cols, p->B, p->Lk, p->ldv, 128))) return rc;
  if ((rc = make_head_tmap(&tm[3], p->d_o, cols, p->B, p->Lq, p->lddo, 128))) return rc;
  if ((rc = make_head_tmap(&tm[4], p->dq, cols, p->B, p->Lq, p->lddq, 128))) return rc;
  if ((rc = make_head_tmap(&tm[5], p->dk, cols, p->B, p->Lk, p->lddk, 128))) return rc;
  if ((rc = make_head_tmap(&tm[6], p->dv, cols, p->B, p->Lk, p->lddv, 128))) return rc;
  AttnBwdV2KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.B = p->B; kp.H = p->H; kp.Lq = p->Lq; kp.Lk = p->Lk;
  kp.n_qt = (p->Lq + 127) / 128; kp.n_kt = (p->Lk + 127) / 128;
  kp.total_items = kp.n_kt * kp.H * kp.B;
  kp.sched = sched_slot(st);
  kp.scale = 1.0f / sqrtf((float)p->head_dim);
  kp.scale_log2 = kp.scale * 1.4426950408889634f;
  kp.lse = p->lse; kp.delta = p->delta;
  // OT_ATTN_BWD_NCG (read once): element-wise warps per TMEM lane quarter, 2 or 4 (default 2: measured faster, profiles/README.md)
  static const int ncg = [] { const char* e = getenv("OT_ATTN_BWD_NCG"); return (e && atoi(e) == 4) ? 4 : 2; }();
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_bwd_v2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM_BYTES));
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_attn_bwd_v2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM_BYTES));
    attr_done = true;
  }
  if ((rc = attn_bwd_prologue(p, st))) return rc;
  const int sms = num_sms();
  const int grid = kp.total_items < sms ? kp.total_items : sms;
  if (ncg == 2) ot_attn_bwd_v2_kernel<2><<<grid, B2Cfg<2>::THREADS, B2_SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], kp);
  else ot_attn_bwd_v2_kernel<4><<<grid, B2Cfg<4>::THREADS, B2_SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
