// ot_wgrad.cu — weight-gradient GEMM for sm_100a:   C[g][m, n] += sum_rows P[row, m] * Q[row, n]
//
// Replaces tape.gradient (OT/train.py:131) for the Dense kernels of the OneTrans path: the shared
// S-token weights contract over up to ~10^6 rows, each NS-token weight set over the B rows of its token
// (OT/model.py:43-54, 142-147).  Both operands are read exactly as the forward pass left them in HBM
// (row-major activations); the contraction runs over rows, so both are MN-major UMMA operands:
// a TMA box of [64 rows x 64 columns] (128-byte swizzle) is one MN-major slab, no transposes anywhere.
//
// Work split: one CTA per (output group, MT x BN tile, row slice), MT = 128 or 256 output rows (two TMEM accumulators).  The
// kernel is bound by L2 -> SM traffic, not by HBM (every CTA of a row slice re-reads the P / Q tiles of the other output tiles:
// 48 KB per 64 rows and 128 x 256 tile against a 42 B/clk/SM L2 port), so the 256-row tile moves a third fewer bytes per FLOP.
// One CTA per (output group, MT x BN tile, row slice).  The row range of an output is cut
// into slices so that the launch fills the device about twice; every CTA keeps its accumulator in
// TMEM for its whole slice and flushes once with fp32 atomics (the caller zeroes C).
//   warp 0 lane*: TMA producer     warp 1 lane*: MMA issuer     all 4 warps: epilogue
//   warps 2-3 (CTAs of the first M tile, when asked): column sums of the Q tiles as they pass through shared
//   memory - the bias gradient of the same Dense layer without a second pass over Q
//   warps 4-11 (PGELU instantiation only): apply GELU to the P tile in shared memory before the MMA reads it.  The fused FFN
//   forward (ot_ffn_fused.cu) keeps h = gelu(pre) on chip and stores only the pre-activation; dW2 = h^T dy (OT/train.py:131 for
//   the second Dense of OT/model.py:138,145) then takes P = pre and rebuilds h tile by tile instead of reading a stored copy
#include <stdlib.h>
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int WG_THREADS = 128;
static constexpr int WG_XFORM_THREADS = 256;   // warps 4-11 of the PGELU instantiation (two per scheduler: one warp could not keep up)
static constexpr int WG_KROWS = 64;  // rows (contraction) per pipeline stage

struct WgradSegDev {
  int n_units, rows_per_unit, group_start, group_stride;
  int kb_per_unit;     // ceil(rows_per_unit / 64)
  int n_out;           // outputs of this segment
  int kb_per_out;      // k-blocks contracted per output
  int slices_per_out, kb_per_slice;
  int item_start;      // first work item of this segment
};

struct WgradKParams {
  int Mdim, Ndim, m_tiles, n_tiles;
  int n_segs;
  WgradSegDev segs[2];
  float* C;
  long long c_group_stride, c_stride_m, c_stride_n;
  float* q_colsum;
  long long q_colsum_group_stride;
  int vec_flush;   // C rows are contiguous along n and 16-byte aligned: flush with 4-wide vector reductions
};

// gelu of eight bf16 values (one 16-byte chunk), same arithmetic as the forward epilogue (ot_common.cuh gelu_erf2)
__device__ __forceinline__ uint4 gelu_chunk(uint4 q) {
  uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    upk2(gelu_erf2(pk2(bf16lo(w[i]), bf16hi(w[i]))), a, b);
    w[i] = pack_bf16x2(a, b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int BN, int SWB, int MT = 128>
struct WgradCfg {
  static constexpr int SLAB_COLS = SWB / 2;                  // columns per MN-major slab
  static constexpr int SLAB_BYTES = WG_KROWS * SWB;          // 64 rows x SWB bytes
  static constexpr int P_SLABS = MT / SLAB_COLS;
  static constexpr int Q_SLABS = BN / SLAB_COLS;
  static constexpr int P_BYTES = P_SLABS * SLAB_BYTES;       // 16 KB
  static constexpr int Q_BYTES = Q_SLABS * SLAB_BYTES;
  static constexpr int STAGE_BYTES = P_BYTES + Q_BYTES;
  static constexpr int STAGES_RAW = (227 * 1024 - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;   // barriers: (3 * STAGES + 1) * 8 + 4 bytes <= 256
  static constexpr int ACC_COLS = (MT / 128) * BN;           // one [128 x BN] fp32 accumulator per 128 output rows
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
};

template <int BN, int SWB, bool PGELU, int MT = 128>
__global__ void __launch_bounds__(WG_THREADS + (PGELU ? WG_XFORM_THREADS : 0), 1)
ot_wgrad_kernel(const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmQ0,
                const __grid_constant__ CUtensorMap tmP1, const __grid_constant__ CUtensorMap tmQ1,
                const __grid_constant__ WgradKParams p) {
  using Cfg = WgradCfg<BN, SWB, MT>;
  static_assert(!PGELU || MT == 128, "the GELU transform is sized for a 128-column P tile");
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* done_bar = bars + 2 * STAGES;
  uint64_t* xform_bar = bars + 2 * STAGES + 1;    // [STAGES] P tile transformed (PGELU only; 4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- decode the work item ----
  int item = blockIdx.x;
  int s = (p.n_segs > 1 && item >= p.segs[1].item_start) ? 1 : 0;
  const WgradSegDev& sg = p.segs[s];
  item -= sg.item_start;
  const int tiles = p.m_tiles * p.n_tiles;
  // order: slice-major so that the CTAs running together read different rows of the same tiles
  const int tile = item % tiles;
  const int rest = item / tiles;
  const int slice = rest % sg.slices_per_out;
  const int out_idx = rest / sg.slices_per_out;
  const int mt = tile / p.n_tiles;
  const int nt = tile - mt * p.n_tiles;
  const int kb_begin = slice * sg.kb_per_slice;
  const int kb_end = min(sg.kb_per_out, kb_begin + sg.kb_per_slice);
  const int num_kb = kb_end - kb_begin;
  const int group = sg.group_start + out_idx * sg.group_stride;
  const CUtensorMap* tmP = s == 0 ? &tmP0 : &tmP1;
  const CUtensorMap* tmQ = s == 0 ? &tmQ0 : &tmQ1;
  // the column sums of a 64-row k-block are shared out over the CTAs of up to 8 M tiles (they all stream the same Q tile)
  int m_split = 1;
  while (m_split * 2 <= p.m_tiles && m_split < 8) m_split *= 2;
  const bool do_colsum = (p.q_colsum != nullptr) && (mt < m_split);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(tmP);
    tma_prefetch_desc(tmQ);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], do_colsum ? 3 : 1);   // tcgen05.commit (+ the two column-sum warps)
    }
    mbar_init(done_bar, 1);
    if (PGELU) for (int i = 0; i < STAGES; ++i) mbar_init(&xform_bar[i], WG_XFORM_THREADS / 32);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          int unit, r0;
          if (sg.group_stride != 0) { unit = out_idx; r0 = kb * WG_KROWS; }
          else { unit = kb / sg.kb_per_unit; r0 = (kb - unit * sg.kb_per_unit) * WG_KROWS; }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sp = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sq = sp + Cfg::P_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < Cfg::P_SLABS; ++i)
            tma_load_3d(sp + i * Cfg::SLAB_BYTES, tmP, &full_bar[stage], mt * MT + i * Cfg::SLAB_COLS, r0, unit);
#pragma unroll
          for (int i = 0; i < Cfg::Q_SLABS; ++i)
            tma_load_3d(sq + i * Cfg::SLAB_BYTES, tmQ, &full_bar[stage], nt * BN + i * Cfg::SLAB_COLS, r0, unit);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);  // both operands MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(PGELU ? &xform_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sp = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sq = sp + Cfg::P_BYTES;
          // LBO = distance between slabs along M/N; each UMMA consumes 16 rows = 16*SWB bytes
          const uint64_t adesc = make_smem_desc<SWB>(sp, Cfg::SLAB_BYTES);
          const uint64_t bdesc = make_smem_desc<SWB>(sq, Cfg::SLAB_BYTES);
#pragma unroll
          for (int k = 0; k < WG_KROWS / 16; ++k) {
            const uint64_t adv = static_cast<uint64_t>((16 * SWB) >> 4) * k;
            umma_bf16_ss(tmem_base, adesc + adv, bdesc + adv, idesc, (i | k) != 0 ? 1u : 0u);
          }
          if (MT == 256) {   // output rows 128..255: the next 128 P columns (slabs 128 / SLAB_COLS onwards), second accumulator
            const uint64_t adesc2 = make_smem_desc<SWB>(sp + (128 / Cfg::SLAB_COLS) * Cfg::SLAB_BYTES, Cfg::SLAB_BYTES);
#pragma unroll
            for (int k = 0; k < WG_KROWS / 16; ++k) {
              const uint64_t adv = static_cast<uint64_t>((16 * SWB) >> 4) * k;
              umma_bf16_ss(tmem_base + BN, adesc2 + adv, bdesc + adv, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(done_bar);
      }
    } else if (PGELU && warp >= 4) {
      // warps 4-11: GELU of the P tile (64 rows x 128 columns = 1024 chunks of 16 bytes, 4 per thread), in place, then hand the
      // stage to the MMA warp.  Rows past the end of a unit arrive zero-filled and stay zero (gelu(0) == 0).
      const int t = threadIdx.x - WG_THREADS;
      constexpr int CPR = 128 / 8;                   // chunks per tile row (all slabs)
      constexpr int CPS = Cfg::SLAB_COLS / 8;        // chunks per slab row
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* sp = smem + stage * Cfg::STAGE_BYTES;
        // all loads, then the arithmetic, then all stores: an in-place loop would serialise on possible aliasing (LDS after STS)
        constexpr int NCH = (WG_KROWS * CPR) / WG_XFORM_THREADS;      // 4 chunks per thread
        uint4 q[NCH];
        uint4* ptr[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int idx = j * WG_XFORM_THREADS + t;
          const int r = idx / CPR, ch = idx % CPR;
          ptr[j] = reinterpret_cast<uint4*>(sp + (ch / CPS) * Cfg::SLAB_BYTES + swz_off<SWB>(r, ch % CPS));
          q[j] = *ptr[j];
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) q[j] = gelu_chunk(q[j]);
#pragma unroll
        for (int j = 0; j < NCH; ++j) *ptr[j] = q[j];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xform_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    } else if (do_colsum && warp < 4) {
      // warps 2-3: thread t owns one 16-byte chunk (8 columns) of the Q tile and a group of its 64 rows
      constexpr int NCHK = BN / 8;                    // 16-byte chunks per tile row (8, 16 or 32)
      constexpr int RG = 64 / NCHK;                   // row groups: 64 threads = NCHK chunks x RG groups
      constexpr int CPS = Cfg::SLAB_COLS / 8;         // chunks per slab row
      const int t = threadIdx.x - 64;
      const int chk = t % NCHK, rg = t / NCHK;
      const int slab = chk / CPS, cin = chk % CPS;
      const int n_r = (WG_KROWS / m_split) / RG;      // rows per thread and k-block (>= 1: m_split <= 8 <= 64 / RG)
      const int r_begin = mt * (WG_KROWS / m_split) + rg * n_r;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* sq = smem + stage * Cfg::STAGE_BYTES + Cfg::P_BYTES + slab * Cfg::SLAB_BYTES;
#pragma unroll 4
        for (int r = r_begin; r < r_begin + n_r; ++r) {
          const uint4 q = *reinterpret_cast<const uint4*>(sq + swz_off<SWB>(r, cin));
          acc[0] += bf16lo(q.x); acc[1] += bf16hi(q.x); acc[2] += bf16lo(q.y); acc[3] += bf16hi(q.y);
          acc[4] += bf16lo(q.z); acc[5] += bf16hi(q.z); acc[6] += bf16lo(q.w); acc[7] += bf16hi(q.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      float* dst = p.q_colsum + (long long)group * p.q_colsum_group_stride + nt * BN + chk * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(dst + e, acc[e]);
    }
    // ---- epilogue: warps 0-3 flush their 32 accumulator rows each ----
    if (warp < 4) {
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int mh = 0; mh < MT / 128; ++mh) {
    const int m = mt * MT + mh * 128 + warp * 32 + lane;
    float* crow = p.C + (long long)group * p.c_group_stride + (long long)m * p.c_stride_m +
                  (long long)(nt * BN) * p.c_stride_n;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + mh * BN + c * 32, v);
      tmem_ld_wait();
      if (p.vec_flush) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + c * 32 + 4 * j), "f"(__uint_as_float(v[4 * j])),
                       "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(crow + (long long)(c * 32 + j) * p.c_stride_n, __uint_as_float(v[j]));
      }
    }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int SWB, bool PGELU = false, int MT = 128>
static int launch_wgrad(const CUtensorMap* tm, const WgradKParams& kp, int items, cudaStream_t st) {
  using Cfg = WgradCfg<BN, SWB, MT>;
  static bool attr_done = false;
  auto kern = ot_wgrad_kernel<BN, SWB, PGELU, MT>;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  kern<<<items, WG_THREADS + (PGELU ? WG_XFORM_THREADS : 0), Cfg::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int wgrad_impl(const ot_wgrad_params* p, cudaStream_t st) {
  if (!p || !p->C) OT_FAIL(OT_ERR_INVALID_ARG, "ot_wgrad: null pointer");
  if (p->p_row_scale) OT_FAIL(OT_ERR_INVALID_ARG, "ot_wgrad: p_row_scale is reserved");
  const int swb = p->swizzle == 0 ? 128 : p->swizzle;
  if (swb != 128 && swb != 64) OT_FAIL(OT_ERR_INVALID_ARG, "ot_wgrad: swizzle must be 0, 64 or 128");
  if (p->Mdim <= 0 || p->Mdim % 128) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_wgrad: Mdim=%d not a multiple of 128", p->Mdim);
  if (p->Ndim <= 0 || p->Ndim % 64) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_wgrad: Ndim=%d not a multiple of 64", p->Ndim);
  if (p->n_segs < 1 || p->n_segs > 2) OT_FAIL(OT_ERR_INVALID_ARG, "ot_wgrad: n_segs=%d", p->n_segs);
  int bn = p->block_n;
  if (bn == 0) bn = (p->Ndim % 256 == 0) ? 256 : (p->Ndim % 128 == 0) ? 128 : 64;
  if ((bn != 64 && bn != 128 && bn != 256) || p->Ndim % bn) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_wgrad: block_n=%d", bn);

  WgradKParams kp;
  memset(&kp, 0, sizeof(kp));
  // 256-row output tiles (two accumulators) where the shape allows: a third less L2 -> SM traffic per FLOP (see the file header);
  // measured on B200: no gain (dW1 1.83 vs 1.88 ms, the 256 x 256 gradients slower), so the 128-row tile stays the default and
  // OT_WGRAD_MT=256 selects the wide tile for A/B runs
  static const int mt_env = [] { const char* e = getenv("OT_WGRAD_MT"); return e ? atoi(e) : 128; }();
  const int mt = (mt_env == 256 && !p->p_gelu && swb == 128 && bn == 256 && p->Mdim % 256 == 0) ? 256 : 128;
  kp.Mdim = p->Mdim; kp.Ndim = p->Ndim; kp.m_tiles = p->Mdim / mt; kp.n_tiles = p->Ndim / bn;
  kp.n_segs = p->n_segs; kp.C = p->C;
  kp.c_group_stride = p->c_group_stride; kp.c_stride_m = p->c_stride_m; kp.c_stride_n = p->c_stride_n;
  kp.q_colsum = p->q_colsum; kp.q_colsum_group_stride = p->q_colsum_group_stride;
  kp.vec_flush = (p->c_stride_n == 1 && (p->c_stride_m % 4) == 0 && (p->c_group_stride % 4) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->C) & 15) == 0) ? 1 : 0;
  const int tiles = kp.m_tiles * kp.n_tiles;
  // default: two CTAs' worth of row slices per SM.  OT_WGRAD_WAVES (read once) is an experiment switch for round 2: more, shorter
  // slices trade accumulator flushes (fp32 reductions) for a shorter tail; unset = the verified behaviour
  static const int waves = [] { const char* e = getenv("OT_WGRAD_WAVES"); const int w = e ? atoi(e) : 0; return (w >= 1 && w <= 16) ? w : 2; }();
  const int target = p->target_ctas > 0 ? p->target_ctas : waves * num_sms();

  // total k-blocks over all outputs, to size the slices evenly
  long long total_kb = 0;
  for (int s = 0; s < p->n_segs; ++s) {
    const ot_wgrad_seg& sg = p->segs[s];
    if (!sg.P || !sg.Q || sg.n_units <= 0 || sg.rows_per_unit <= 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_wgrad: bad segment %d", s);
    total_kb += (long long)sg.n_units * ((sg.rows_per_unit + WG_KROWS - 1) / WG_KROWS);
  }
  long long kb_target = (total_kb * tiles + target - 1) / target;  // k-blocks per CTA
  if (kb_target < 8) kb_target = 8;

  CUtensorMap tm[4];
  int items = 0;
  for (int s = 0; s < p->n_segs; ++s) {
    const ot_wgrad_seg& sg = p->segs[s];
    WgradSegDev& d = kp.segs[s];
    d.n_units = sg.n_units; d.rows_per_unit = sg.rows_per_unit; d.group_start = sg.group_start; d.group_stride = sg.group_stride;
    d.kb_per_unit = (sg.rows_per_unit + WG_KROWS - 1) / WG_KROWS;
    if (sg.group_stride != 0) { d.n_out = sg.n_units; d.kb_per_out = d.kb_per_unit; }
    else { d.n_out = 1; d.kb_per_out = d.kb_per_unit * sg.n_units; }
    d.slices_per_out = (int)((d.kb_per_out + kb_target - 1) / kb_target);
    if (d.slices_per_out < 1) d.slices_per_out = 1;
    d.kb_per_slice = (d.kb_per_out + d.slices_per_out - 1) / d.slices_per_out;
    d.slices_per_out = (d.kb_per_out + d.kb_per_slice - 1) / d.kb_per_slice;
    d.item_start = items;
    items += d.n_out * d.slices_per_out * tiles;
    // rank-3 maps (col, row_in_unit, unit); rows past rows_per_unit are zero-filled by TMA
    const int slab_cols = swb / 2;
    {
      uint64_t dims[3] = {(uint64_t)p->Mdim, (uint64_t)sg.rows_per_unit, (uint64_t)sg.n_units};
      uint64_t str[2] = {(uint64_t)sg.p_stride_row * 2, (uint64_t)(sg.n_units > 1 ? sg.p_stride_unit : sg.p_stride_row * sg.rows_per_unit) * 2};
      uint32_t box[3] = {(uint32_t)slab_cols, (uint32_t)WG_KROWS, 1};
      int rc = make_tmap_bf16(&tm[2 * s], sg.P, 3, dims, str, box, swb);
      if (rc) return rc;
    }
    {
      uint64_t dims[3] = {(uint64_t)p->Ndim, (uint64_t)sg.rows_per_unit, (uint64_t)sg.n_units};
      uint64_t str[2] = {(uint64_t)sg.q_stride_row * 2, (uint64_t)(sg.n_units > 1 ? sg.q_stride_unit : sg.q_stride_row * sg.rows_per_unit) * 2};
      uint32_t box[3] = {(uint32_t)slab_cols, (uint32_t)WG_KROWS, 1};
      int rc = make_tmap_bf16(&tm[2 * s + 1], sg.Q, 3, dims, str, box, swb);
      if (rc) return rc;
    }
  }
  if (p->n_segs == 1) { tm[2] = tm[0]; tm[3] = tm[1]; }

  if (p->p_gelu) {
    if (swb != 128 || bn != 256) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_wgrad: p_gelu is built for block_n 256 with the 128-byte swizzle (Ndim=%d)", p->Ndim);
    return launch_wgrad<256, 128, true>(tm, kp, items, st);
  }
  if (mt == 256) return launch_wgrad<256, 128, false, 256>(tm, kp, items, st);
  if (swb == 128) {
    if (bn == 256) return launch_wgrad<256, 128>(tm, kp, items, st);
    if (bn == 128) return launch_wgrad<128, 128>(tm, kp, items, st);
    return launch_wgrad<64, 128>(tm, kp, items, st);
  } else {
    if (bn == 256) return launch_wgrad<256, 64>(tm, kp, items, st);
    if (bn == 128) return launch_wgrad<128, 64>(tm, kp, items, st);
    return launch_wgrad<64, 64>(tm, kp, items, st);
  }
}

}  // namespace ot
