// ot_gemm.cu — mixed-parameter grouped GEMM for sm_100a.
//
//   out[row, n] = epilogue( sum_k A[row, k] * W[group(row)][n, k] )
//
// Replaces the per-position Dense loops of the reference (OT/model.py:84-92 QKV, :117 Wo,
// :154-163 FFN, :262-265 sequence projections) and their input gradients.  The rule "which rows use
// which weight set" (OT/model.py:67-74) arrives as data (ot_gemm_seg), so the shared S-token run and
// the per-token NS runs are tiles of ONE persistent launch.
//
// Kernel structure (one CTA per SM, persistent; 20 warps, registers re-balanced with setmaxnreg: 32 per control
// thread, 112 per epilogue thread).  Tiles are scheduled dynamically: every CTA starts with tile blockIdx.x and then
// draws from a device-wide counter, so CTAs that start late (an SM held by a collective) take fewer tiles.
//   warp 0   : TMA producer   — tile index -> ring (to warps 1, 4-19); A tile [128 x BK] and W tile [BN x BK] into a
//                               STAGES-deep smem ring
//   warp 1   : MMA issuer     — tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulators
//   warps 4-19: epilogue      — tcgen05.ld -> packed fp32 math (row scale, bias, GELU, GELU', residual, dropout)
//                               -> bf16 -> swizzled smem staging tile -> TMA store; residual / GELU' inputs arrive
//                               by TMA one tile ahead (four sets of four warps, one 64-column chunk each; with
//                               BN < 256 the sets split into groups that take tiles in turn); OT_EPI_NORM adds the
//                               RMSNorm of the finished rows as a second output (row statistics exchanged in smem)
// The two TMEM accumulators (2*BN <= 512 columns) let the epilogue of tile i overlap the MMAs of
// tile i+1.  Roofline: at d=256 every GEMM of the block is HBM-bound, so the
// epilogue reads/writes each activation byte exactly once and in full 128-byte lines.
#include "ot_gemm_common.cuh"
#include "ot_host.h"

namespace ot {

static constexpr int EPI_SETS = 4;
static constexpr int EPI_SET_THREADS = 128;
static constexpr int EPI_THREADS = EPI_SETS * EPI_SET_THREADS;
static constexpr int CTRL_THREADS = 128;      // one control warpgroup: warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle
static constexpr int GEMM_THREADS = CTRL_THREADS + EPI_THREADS;   // + 16 epilogue warps (warps 4..19)
static constexpr int EPI_REGS = 112, CTRL_REGS = 32;   // setmaxnreg: the control warpgroup releases 4*32*(96-32) registers into the CTA pool, exactly what the four epilogue warpgroups take (4*128*(112-96)); asking for more than was released spins forever
static constexpr int EPI_BAR_ID = 1;         // named barriers 1..4: one per epilogue set
static constexpr int EPI_BAR_NORM = 5;       // 5..8: the sets working on one tile (row statistics of the fused RMSNorm)
static constexpr int CHUNK = 64;             // epilogue column chunk (128 bytes of bf16 per row)
static constexpr int CH_BYTES = 128 * 64 * 2;   // one staged chunk: 128 rows x 128 bytes
#ifndef OT_DUAL_LATE_WAIT
#define OT_DUAL_LATE_WAIT 0     // experiment switch, see the dual-output pass of the epilogue; 0 = the verified build
#endif
static constexpr int TILE_SLOTS = 8;         // tile-index ring: producer -> MMA issuer and epilogue warps

struct GemmKParams {
  int N, K;
  int n_segs;
  int total_mblks, n_nblks;
  uint32_t nn_rcp;         // ceil(2^32 / n_nblks)
  int a_transposed;
  int flags;
  GemmSegDev segs[3];
  __nv_bfloat16* out;  long long ldo;
  __nv_bfloat16* out2; long long ldo2;
  const __nv_bfloat16* res; long long ldr;
  const __nv_bfloat16* aux; long long ldaux;
  const float* bias; long long bias_group_stride;
  const float* row_scale;
  const float* res_hp; float* out_hp; long long ld_hp; long long hp_row0;
  uint32_t drop_seed, drop_thr16; float drop_scale;
  __nv_bfloat16* norm_out; long long ld_norm; const float* norm_gain; float* norm_rstd; float norm_eps;
  int* sched;              // dynamic tile counter (zeroed by the launcher) or NULL = static round-robin
};

template <int BN, int SWB>
struct GemmCfg {
  static constexpr int BK = SWB / 2;                 // bf16 elements per swizzle row
  static constexpr int A_BYTES = BM * SWB;
  static constexpr int B_BYTES = BN * SWB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = EPI_SETS * BM * CHUNK * 2;  // one 128x64 bf16 buffer per epilogue set
  static constexpr int NORM_BYTES = 2 * EPI_SETS * BM * 4;        // fused RMSNorm: per-set partial sums of squares, double-buffered
  static constexpr int BUDGET = 227 * 1024 - STAGING_BYTES - NORM_BYTES - 512;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + NORM_BYTES + 512;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int SWB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
ot_mixed_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                     const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN, SWB>;
  constexpr int BK = Cfg::BK;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];

  uint8_t* stage_base = smem;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  float* ss_part = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES);          // [2][EPI_SETS][BM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES + Cfg::NORM_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* in_bars = bars + 2 * STAGES + 4;     // [EPI_SETS] auxiliary-input tiles
  uint64_t* tile_full = bars + 2 * STAGES + 4 + EPI_SETS;                 // [TILE_SLOTS] tile index published
  uint64_t* tile_empty = tile_full + TILE_SLOTS;                          // [TILE_SLOTS] read by the MMA issuer + 16 epilogue warps
  int* tile_ring = reinterpret_cast<int*>(tile_empty + TILE_SLOTS);       // [TILE_SLOTS] tile index or -1 (no more work)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ring + TILE_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.total_mblks * p.n_nblks;
  const int num_kb = p.K / BK;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * EPI_SETS);
    }
    for (int i = 0; i < EPI_SETS; ++i) mbar_init(&in_bars[i], 1);
    for (int i = 0; i < TILE_SLOTS; ++i) { mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1 + 4 * EPI_SETS); }
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

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // Tile schedule: every CTA starts with tile blockIdx.x; afterwards tiles come from a device-wide counter
      // (p.sched) so that CTAs which start late - e.g. because a collective holds their SM - simply take fewer tiles
      // instead of becoming stragglers.  The index travels to the MMA issuer and the epilogue warps through a small
      // ring; -1 ends the kernel.  The next index is fetched while this tile's loads are issued.
      int stage = 0;
      uint32_t phase = 0;
      int n = 0;
      int tile = blockIdx.x;
      while (true) {
        const int slot = n & (TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_empty[slot], ((n / TILE_SLOTS) & 1) ^ 1);
        tile_ring[slot] = tile;
        mbar_arrive(&tile_full[slot]);
        if (tile < 0) break;
        int nxt;
        if (p.sched != nullptr) nxt = (int)gridDim.x + atomicAdd(p.sched, 1);
        else nxt = tile + (int)gridDim.x;
        if (nxt >= total_tiles) nxt = -1;
        const int mblk = div_rcp(tile, p.nn_rcp, p.n_nblks);
        const int nblk = tile - mblk * p.n_nblks;
        const TileInfo t = decode_tile(p, mblk);
        const int w_row = t.group * p.N + nblk * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(sa, &tmA, &full_bar[stage], kb * BK, t.a_c1, t.a_c2);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, w_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tile = nxt;
        ++n;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const int slot = it & (TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / TILE_SLOTS) & 1, 32);
        const int tile = tile_ring[slot];
        mbar_arrive(&tile_empty[slot]);
        if (tile < 0) break;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_backoff(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_backoff(&full_bar[stage], phase, 32);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = make_smem_desc<SWB>(sa, 16);
          const uint64_t bdesc = make_smem_desc<SWB>(sb, 16);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the 16-byte address field
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..19: four sets of four) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    // Set s owns the 64-column chunk s of every tile (BN <= 256 -> at most one chunk per set and tile), one 16 KB
    // staging tile and one mbarrier.  Inside a set, warp w may read TMEM lanes 32*(w%4)..+31 (one accumulator row
    // per thread).  Data path of a chunk:
    //   residual / GELU' input: TMA load into the staging tile, issued one TILE ahead by the set's IO thread
    //   tcgen05.ld -> packed fp32 math (FFMA2) -> bf16 -> the thread's own row of the staging tile (in place)
    //   TMA store of the staging tile; the IO thread waits for it to drain before it issues the next prefetch
    // Tiles with fewer than 128 valid rows (small batches) take the cooperative 16-byte load/store path instead.
    const int set = (warp - 4) >> 2;
    const int et = threadIdx.x - CTRL_THREADS - set * EPI_SET_THREADS;   // 0..127 inside the set
    const int lgrp = warp & 3;
    const int r_own = lgrp * 32 + lane;
    const int ld_row = et >> 3;                // cooperative copy: 16 rows per pass, 8 x 16 B per row
    const int ld_ch = et & 7;
    const uint32_t bar_id = EPI_BAR_ID + set;
    uint8_t* buf = staging + set * CH_BYTES;
    uint64_t* in_bar = &in_bars[set];
    uint32_t in_phase = 0;
    // A tile has BN/64 chunks; with BN < 256 the sets form 4/(BN/64) groups that take the CTA's tiles in turn (group g
    // owns tiles it % n_grp == g, i.e. always the same TMEM accumulator when n_grp == 2), so no set idles and every set
    // has n_grp tile periods between two of its chunks - room to prefetch its auxiliary input.
    constexpr int SPT = BN / CHUNK;                 // sets per tile
    constexpr int N_GRP = EPI_SETS / SPT;           // set groups
    const int my_grp = set / SPT;
    const bool io_thread = (et == 0);
    const bool f_bias = p.flags & OT_EPI_BIAS, f_gelu = p.flags & OT_EPI_GELU, f_res = p.flags & OT_EPI_RESIDUAL;
    const bool f_ggrad = p.flags & OT_EPI_GELU_GRAD, f_rs = p.flags & OT_EPI_ROW_SCALE, f_drop = p.flags & OT_EPI_DROPOUT;
    const bool f_norm = p.flags & OT_EPI_NORM;
    const bool has_in = f_res || f_ggrad;
    const bool dual = f_gelu && (p.out2 != nullptr);
    const __nv_bfloat16* in_ptr = f_res ? p.res : p.aux;
    const long long in_ld = f_res ? p.ldr : p.ldaux;
    // does tile t take its bf16 auxiliary input through TMA?  (tile-uniform; NS tiles with an fp32 residual do not)
    auto tma_in_tile = [&](const TileInfo& t) {
      return has_in && t.valid == BM && !(f_res && p.res_hp != nullptr && t.row0 >= p.hp_row0);
    };
    auto issue_in = [&](int tile) {            // IO thread only
      const int mblk = div_rcp(tile, p.nn_rcp, p.n_nblks);
      const TileInfo t = decode_tile(p, mblk);
      if (tma_in_tile(t)) {
        mbar_arrive_expect_tx(in_bar, CH_BYTES);
        tma_load_2d(buf, &tmIn, in_bar, (tile - mblk * p.n_nblks) * BN + (set % SPT) * CHUNK, t.row0);
      }
    };
    // tile index of ring entry n without consuming it (the entry is released when this warp reaches iteration n)
    auto peek = [&](int n) {
      const int slot = n & (TILE_SLOTS - 1);
      mbar_wait_backoff(&tile_full[slot], (n / TILE_SLOTS) & 1, 32);
      return tile_ring[slot];
    };
    // the tile `ahead` entries after entry n, or -1 when the schedule ends before it
    auto tile_after = [&](int n, int ahead) {
      int tl = -1;
      for (int j = 1; j <= ahead; ++j) {
        tl = peek(n + j);
        if (tl < 0) break;
      }
      return tl;
    };
    if (io_thread && has_in) {               // this set's first tile is ring entry my_grp
      const int first = my_grp == 0 ? (int)blockIdx.x : tile_after(0, my_grp);
      if (first >= 0) issue_in(first);
    }

    for (int it = 0;; ++it) {
      int tile;
      {
        const int slot = it & (TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / TILE_SLOTS) & 1, 32);
        tile = tile_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&tile_empty[slot]);
      }
      if (tile < 0) break;
      const int mblk = div_rcp(tile, p.nn_rcp, p.n_nblks);
      const int nblk = tile - mblk * p.n_nblks;
      const TileInfo t = decode_tile(p, mblk);
      const int n0 = nblk * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;

      float rs = 1.0f;
      if (f_rs && r_own < t.valid) rs = p.row_scale[t.row0 + r_own];
      // NS-token tiles may carry an fp32 residual stream (tile-uniform: NS units start on tile boundaries)
      const bool hp_tile = f_res && (p.res_hp != nullptr) && (t.row0 >= p.hp_row0);
      const bool tile_in = has_in && !hp_tile;
      const bool tma_in = tma_in_tile(t);
      const bool tma_out = (t.valid == BM);

      mbar_wait_backoff(&tfull_bar[acc], acc_phase, 32);
      tc_fence_after();

      if ((it % N_GRP) == my_grp) {
        const int c = set % SPT;
        const int col0 = n0 + c * CHUNK;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + acc * BN + c * CHUNK;
        const float* bias_g = f_bias ? p.bias + (long long)t.group * p.bias_group_stride + col0 : nullptr;
        auto store_rows = [&](__nv_bfloat16* dst, long long ld) {   // staging -> global, full 128-byte rows
#pragma unroll
          for (int i = 0; i < BM / 16; ++i) {
            const int r = i * 16 + ld_row;
            if (r < t.valid)
              *reinterpret_cast<uint4*>(dst + (long long)(t.row0 + r) * ld + col0 + ld_ch * 8) =
                  *reinterpret_cast<const uint4*>(buf + swz_off<128>(r, ld_ch));
          }
        };
        if (tma_in) {
          mbar_wait(in_bar, in_phase);           // prefetched one tile ago, behind the drain of the last store
          in_phase ^= 1;
        } else {
          if (io_thread) bulk_wait_read0();      // the previous TMA store has left the staging tile
          named_bar_sync(bar_id, EPI_SET_THREADS);
          if (tile_in) {   // partial tile: cooperative global -> staging
#pragma unroll
            for (int i = 0; i < BM / 16; ++i) {
              const int r = i * 16 + ld_row;
              uint4 q = make_uint4(0, 0, 0, 0);
              if (r < t.valid)
                q = *reinterpret_cast<const uint4*>(in_ptr + (long long)(t.row0 + r) * in_ld + col0 + ld_ch * 8);
              *reinterpret_cast<uint4*>(buf + swz_off<128>(r, ld_ch)) = q;
            }
            named_bar_sync(bar_id, EPI_SET_THREADS);
          }
        }
        // dual mode (FFN-1 forward keeps the pre-activation for the backward pass): pass 0 stores the pre-activation,
        // pass 1 re-reads the accumulator (TMEM reads are cheap, registers are not) and stores GELU of it
        f32x2 ss2 = pk2(0.0f);                 // OT_EPI_NORM: sum of squares of my 64 result columns (fp32, before rounding)
        const int n_pass = dual ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
#if OT_DUAL_LATE_WAIT
          // experiment switch (round 2): when nothing is read from the staging tile, wait for the drain of the pre-activation
          // store only right before pass 1 first WRITES the tile, so its TMEM load and GELU arithmetic overlap the drain
          const bool late_wait = !tile_in;
          if (pass == 1 && !late_wait) {
            if (tma_out && io_thread) bulk_wait_read0();
            named_bar_sync(bar_id, EPI_SET_THREADS);
          }
#else
          if (pass == 1) {                       // pre-activation rows have left the staging tile
            if (tma_out && io_thread) bulk_wait_read0();
            named_bar_sync(bar_id, EPI_SET_THREADS);
          }
#endif
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_x32(t_row + half * 32, v);
            tmem_ld_wait();
            f32x2 f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            if (f_rs) {
              const f32x2 rs2 = pk2(rs);
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = mul2(f[j], rs2);
            }
            if (f_bias) {                        // 32 floats shared by the whole warp: broadcast loads, L1-resident
              const float4* b4 = reinterpret_cast<const float4*>(bias_g + half * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(b4 + j);
                f[2 * j] = add2(f[2 * j], pk2(bb.x, bb.y));
                f[2 * j + 1] = add2(f[2 * j + 1], pk2(bb.z, bb.w));
              }
            }
            if (f_drop) {   // inverted dropout on the branch output, before the residual add
              const uint32_t grow = static_cast<uint32_t>(t.row0 + r_own);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t hb = dropout_bits(p.drop_seed, grow, static_cast<uint32_t>(col0 + half * 32 + 2 * j), static_cast<uint32_t>(p.N));
                float a, b;
                upk2(f[j], a, b);
                a = ((hb & 0xFFFFu) >= p.drop_thr16) ? a * p.drop_scale : 0.0f;
                b = ((hb >> 16) >= p.drop_thr16) ? b * p.drop_scale : 0.0f;
                f[j] = pk2(a, b);
              }
            }
            if (hp_tile) {   // fp32 residual in, fp32 result out: one 128-byte line per thread and half-chunk
              if (r_own < t.valid) {
                const long long hr = (long long)(t.row0 + r_own) - p.hp_row0;
                const float4* rp = reinterpret_cast<const float4*>(p.res_hp + hr * p.ld_hp + col0 + half * 32);
                float4* op = reinterpret_cast<float4*>(p.out_hp + hr * p.ld_hp + col0 + half * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 rr = rp[j];
                  f[2 * j] = add2(f[2 * j], pk2(rr.x, rr.y));
                  f[2 * j + 1] = add2(f[2 * j + 1], pk2(rr.z, rr.w));
                  float4 o;
                  upk2(f[2 * j], o.x, o.y);
                  upk2(f[2 * j + 1], o.z, o.w);
                  op[j] = o;
                }
              }
            }
            if (tile_in) {
              // the mode test stays outside the element loops: a branch per element doubled the instruction count
              uint32_t w[16];
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                const uint4 q = *reinterpret_cast<const uint4*>(buf + swz_off<128>(r_own, half * 4 + ch));
                w[ch * 4 + 0] = q.x; w[ch * 4 + 1] = q.y; w[ch * 4 + 2] = q.z; w[ch * 4 + 3] = q.w;
              }
              if (f_ggrad) {
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = mul2(f[e], gelu_erf_grad2(pk2(bf16lo(w[e]), bf16hi(w[e]))));
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = add2(f[e], pk2(bf16lo(w[e]), bf16hi(w[e])));
              }
            }
            if (f_gelu && (pass == 1 || !dual)) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = gelu_erf2(f[j]);
            }
            if (f_norm) {
#pragma unroll
              for (int j = 0; j < 16; ++j) ss2 = fma2(f[j], f[j], ss2);
            }
#if OT_DUAL_LATE_WAIT
            if (pass == 1 && late_wait && half == 0) {
              if (tma_out && io_thread) bulk_wait_read0();
              named_bar_sync(bar_id, EPI_SET_THREADS);
            }
#endif
            // each thread (over)writes only its own row of the staging tile
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              float a0, a1, a2, a3, a4, a5, a6, a7;
              upk2(f[ch * 4 + 0], a0, a1); upk2(f[ch * 4 + 1], a2, a3); upk2(f[ch * 4 + 2], a4, a5); upk2(f[ch * 4 + 3], a6, a7);
              uint4 q;
              q.x = pack_bf16x2(a0, a1); q.y = pack_bf16x2(a2, a3); q.z = pack_bf16x2(a4, a5); q.w = pack_bf16x2(a6, a7);
              *reinterpret_cast<uint4*>(buf + swz_off<128>(r_own, half * 4 + ch)) = q;
            }
          }
          if (tma_out) fence_proxy_async_smem();   // my row -> visible to the TMA store
          named_bar_sync(bar_id, EPI_SET_THREADS);
          const bool to_out2 = dual && pass == 0;
          if (tma_out) {
            if (io_thread) {
              tma_store_2d(to_out2 ? &tmOut2 : &tmOut, buf, col0, t.row0);
              bulk_commit();
            }
          } else {
            if (to_out2) store_rows(p.out2, p.ldo2); else store_rows(p.out, p.ldo);
            named_bar_sync(bar_id, EPI_SET_THREADS);   // readers done before anything lands in the tile again
          }
        }
        if (f_norm) {
          // ---- fused RMSNorm of the rows just produced (OT/model.py:19-23): the tile holds whole rows (N == BN),
          // each set its 64 columns.  The row statistics use the fp32 results of the main pass; the values to scale come
          // back from the staging tile (bf16, exactly what went to HBM) or, on fp32-residual rows, from out_hp. ----
          const long long hr = (long long)(t.row0 + r_own) - p.hp_row0;
          const bool hp_row = hp_tile && r_own < t.valid;
          auto load8 = [&](int ch, float (&x)[8]) {
            if (hp_tile) {
              if (hp_row) {
                const float4* rp = reinterpret_cast<const float4*>(p.out_hp + hr * p.ld_hp + col0 + ch * 8);
                const float4 a = rp[0], b = rp[1];
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = 0.0f;
              }
            } else {
              const uint4 q = *reinterpret_cast<const uint4*>(buf + swz_off<128>(r_own, ch));
              x[0] = bf16lo(q.x); x[1] = bf16hi(q.x); x[2] = bf16lo(q.y); x[3] = bf16hi(q.y);
              x[4] = bf16lo(q.z); x[5] = bf16hi(q.z); x[6] = bf16lo(q.w); x[7] = bf16hi(q.w);
            }
          };
          float ss, ss_hi;
          upk2(ss2, ss, ss_hi);
          ss += ss_hi;                            // accumulated in the main pass: no extra trip through the staging tile
          float* ssb = ss_part + ((it / N_GRP) & 1) * (EPI_SETS * BM);   // alternates between two consecutive tiles of a set
          ssb[set * BM + r_own] = ss;
          if (tma_out && io_thread) bulk_wait_read0();          // the `out` rows have left the staging tile
          named_bar_sync(EPI_BAR_NORM + my_grp, EPI_SET_THREADS * SPT);
          float tot = 0.0f;
#pragma unroll
          for (int s2 = 0; s2 < SPT; ++s2) tot += ssb[(my_grp * SPT + s2) * BM + r_own];
          const float rstd = rsqrtf(tot / (float)p.N + p.norm_eps);
          if (c == 0 && r_own < t.valid && p.norm_rstd != nullptr) p.norm_rstd[t.row0 + r_own] = rstd;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float x[8];
            load8(ch, x);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.norm_gain + col0 + ch * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.norm_gain + col0 + ch * 8) + 1);
            uint4 q;
            q.x = pack_bf16x2(x[0] * rstd * g0.x, x[1] * rstd * g0.y);
            q.y = pack_bf16x2(x[2] * rstd * g0.z, x[3] * rstd * g0.w);
            q.z = pack_bf16x2(x[4] * rstd * g1.x, x[5] * rstd * g1.y);
            q.w = pack_bf16x2(x[6] * rstd * g1.z, x[7] * rstd * g1.w);
            *reinterpret_cast<uint4*>(buf + swz_off<128>(r_own, ch)) = q;
          }
          if (tma_out) fence_proxy_async_smem();
          named_bar_sync(bar_id, EPI_SET_THREADS);
          if (tma_out) {
            if (io_thread) {
              tma_store_2d(&tmOut2, buf, col0, t.row0);       // tmOut2 maps norm_out in this mode
              bulk_commit();
            }
          } else {
            store_rows(p.norm_out, p.ld_norm);
            named_bar_sync(bar_id, EPI_SET_THREADS);
          }
        }
        // prefetch the auxiliary input of this set's chunk in the CTA's next tile
        if (io_thread && has_in) {
          const int nxt = tile_after(it, N_GRP);
          if (nxt >= 0) {
            bulk_wait_read0();
            issue_in(nxt);
          }
        }
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (io_thread) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// -------------------------------------------------------------------------------------------------
// host launcher
// -------------------------------------------------------------------------------------------------
template <int BN, int SWB>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmOut2,
                       const CUtensorMap& tmIn, const GemmKParams& kp, cudaStream_t st) {
  using Cfg = GemmCfg<BN, SWB>;
  static bool attr_done = false;
  auto kern = ot_mixed_gemm_kernel<BN, SWB>;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int total_tiles = kp.total_mblks * kp.n_nblks;
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmOut, tmOut2, tmIn, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int mixed_gemm_impl(const ot_gemm_params* p, cudaStream_t st) {
  if (!p || !p->A || !p->W || !p->out) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: null pointer");
  const int swb = p->swizzle == 0 ? 128 : p->swizzle;
  if (swb != 128 && swb != 64) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: swizzle must be 0, 64 or 128");
  const int bk = swb / 2;
  if (p->K <= 0 || p->K % bk != 0) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: K=%d not a multiple of %d", p->K, bk);
  if (p->N <= 0 || p->N % 64 != 0) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: N=%d not a multiple of 64", p->N);
  if (p->n_segs < 1 || p->n_segs > 3) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: n_segs=%d", p->n_segs);
  if ((p->flags & OT_EPI_BIAS) && !p->bias) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: bias flag without bias");
  if ((p->flags & OT_EPI_BIAS) && ((reinterpret_cast<uintptr_t>(p->bias) & 15) || (p->bias_group_stride % 4)))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: bias must be 16-byte aligned with a group stride that is a multiple of 4");
  if ((p->flags & OT_EPI_RESIDUAL) && !p->res) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: residual flag without res");
  if ((p->flags & OT_EPI_GELU_GRAD) && !p->aux) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: gelu-grad flag without aux");
  if ((p->flags & OT_EPI_ROW_SCALE) && !p->row_scale) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: row-scale flag without row_scale");
  if ((p->flags & OT_EPI_RESIDUAL) && (p->flags & OT_EPI_GELU_GRAD))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: residual and gelu-grad epilogues cannot be combined (one auxiliary input per launch)");
  if (p->flags & OT_EPI_NORM) {
    if (!p->norm_out || !p->norm_gain || p->out2 || (p->ld_norm % 8) || (reinterpret_cast<uintptr_t>(p->norm_gain) & 15))
      OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: OT_EPI_NORM needs norm_out, a 16-byte aligned norm_gain, ld_norm %% 8 == 0 and no out2");
    if (p->N > 256 || (p->block_n != 0 && p->block_n != p->N))
      OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: OT_EPI_NORM needs whole rows in one tile (N=%d <= 256, block_n == N)", p->N);
  }
  if ((p->ldo % 8) || (p->out2 && (p->ldo2 % 8)) || (p->res && (p->ldr % 8)) || (p->aux && (p->ldaux % 8)))
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: leading dimensions must be multiples of 8 elements");

  int bn = p->block_n;
  if (bn == 0 && (p->flags & OT_EPI_NORM)) bn = p->N;
  if (bn == 0) bn = (p->N % 256 == 0) ? 256 : (p->N % 128 == 0) ? 128 : 64;
  if ((bn != 64 && bn != 128 && bn != 256) || p->N % bn != 0)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: block_n=%d does not tile N=%d", bn, p->N);

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.N = p->N; kp.K = p->K; kp.n_segs = p->n_segs; kp.a_transposed = p->a_transposed; kp.flags = p->flags;
  int mb = 0;
  for (int s = 0; s < p->n_segs; ++s) {
    const ot_gemm_seg& sg = p->segs[s];
    if (sg.n_units <= 0 || sg.rows_per_unit <= 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: empty segment %d", s);
    GemmSegDev& d = kp.segs[s];
    d.row_start = sg.row_start; d.n_units = sg.n_units; d.rows_per_unit = sg.rows_per_unit;
    d.group_start = sg.group_start; d.group_stride = sg.group_stride; d.a_row_start = sg.a_row_start;
    d.mblk_start = mb; d.mblk_per_unit = (sg.rows_per_unit + BM - 1) / BM;
    d.mpu_rcp = (uint32_t)(((1ull << 32) + d.mblk_per_unit - 1) / d.mblk_per_unit);
    mb += d.mblk_per_unit * sg.n_units;
    const int last_group = sg.group_start + (sg.n_units - 1) * sg.group_stride;
    if (sg.group_start < 0 || last_group >= p->n_groups)
      OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: segment %d addresses weight group %d of %d", s, last_group, p->n_groups);
  }
  kp.total_mblks = mb; kp.n_nblks = p->N / bn;
  kp.nn_rcp = (uint32_t)(((1ull << 32) + kp.n_nblks - 1) / kp.n_nblks);
  if ((long long)mb * kp.n_nblks >= (1ll << 24)) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_mixed_gemm: %lld tiles exceed the tile-index range", (long long)mb * kp.n_nblks);
  kp.out = (__nv_bfloat16*)p->out; kp.ldo = p->ldo; kp.out2 = (__nv_bfloat16*)p->out2; kp.ldo2 = p->ldo2;
  kp.res = (const __nv_bfloat16*)p->res; kp.ldr = p->ldr; kp.aux = (const __nv_bfloat16*)p->aux; kp.ldaux = p->ldaux;
  kp.bias = p->bias; kp.bias_group_stride = p->bias_group_stride; kp.row_scale = p->row_scale;
  kp.res_hp = p->res_hp; kp.out_hp = p->out_hp; kp.ld_hp = p->ld_hp; kp.hp_row0 = p->hp_row0;
  kp.norm_out = (__nv_bfloat16*)p->norm_out; kp.ld_norm = p->ld_norm; kp.norm_gain = p->norm_gain; kp.norm_rstd = p->norm_rstd;
  kp.norm_eps = p->norm_eps;
  kp.sched = sched_slot(st);
  if (p->flags & OT_EPI_DROPOUT) {
    if (!(p->drop_rate >= 0.0f && p->drop_rate < 1.0f)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: drop_rate=%f", (double)p->drop_rate);
    kp.drop_seed = p->drop_seed; kp.drop_thr16 = (uint32_t)(p->drop_rate * 65536.0f + 0.5f); kp.drop_scale = 1.0f / (1.0f - p->drop_rate);
  }
  if (p->res_hp || p->out_hp) {
    if (!(p->flags & OT_EPI_RESIDUAL) || !p->res_hp || !p->out_hp || (p->ld_hp % 4))
      OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: res_hp/out_hp need OT_EPI_RESIDUAL, both pointers and ld_hp %% 4 == 0");
    bool ok = false;
    for (int s2 = 0; s2 < p->n_segs; ++s2) ok = ok || (p->segs[s2].row_start == p->hp_row0);
    if (!ok) OT_FAIL(OT_ERR_INVALID_ARG, "ot_mixed_gemm: hp_row0=%lld is not the first row of a segment", (long long)p->hp_row0);
  }

  // tensor maps: A always rank 3 (k, c1, c2); W rank 2 (k, group*N + n)
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[3]; uint64_t str[2]; uint32_t box[3];
    dims[0] = (uint64_t)p->K; dims[1] = (uint64_t)p->a_dim1; dims[2] = (uint64_t)(p->a_dim2 > 0 ? p->a_dim2 : 1);
    str[0] = (uint64_t)p->a_stride1 * 2; str[1] = (uint64_t)(p->a_dim2 > 1 ? p->a_stride2 : p->a_stride1 * p->a_dim1) * 2;
    box[0] = bk;
    if (!p->a_transposed) { box[1] = BM; box[2] = 1; } else { box[1] = 1; box[2] = BM; }
    int rc = make_tmap_bf16(&tmA, p->A, 3, dims, str, box, swb);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->n_groups * (uint64_t)p->N};
    uint64_t str[1] = {(uint64_t)p->ldw * 2};
    uint32_t box[2] = {(uint32_t)bk, (uint32_t)bn};
    int rc = make_tmap_bf16(&tmB, p->W, 2, dims, str, box, swb);
    if (rc) return rc;
  }
  // epilogue I/O maps: [rows, N] bf16, box 128 rows x 64 columns in the 128-byte swizzle of the staging tiles
  CUtensorMap tmOut, tmOut2, tmIn;
  {
    long long row_extent = 0;
    for (int s = 0; s < p->n_segs; ++s) {
      const long long e = (long long)p->segs[s].row_start + (long long)p->segs[s].n_units * p->segs[s].rows_per_unit;
      if (e > row_extent) row_extent = e;
    }
    uint64_t dims[2] = {(uint64_t)p->N, (uint64_t)row_extent};
    uint32_t box[2] = {(uint32_t)CHUNK, (uint32_t)BM};
    uint64_t str[1] = {(uint64_t)p->ldo * 2};
    int rc = make_tmap_bf16(&tmOut, p->out, 2, dims, str, box, 128);
    if (rc) return rc;
    tmOut2 = tmOut; tmIn = tmOut;
    if (p->out2) {
      str[0] = (uint64_t)p->ldo2 * 2;
      rc = make_tmap_bf16(&tmOut2, p->out2, 2, dims, str, box, 128);
      if (rc) return rc;
    }
    if (p->flags & OT_EPI_NORM) {
      str[0] = (uint64_t)p->ld_norm * 2;
      rc = make_tmap_bf16(&tmOut2, p->norm_out, 2, dims, str, box, 128);
      if (rc) return rc;
    }
    const void* in_base = (p->flags & OT_EPI_RESIDUAL) ? p->res : (p->flags & OT_EPI_GELU_GRAD) ? p->aux : nullptr;
    if (in_base) {
      str[0] = (uint64_t)((p->flags & OT_EPI_RESIDUAL) ? p->ldr : p->ldaux) * 2;
      rc = make_tmap_bf16(&tmIn, in_base, 2, dims, str, box, 128);
      if (rc) return rc;
    }
  }
  if (swb == 128) {
    if (bn == 256) return launch_gemm<256, 128>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
    if (bn == 128) return launch_gemm<128, 128>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
    return launch_gemm<64, 128>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
  } else {
    if (bn == 256) return launch_gemm<256, 64>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
    if (bn == 128) return launch_gemm<128, 64>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
    return launch_gemm<64, 64>(tmA, tmB, tmOut, tmOut2, tmIn, kp, st);
  }
}

}  // namespace ot
