// ot_ffn_fused.cu — the whole MixedFFN (OT/model.py:149-163) plus the block's second residual and the next RMSNorm
// (OT/model.py:196-198, :191 of the next block) as ONE persistent sm_100a kernel:
//
//   y[row, :] = res[row, :] + drop( gelu(zn[row, :] W1[g] + b1[g]) W2[g] + b2[g] ),   g = group(row)   (OT/model.py:67-74 as data)
//   norm_out[row, :] = y[row, :] * rsqrt(mean(y^2) + eps) * gain
//
// The hidden activation h = gelu(.) [rows, F] never reaches HBM: per 128-row tile the kernel walks F in chunks of 128 columns,
// FFN-1 of chunk c accumulates into one of two TMEM buffers, the epilogue warps add the bias, apply GELU and leave the bf16 chunk
// in shared memory as the A operand of FFN-2, which accumulates the [128 x d] output tile in a third TMEM buffer across all
// chunks.  Only the pre-activation (the backward's GELU' input) is stored, once.  Per row at d 256 / F 1024 the two-kernel
// form moved 0.5 + 2 + 2 (FFN-1) + 2 + 0.5 + 0.5 + 0.5 (FFN-2) = 8 KB through HBM, this kernel moves 0.5 + 2 + 0.5 + 0.5 + 0.5 = 4 KB
// (2 KB without the saved pre-activation, i.e. in evaluation).
//
// Roles (one CTA per SM, 20 warps, registers re-balanced with setmaxnreg as in ot_gemm.cu):
//   warp 0    TMA producer : tile index -> ring; the A tile (zn, [128 x d], loaded once per tile) and the weight stream:
//                            per chunk W1[g][c*128.., :] ([128 x d], 2 ring slots of 32 KB) and W2[g][:, c*128..] ([d x 128],
//                            2 slots), in exactly the order the MMA warp consumes them
//   warp 1    MMA issuer   : FFN1(c) -> acc1[c&1] (M128 N128 K256), FFN2(c-1) -> acc2 (M128 N256 K128); FFN1(c+1) is issued
//                            before FFN2(c) so that the tensor pipe has work while chunk c sits in the epilogue
//   warps 4-19 epilogue    : four sets of four warps; in a chunk, set s owns 32 of the 128 columns: tcgen05.ld -> +b1 ->
//                            (bf16 pre-activation -> per-warp staging patch -> TMA store) -> GELU -> bf16 h -> swizzled A-operand tile;
//                            at the end of a tile, set s owns 64 of the d columns of the output: +b2, dropout, residual
//                            (TMA-loaded into the staging tile, or the fp32 stream of the NS rows), y -> staging -> TMA store,
//                            y kept in TMEM (tcgen05.st) for the fused RMSNorm pass once the row statistics are exchanged
// Shared memory (227 KB): A tile 64 KB | weight ring 3 x 32 KB | h tile 32 KB | pre-activation staging 32 KB (the last two double
// as the 64 KB staging tile of the final epilogue) | row statistics | barriers.
// TMEM (512 columns): acc1[0] 0-127, acc1[1] 128-255, acc2 256-511.
#include "ot_gemm_common.cuh"
#include "ot_host.h"

namespace ot {

static constexpr int FF_D = 256;                  // model width this kernel is built for (acc2 = 256 TMEM columns)
static constexpr int FF_FC = 128;                 // F chunk
static constexpr int FF_SETS = 4;
static constexpr int FF_SET_THREADS = 128;
static constexpr int FF_EPI_THREADS = FF_SETS * FF_SET_THREADS;
static constexpr int FF_CTRL_THREADS = 128;
static constexpr int FF_THREADS = FF_CTRL_THREADS + FF_EPI_THREADS;
static constexpr int FF_EPI_REGS = 112, FF_CTRL_REGS = 32;   // same balance as ot_gemm.cu (4*32*(96-32) released == 4*128*(112-96) taken)
static constexpr int FF_RING = 3;
static constexpr int FF_SLOT_BYTES = 32 * 1024;
static constexpr int FF_BOX_BYTES = 128 * 128;    // [128 rows x 64 bf16], 128-byte swizzle
static constexpr int FF_TILE_SLOTS = 8;
static constexpr int FF_A_BYTES = 128 * FF_D * 2;                            // 64 KB
static constexpr int FF_OFF_RING = FF_A_BYTES;
static constexpr int FF_OFF_STAGE = FF_OFF_RING + FF_RING * FF_SLOT_BYTES;    // h tile (boxes 0,1) then pre staging (boxes 2,3)
static constexpr int FF_OFF_SS = FF_OFF_STAGE + 4 * FF_BOX_BYTES;
static constexpr int FF_OFF_BARS = FF_OFF_SS + FF_SETS * BM * 4;
static constexpr int FF_SMEM_BYTES = FF_OFF_BARS + 1024;
static_assert(FF_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static constexpr uint32_t FF_T_ACC1 = 0, FF_T_ACC2 = 256;
enum { FF_BAR_SET = 1, FF_BAR_GRP = 5, FF_BAR_ALL = 7 };   // named barriers: 1-4 one per set, 5-6 one per pair of sets, 7 all epilogue warps

struct FfnKParams {
  int F, NC;
  int n_segs, total_mblks;
  int a_transposed;        // always 0; read by decode_tile
  int flags;               // OT_EPI_RESIDUAL | OT_EPI_DROPOUT | OT_EPI_NORM
  int bwd;                 // 1: input-gradient form (ot_ffn_bwd): no biases, chunk epilogue multiplies by gelu'(pre) instead of applying gelu
  GemmSegDev segs[3];
  const float* b1; long long b1_gs;
  const float* b2; long long b2_gs;
  __nv_bfloat16* out; long long ldo;
  __nv_bfloat16* pre; long long ldpre;
  __nv_bfloat16* hout; long long ldh;     // optional copy of h = gelu(pre) for the dW2 weight gradient, stored from the h tile itself
  const __nv_bfloat16* res; long long ldr;
  const float* res_hp; float* out_hp; long long ld_hp; long long hp_row0;
  uint32_t drop_seed, drop_thr16; float drop_scale;
  __nv_bfloat16* norm_out; long long ld_norm; const float* norm_gain; float* norm_rstd; float norm_eps;
  int* sched;
};

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(FF_THREADS, 1)
ot_ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmPre,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmNorm,
                    const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ FfnKParams p) {
  constexpr int D = FF_D;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* ring = smem + FF_OFF_RING;
  uint8_t* stage = smem + FF_OFF_STAGE;               // 4 boxes of 16 KB
  uint8_t* sH = stage;                                // boxes 0, 1: h chunk [128 x 128] as two K slabs
  uint8_t* sP = stage + 2 * FF_BOX_BYTES;             // boxes 2, 3: pre-activation staging
  float* ss_part = reinterpret_cast<float*>(smem + FF_OFF_SS);            // [FF_SETS][BM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FF_OFF_BARS);
  uint64_t* a_full = bars;                   // A tile landed                                   (TMA -> MMA)
  uint64_t* a_empty = bars + 1;              // last FFN-1 MMA of the tile done                 (MMA commit -> producer)
  uint64_t* ring_full = bars + 2;            // [3]
  uint64_t* ring_empty = bars + 5;           // [3]
  uint64_t* acc1_full = bars + 8;            // [2] FFN-1 chunk complete                        (MMA commit -> epilogue)
  uint64_t* acc1_empty = bars + 10;          // [2] 16 arrivals: chunk pulled out of TMEM       (epilogue -> MMA)
  uint64_t* h_full = bars + 12;              // 16 arrivals: h chunk written                    (epilogue -> MMA)
  uint64_t* h_empty = bars + 13;             // FFN-2 of the chunk done reading h               (MMA commit -> epilogue)
  uint64_t* acc2_full = bars + 14;           // output tile complete                            (MMA commit -> epilogue)
  uint64_t* acc2_empty = bars + 15;          // 16 arrivals: output tile pulled out             (epilogue -> MMA)
  uint64_t* res_full = bars + 16;            // [4] residual box landed (one per set)
  uint64_t* tile_full = bars + 20;           // [8]
  uint64_t* tile_empty = bars + 28;          // [8] read by the MMA issuer, the h-store thread, the pre loader and 16 epilogue warps
  uint64_t* h_stored = bars + 36;            // the TMA store of the h chunk has read the h tile  (store thread -> epilogue)
  uint64_t* p_full = bars + 37;              // bwd: pre-activation chunk landed                   (TMA -> epilogue)
  uint64_t* p_free = bars + 38;              // bwd: 16 arrivals: chunk read out                   (epilogue -> pre loader)
  uint64_t* stage_free = bars + 39;          // bwd: 16 arrivals: the previous tile's output stores have read the staging boxes
  int* tile_ring = reinterpret_cast<int*>(bars + 40);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ring + FF_TILE_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NC = p.NC;
  constexpr int KS1 = D / 128;               // ring slots per W1 chunk

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmOut);
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int i = 0; i < FF_RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 4 * FF_SETS); }
    mbar_init(h_full, 4 * FF_SETS); mbar_init(h_empty, 1);
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 4 * FF_SETS);
    for (int i = 0; i < FF_SETS; ++i) mbar_init(&res_full[i], 1);
    for (int i = 0; i < FF_TILE_SLOTS; ++i) { mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 3 + 4 * FF_SETS); }
    mbar_init(h_stored, 1); mbar_init(p_full, 1); mbar_init(p_free, 4 * FF_SETS); mbar_init(stage_free, 4 * FF_SETS);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FF_CTRL_REGS));
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t r = 0;                        // ring slots filled so far
      int n = 0;
      int tile = blockIdx.x;
      auto slot_acquire = [&](uint32_t bytes) -> uint8_t* {
        const uint32_t slot = r % FF_RING;
        mbar_wait_backoff(&ring_empty[slot], ((r / FF_RING) & 1) ^ 1);
        mbar_arrive_expect_tx(&ring_full[slot], bytes);
        return ring + slot * FF_SLOT_BYTES;
      };
      while (true) {
        const int slot = n & (FF_TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_empty[slot], ((n / FF_TILE_SLOTS) & 1) ^ 1);
        tile_ring[slot] = tile;
        mbar_arrive(&tile_full[slot]);
        if (tile < 0) break;
        int nxt;
        if (p.sched != nullptr) nxt = (int)gridDim.x + atomicAdd(p.sched, 1);
        else nxt = tile + (int)gridDim.x;
        if (nxt >= p.total_mblks) nxt = -1;
        const TileInfo t = decode_tile(p, tile);
        // A tile: D/64 K slabs of [128 rows x 64 columns]
        mbar_wait_backoff(a_empty, (n & 1) ^ 1);
        mbar_arrive_expect_tx(a_full, FF_A_BYTES);
#pragma unroll
        for (int ks = 0; ks < D / 64; ++ks) tma_load_2d(sA + ks * FF_BOX_BYTES, &tmA, a_full, ks * 64, t.a_c1);
        const int w1_row = t.group * p.F;    // W1 as [G*F, D]: N = F rows, K = D
        const int w2_row = t.group * D;      // W2 as [G*D, F]: N = D rows, K = F
        for (int c = 0; c <= NC; ++c) {
          if (c < NC) {
#pragma unroll
            for (int ks = 0; ks < KS1; ++ks) {     // W1 chunk c, K half ks: [128 N rows x 128 K] = two K slabs
              uint8_t* dst = slot_acquire(FF_SLOT_BYTES);
              const uint32_t slot_i = r % FF_RING;
              tma_load_2d(dst, &tmW1, &ring_full[slot_i], ks * 128, w1_row + c * FF_FC);
              tma_load_2d(dst + FF_BOX_BYTES, &tmW1, &ring_full[slot_i], ks * 128 + 64, w1_row + c * FF_FC);
              ++r;
            }
          }
          if (c >= 1) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {       // W2 chunk c-1, K quarter ks: [D N rows x 64 K] = one K slab
              uint8_t* dst = slot_acquire(D * 128);
              const uint32_t slot_i = r % FF_RING;
              tma_load_2d(dst, &tmW2, &ring_full[slot_i], (c - 1) * FF_FC + ks * 64, w2_row);
              ++r;
            }
          }
        }
        tile = nxt;
        ++n;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc_bf16(BM, FF_FC, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_bf16(BM, D, 0, 0);
      uint32_t r = 0;
      const uint32_t a_addr = smem_u32(sA), h_addr = smem_u32(sH), ring_addr = smem_u32(ring);
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t slot = r % FF_RING;
        mbar_wait_backoff(&ring_full[slot], (r / FF_RING) & 1, 32);
        tc_fence_after();
        return ring_addr + slot * FF_SLOT_BYTES;
      };
      auto slot_release = [&]() { umma_commit(&ring_empty[r % FF_RING]); ++r; };
      for (int it = 0;; ++it) {
        const int slot = it & (FF_TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / FF_TILE_SLOTS) & 1, 32);
        const int tile = tile_ring[slot];
        mbar_arrive(&tile_empty[slot]);
        if (tile < 0) break;
        mbar_wait_backoff(a_full, it & 1, 32);
        tc_fence_after();
        for (int c = 0; c <= NC; ++c) {
          if (c < NC) {
            // ---- FFN1(c): acc1[g & 1] = zn_tile [128 x D] . W1_chunk^T ----
            const uint32_t g = (uint32_t)it * NC + c;
            const uint32_t b = g & 1;
            mbar_wait_backoff(&acc1_empty[b], ((g >> 1) & 1) ^ 1, 32);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + FF_T_ACC1 + b * FF_FC;
#pragma unroll
            for (int ks = 0; ks < KS1; ++ks) {
              const uint32_t sb = slot_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint64_t adesc = make_smem_desc<128>(a_addr + (ks * 2 + (j >> 2)) * FF_BOX_BYTES, 16) + 2 * (j & 3);
                const uint64_t bdesc = make_smem_desc<128>(sb + (j >> 2) * FF_BOX_BYTES, 16) + 2 * (j & 3);
                umma_bf16_ss(d_tmem, adesc, bdesc, idesc1, (ks | j) != 0 ? 1u : 0u);
              }
              slot_release();
            }
            umma_commit(&acc1_full[b]);
            if (c == NC - 1) umma_commit(a_empty);        // the A tile may be overwritten by the next tile's
          }
          if (c >= 1) {
            // ---- FFN2(c-1): acc2 (+)= h_chunk [128 x 128] . W2_chunk^T ----
            const int cc = c - 1;
            const uint32_t g2 = (uint32_t)it * NC + cc;
            mbar_wait_backoff(h_full, g2 & 1, 32);
            if (cc == 0) mbar_wait_backoff(acc2_empty, (it & 1) ^ 1, 32);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t sb = slot_wait();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t adesc = make_smem_desc<128>(h_addr + ks * FF_BOX_BYTES, 16) + 2 * j;
                const uint64_t bdesc = make_smem_desc<128>(sb, 16) + 2 * j;
                umma_bf16_ss(tmem_base + FF_T_ACC2, adesc, bdesc, idesc2, (cc | ks | j) != 0 ? 1u : 0u);
              }
              slot_release();
            }
            umma_commit(h_empty);
            if (cc == NC - 1) umma_commit(acc2_full);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== h store (optional second output) =====================
    // The bf16 h chunk in shared memory is already laid out as two TMA boxes ([128 rows x 64 columns], 128-byte swizzle): when
    // the caller wants h = gelu(pre) for the dW2 weight gradient (OT/train.py:131), one thread stores the tile as it is - no
    // epilogue instruction is spent on it.  The epilogue waits for the store to have READ the tile before it writes the next chunk.
    if (elect_one()) {
      const bool save_h = p.hout != nullptr;
      for (int it = 0;; ++it) {
        const int slot = it & (FF_TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / FF_TILE_SLOTS) & 1, 32);
        const int tile = tile_ring[slot];
        mbar_arrive(&tile_empty[slot]);
        if (tile < 0) break;
        if (!save_h) continue;
        const TileInfo t = decode_tile(p, tile);
        for (int c = 0; c < NC; ++c) {
          const uint32_t g = (uint32_t)it * NC + c;
          mbar_wait_backoff(h_full, g & 1, 32);
          if (t.valid == BM) {                       // partial tiles are written by the epilogue threads themselves (row masks)
            tma_store_2d(&tmH, sH, c * FF_FC, t.row0);
            tma_store_2d(&tmH, sH + FF_BOX_BYTES, c * FF_FC + 64, t.row0);
            bulk_commit();
            bulk_wait_read0();
          }
          mbar_arrive(h_stored);
        }
      }
      bulk_wait_all();
    }
  } else if (warp == 3) {
    // ===================== pre-activation loader (input-gradient form only) =====================
    // dpre = (dy W2^T) o gelu'(pre): the saved pre-activation chunk [128 x 128] arrives by TMA in the 32 KB that the forward form
    // uses for its per-warp store patches, one chunk ahead of the epilogue.  The same 32 KB are half of the output staging tile
    // of the final epilogue, so a tile's first chunk waits until the previous tile's output stores have read them.
    if (elect_one()) {
      for (int it = 0;; ++it) {
        const int slot = it & (FF_TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / FF_TILE_SLOTS) & 1, 32);
        const int tile = tile_ring[slot];
        mbar_arrive(&tile_empty[slot]);
        if (tile < 0) break;
        if (!p.bwd) continue;
        const TileInfo t = decode_tile(p, tile);
        for (int c = 0; c < NC; ++c) {
          const uint32_t g = (uint32_t)it * NC + c;
          if (g > 0) mbar_wait_backoff(p_free, (g - 1) & 1, 32);
          if (c == 0 && it > 0) mbar_wait_backoff(stage_free, (it - 1) & 1, 32);
          mbar_arrive_expect_tx(p_full, 2 * FF_BOX_BYTES);
          tma_load_2d(sP, &tmRes, p_full, c * FF_FC, t.row0);                       // tmRes maps the pre-activation in this form
          tma_load_2d(sP + FF_BOX_BYTES, &tmRes, p_full, c * FF_FC + 64, t.row0);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..19: four sets of four) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FF_EPI_REGS));
    const int set = (warp - 4) >> 2;
    const int grp = set >> 1;                                              // pair of sets sharing a 64-column box in a chunk
    const int et = threadIdx.x - FF_CTRL_THREADS - set * FF_SET_THREADS;   // 0..127 inside the set
    const int lgrp = warp & 3;
    const int r_own = lgrp * 32 + lane;
    const bool set_io = (et == 0);                     // issues this set's TMA traffic of the final epilogue
    const bool f_res = p.flags & OT_EPI_RESIDUAL, f_drop = p.flags & OT_EPI_DROPOUT, f_norm = p.flags & OT_EPI_NORM;
    const bool bwd = p.bwd != 0;
    const bool save_pre = !bwd && p.pre != nullptr, save_h = p.hout != nullptr;
    uint8_t* box_h = sH + grp * FF_BOX_BYTES;          // chunk phase: h columns [64 grp, +64)
    uint8_t* patch = sP + (warp - 4) * 2048;           // chunk phase: this warp's pre-activation patch, [32 rows x 64 bytes]
    uint8_t* box_y = stage + set * FF_BOX_BYTES;       // final phase: output columns [64 set, +64)
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16);
    uint32_t res_phase = 0;

    for (int it = 0;; ++it) {
      int tile;
      {
        const int slot = it & (FF_TILE_SLOTS - 1);
        mbar_wait_backoff(&tile_full[slot], (it / FF_TILE_SLOTS) & 1, 32);
        tile = tile_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&tile_empty[slot]);
      }
      if (tile < 0) break;
      const TileInfo t = decode_tile(p, tile);
      const bool full = (t.valid == BM);
      const bool hp_tile = f_res && (p.res_hp != nullptr) && (t.row0 >= p.hp_row0);
      // the staging boxes still hold the previous tile's output until its TMA stores have read them
      if (it > 0) {
        if (lane == 0) bulk_wait_read0();
        named_bar_sync(FF_BAR_ALL, FF_EPI_THREADS);
        if (bwd && lane == 0) mbar_arrive(stage_free);       // the pre loader may refill its half of the staging tile
      }

      // -------------------------------- F chunks --------------------------------
      for (int c = 0; c < NC; ++c) {
        const uint32_t g = (uint32_t)it * NC + c;
        const uint32_t b = g & 1;
        mbar_wait_backoff(&acc1_full[b], (g >> 1) & 1, 32);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_x32(t_lane + FF_T_ACC1 + b * FF_FC + set * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);
        f32x2 f[16];
        if (!bwd) {
          const float4* b4 = reinterpret_cast<const float4*>(p.b1 + (long long)t.group * p.b1_gs + c * FF_FC + set * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            f[2 * j] = add2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pk2(bb.x, bb.y));
            f[2 * j + 1] = add2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pk2(bb.z, bb.w));
          }
        } else {
          // input-gradient form: dpre = (dy W2^T) o gelu'(pre), pre from the TMA-loaded chunk (own row, own 32 columns)
          mbar_wait(p_full, g & 1);
          uint32_t w[16];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 q = *reinterpret_cast<const uint4*>(sP + grp * FF_BOX_BYTES + swz_off<128>(r_own, (set & 1) * 4 + ch));
            w[ch * 4 + 0] = q.x; w[ch * 4 + 1] = q.y; w[ch * 4 + 2] = q.z; w[ch * 4 + 3] = q.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(p_free);
#pragma unroll
          for (int e = 0; e < 16; ++e)
            f[e] = mul2(pk2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), gelu_erf_grad2(pk2(bf16lo(w[e]), bf16hi(w[e]))));
        }
        if (save_pre) {
          // pre-activation -> this warp's private [32 rows x 32 columns] staging patch (64-byte swizzle) -> HBM by one TMA store
          // per warp and chunk: no barrier between warps on this path (the first build staged per pair of sets behind two
          // 256-thread barriers and a store-drain wait: + 1700 cycles per chunk, profiles/README.md).  It is the only
          // [rows, F] tensor this kernel writes.
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { float a, bq; upk2(f[j], a, bq); pk[j] = pack_bf16x2(a, bq); }
          if (full) {
            if (lane == 0) bulk_wait_read0();                    // this warp's previous store has read the patch
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              *reinterpret_cast<uint4*>(patch + swz_off<64>(lane, ch)) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&tmPre, patch, c * FF_FC + set * 32, t.row0 + lgrp * 32); bulk_commit(); }
          } else if (r_own < t.valid) {
            // partial tile (small batches): the row's 64 bytes straight from registers
            uint4* dst = reinterpret_cast<uint4*>(p.pre + (long long)(t.row0 + r_own) * p.ldpre + c * FF_FC + set * 32);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) dst[ch] = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
          }
        }
        if (!bwd) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = gelu_erf2(f[j]);
        }
        if (g > 0) {
          mbar_wait(h_empty, (g - 1) & 1);                     // FFN2 of the previous chunk has read the h tile
          if (save_h) mbar_wait(h_stored, (g - 1) & 1);        // ... and so has its TMA store
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float a0, a1, a2, a3, a4, a5, a6, a7;
          upk2(f[ch * 4 + 0], a0, a1); upk2(f[ch * 4 + 1], a2, a3); upk2(f[ch * 4 + 2], a4, a5); upk2(f[ch * 4 + 3], a6, a7);
          uint4 q;
          q.x = pack_bf16x2(a0, a1); q.y = pack_bf16x2(a2, a3); q.z = pack_bf16x2(a4, a5); q.w = pack_bf16x2(a6, a7);
          *reinterpret_cast<uint4*>(box_h + swz_off<128>(r_own, (set & 1) * 4 + ch)) = q;
          if (save_h && !full && r_own < t.valid)
            *reinterpret_cast<uint4*>(p.hout + (long long)(t.row0 + r_own) * p.ldh + c * FF_FC + set * 32 + ch * 8) = q;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full);
      }

      // -------------------------------- output tile --------------------------------
      mbar_wait_backoff(acc2_full, it & 1, 32);
      if (save_h) mbar_wait(h_stored, ((uint32_t)it * NC + NC - 1) & 1);   // the last chunk's h store has read the h tile too
      tc_fence_after();
      if (save_pre && lane == 0) bulk_wait_read0();            // this warp's last pre-activation store has read its patch
      named_bar_sync(FF_BAR_ALL, FF_EPI_THREADS);              // h tile free (acc2_full), pre staging free: 4 boxes of staging
      const int col0 = set * 64;
      const bool tile_res = f_res && !hp_tile;
      if (tile_res) {
        if (full) {
          if (set_io) { mbar_arrive_expect_tx(&res_full[set], FF_BOX_BYTES); tma_load_2d(box_y, &tmRes, &res_full[set], col0, t.row0); }
          mbar_wait(&res_full[set], res_phase);
          res_phase ^= 1;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + et, rr = idx >> 3, ch = idx & 7;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (rr < t.valid) q = *reinterpret_cast<const uint4*>(p.res + (long long)(t.row0 + rr) * p.ldr + col0 + ch * 8);
            *reinterpret_cast<uint4*>(box_y + swz_off<128>(rr, ch)) = q;
          }
          named_bar_sync(FF_BAR_SET + set, FF_SET_THREADS);
        }
      }
      f32x2 ss2 = pk2(0.0f);
      const float* bias2 = bwd ? nullptr : p.b2 + (long long)t.group * p.b2_gs + col0;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld_x32(t_lane + FF_T_ACC2 + col0 + half * 32, v);
        tmem_ld_wait();
        f32x2 f[16];
        if (!bwd) {
          const float4* b4 = reinterpret_cast<const float4*>(bias2 + half * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            f[2 * j] = add2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pk2(bb.x, bb.y));
            f[2 * j + 1] = add2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pk2(bb.z, bb.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        }
        if (f_drop) {   // inverted dropout on the branch output, before the residual add (OT/model.py:198)
          const uint32_t grow = static_cast<uint32_t>(t.row0 + r_own);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t hb = dropout_bits(p.drop_seed, grow, static_cast<uint32_t>(col0 + half * 32 + 2 * j), static_cast<uint32_t>(D));
            float a, bq;
            upk2(f[j], a, bq);
            a = ((hb & 0xFFFFu) >= p.drop_thr16) ? a * p.drop_scale : 0.0f;
            bq = ((hb >> 16) >= p.drop_thr16) ? bq * p.drop_scale : 0.0f;
            f[j] = pk2(a, bq);
          }
        }
        if (hp_tile) {   // fp32 residual in, fp32 result out (NS-token rows)
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
        } else if (tile_res) {
          uint32_t w[16];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 q = *reinterpret_cast<const uint4*>(box_y + swz_off<128>(r_own, half * 4 + ch));
            w[ch * 4 + 0] = q.x; w[ch * 4 + 1] = q.y; w[ch * 4 + 2] = q.z; w[ch * 4 + 3] = q.w;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] = add2(f[e], pk2(bf16lo(w[e]), bf16hi(w[e])));
        }
        if (f_norm) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ss2 = fma2(f[j], f[j], ss2);
          // keep the fp32 row in TMEM for the normalisation pass (the accumulator is dead once read)
#pragma unroll
          for (int j = 0; j < 16; ++j) { float a, bq; upk2(f[j], a, bq); v[2 * j] = __float_as_uint(a); v[2 * j + 1] = __float_as_uint(bq); }
          tmem_st_x32(t_lane + FF_T_ACC2 + col0 + half * 32, v);
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float a0, a1, a2, a3, a4, a5, a6, a7;
          upk2(f[ch * 4 + 0], a0, a1); upk2(f[ch * 4 + 1], a2, a3); upk2(f[ch * 4 + 2], a4, a5); upk2(f[ch * 4 + 3], a6, a7);
          uint4 q;
          q.x = pack_bf16x2(a0, a1); q.y = pack_bf16x2(a2, a3); q.z = pack_bf16x2(a4, a5); q.w = pack_bf16x2(a6, a7);
          *reinterpret_cast<uint4*>(box_y + swz_off<128>(r_own, half * 4 + ch)) = q;
        }
      }
      auto store_box = [&](__nv_bfloat16* dst, long long ld, const CUtensorMap* tm) {
        if (full) fence_proxy_async_smem();
        named_bar_sync(FF_BAR_SET + set, FF_SET_THREADS);
        if (full) {
          if (set_io) { tma_store_2d(tm, box_y, col0, t.row0); bulk_commit(); }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + et, rr = idx >> 3, ch = idx & 7;
            if (rr < t.valid)
              *reinterpret_cast<uint4*>(dst + (long long)(t.row0 + rr) * ld + col0 + ch * 8) =
                  *reinterpret_cast<const uint4*>(box_y + swz_off<128>(rr, ch));
          }
          named_bar_sync(FF_BAR_SET + set, FF_SET_THREADS);      // readers done before the box is written again
        }
      };
      store_box(p.out, p.ldo, &tmOut);
      if (f_norm) {
        // ---- fused RMSNorm of the finished rows (OT/model.py:19-23; the next block's norm1, OT/model.py:191) ----
        float ss, ss_hi;
        upk2(ss2, ss, ss_hi);
        ss_part[set * BM + r_own] = ss + ss_hi;
        tmem_st_wait();
        if (full && set_io) bulk_wait_read0();                   // the y rows have left the staging box
        named_bar_sync(FF_BAR_ALL, FF_EPI_THREADS);
        const float tot = (ss_part[r_own] + ss_part[BM + r_own]) + (ss_part[2 * BM + r_own] + ss_part[3 * BM + r_own]);
        const float rstd = rsqrtf(tot / (float)D + p.norm_eps);
        if (set == 0 && r_own < t.valid && p.norm_rstd != nullptr) p.norm_rstd[t.row0 + r_own] = rstd;
        const f32x2 rstd2 = pk2(rstd);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_x32(t_lane + FF_T_ACC2 + col0 + half * 32, v);
          tmem_ld_wait();
          const float4* g4 = reinterpret_cast<const float4*>(p.norm_gain + col0 + half * 32);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 ga = __ldg(g4 + 2 * ch), gb = __ldg(g4 + 2 * ch + 1);
            const f32x2 x0 = mul2(mul2(pk2(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1])), rstd2), pk2(ga.x, ga.y));
            const f32x2 x1 = mul2(mul2(pk2(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3])), rstd2), pk2(ga.z, ga.w));
            const f32x2 x2 = mul2(mul2(pk2(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5])), rstd2), pk2(gb.x, gb.y));
            const f32x2 x3 = mul2(mul2(pk2(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7])), rstd2), pk2(gb.z, gb.w));
            float a0, a1, a2, a3, a4, a5, a6, a7;
            upk2(x0, a0, a1); upk2(x1, a2, a3); upk2(x2, a4, a5); upk2(x3, a6, a7);
            uint4 q;
            q.x = pack_bf16x2(a0, a1); q.y = pack_bf16x2(a2, a3); q.z = pack_bf16x2(a4, a5); q.w = pack_bf16x2(a6, a7);
            *reinterpret_cast<uint4*>(box_y + swz_off<128>(r_own, half * 4 + ch)) = q;
          }
        }
        store_box(p.norm_out, p.ld_norm, &tmNorm);
      }
      // every TMEM access to acc2 by this warp is complete -> the MMA warp may start the next tile's FFN-2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// -------------------------------------------------------------------------------------------------
// host launcher
// -------------------------------------------------------------------------------------------------
static int make_2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, long long ld, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {64u, box_rows};
  return make_tmap_bf16(tm, base, 2, dims, str, box, 128);
}

static int ffn_launch(const ot_ffn_params* p, cudaStream_t st, bool bwd) {
  const char* who = bwd ? "ot_ffn_bwd" : "ot_ffn_fwd";
  if (!p || !p->zn || !p->W1 || !p->W2 || !p->out) OT_FAIL(OT_ERR_INVALID_ARG, "%s: null pointer", who);
  if (!bwd && (!p->b1 || !p->b2)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: null bias");
  if (bwd && (!p->pre || !p->h || p->flags != 0 || p->res || p->res_hp))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_bwd: needs pre (saved pre-activation, input) and h (dpre, output); no flags / residual");
  if (p->d != FF_D) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ffn_fwd: d=%d (this kernel is built for d=%d; use the two-GEMM path)", p->d, FF_D);
  if (p->F <= 0 || p->F % FF_FC != 0) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ffn_fwd: F=%d not a multiple of %d", p->F, FF_FC);
  if (p->n_segs < 1 || p->n_segs > 3) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: n_segs=%d", p->n_segs);
  if (!bwd && ((reinterpret_cast<uintptr_t>(p->b1) & 15) || (reinterpret_cast<uintptr_t>(p->b2) & 15) || (p->b1_group_stride % 4) || (p->b2_group_stride % 4)))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: biases must be 16-byte aligned with group strides that are multiples of 4");
  if ((p->ldzn % 8) || (p->ldo % 8) || (p->pre && (p->ldpre % 8)) || (p->h && (p->ldh % 8)) || (p->res && (p->ldr % 8)) || (p->ldw1 % 8) || (p->ldw2 % 8))
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ffn_fwd: leading dimensions must be multiples of 8 elements");
  if ((p->flags & OT_EPI_RESIDUAL) && !p->res) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: residual flag without res");
  if (p->flags & ~(OT_EPI_RESIDUAL | OT_EPI_DROPOUT | OT_EPI_NORM)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: unsupported flags %d", p->flags);
  if ((p->flags & OT_EPI_NORM) && (!p->norm_out || !p->norm_gain || (p->ld_norm % 8) || (reinterpret_cast<uintptr_t>(p->norm_gain) & 15)))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: OT_EPI_NORM needs norm_out, a 16-byte aligned norm_gain and ld_norm %% 8 == 0");

  FfnKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.F = p->F; kp.NC = p->F / FF_FC; kp.n_segs = p->n_segs; kp.flags = p->flags; kp.bwd = bwd ? 1 : 0;
  long long row_extent = 0;
  for (int s = 0; s < p->n_segs; ++s) {
    const ot_gemm_seg& sg = p->segs[s];
    if (sg.n_units <= 0 || sg.rows_per_unit <= 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: empty segment %d", s);
    const int last_group = sg.group_start + (sg.n_units - 1) * sg.group_stride;
    if (sg.group_start < 0 || last_group >= p->n_groups)
      OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: segment %d addresses weight group %d of %d", s, last_group, p->n_groups);
    if (sg.a_row_start != sg.row_start) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: input and output rows must coincide (segment %d)", s);
    const long long e = (long long)sg.row_start + (long long)sg.n_units * sg.rows_per_unit;
    if (e > row_extent) row_extent = e;
  }
  kp.total_mblks = fill_seg_table(kp.segs, p->segs, p->n_segs);
  if (kp.total_mblks >= (1 << 24)) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_ffn_fwd: %d tiles exceed the tile-index range", kp.total_mblks);
  kp.b1 = p->b1; kp.b1_gs = p->b1_group_stride; kp.b2 = p->b2; kp.b2_gs = p->b2_group_stride;
  kp.out = (__nv_bfloat16*)p->out; kp.ldo = p->ldo; kp.pre = (__nv_bfloat16*)p->pre; kp.ldpre = p->ldpre;
  kp.hout = (__nv_bfloat16*)p->h; kp.ldh = p->ldh;
  kp.res = (const __nv_bfloat16*)p->res; kp.ldr = p->ldr;
  kp.res_hp = p->res_hp; kp.out_hp = p->out_hp; kp.ld_hp = p->ld_hp; kp.hp_row0 = p->hp_row0;
  kp.norm_out = (__nv_bfloat16*)p->norm_out; kp.ld_norm = p->ld_norm; kp.norm_gain = p->norm_gain; kp.norm_rstd = p->norm_rstd;
  kp.norm_eps = p->norm_eps;
  if (p->flags & OT_EPI_DROPOUT) {
    if (!(p->drop_rate >= 0.0f && p->drop_rate < 1.0f)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: drop_rate=%f", (double)p->drop_rate);
    kp.drop_seed = p->drop_seed; kp.drop_thr16 = (uint32_t)(p->drop_rate * 65536.0f + 0.5f); kp.drop_scale = 1.0f / (1.0f - p->drop_rate);
  }
  if (p->res_hp || p->out_hp) {
    if (!(p->flags & OT_EPI_RESIDUAL) || !p->res_hp || !p->out_hp || (p->ld_hp % 4))
      OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: res_hp/out_hp need OT_EPI_RESIDUAL, both pointers and ld_hp %% 4 == 0");
    bool ok = false;
    for (int s2 = 0; s2 < p->n_segs; ++s2) ok = ok || (p->segs[s2].row_start == p->hp_row0);
    if (!ok) OT_FAIL(OT_ERR_INVALID_ARG, "ot_ffn_fwd: hp_row0=%lld is not the first row of a segment", (long long)p->hp_row0);
  }
  kp.sched = sched_slot(st);

  CUtensorMap tmA, tmW1, tmW2, tmPre, tmOut, tmNorm, tmRes, tmH;
  int rc;
  if ((rc = make_2d(&tmA, p->zn, FF_D, (uint64_t)row_extent, p->ldzn, BM))) return rc;
  if ((rc = make_2d(&tmW1, p->W1, FF_D, (uint64_t)p->n_groups * p->F, p->ldw1, FF_FC))) return rc;
  if ((rc = make_2d(&tmW2, p->W2, (uint64_t)p->F, (uint64_t)p->n_groups * FF_D, p->ldw2, FF_D))) return rc;
  if ((rc = make_2d(&tmOut, p->out, FF_D, (uint64_t)row_extent, p->ldo, BM))) return rc;
  tmPre = tmOut; tmNorm = tmOut; tmRes = tmOut; tmH = tmOut;
  if (p->h && (rc = make_2d(&tmH, p->h, (uint64_t)p->F, (uint64_t)row_extent, p->ldh, BM))) return rc;
  if (bwd) {      // the saved pre-activation is an INPUT here: [128 x 64] boxes like a residual tile
    if ((rc = make_2d(&tmRes, p->pre, (uint64_t)p->F, (uint64_t)row_extent, p->ldpre, BM))) return rc;
  } else if (p->pre) {   // per-warp patches: 32 columns x 32 rows, 64-byte swizzle
    uint64_t dims[2] = {(uint64_t)p->F, (uint64_t)row_extent};
    uint64_t str[1] = {(uint64_t)p->ldpre * 2};
    uint32_t box[2] = {32u, 32u};
    if ((rc = make_tmap_bf16(&tmPre, p->pre, 2, dims, str, box, 64))) return rc;
  }
  if ((p->flags & OT_EPI_NORM) && (rc = make_2d(&tmNorm, p->norm_out, FF_D, (uint64_t)row_extent, p->ld_norm, BM))) return rc;
  if ((p->flags & OT_EPI_RESIDUAL) && (rc = make_2d(&tmRes, p->res, FF_D, (uint64_t)row_extent, p->ldr, BM))) return rc;

  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(ot_ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    attr_done = true;
  }
  const int grid = kp.total_mblks < num_sms() ? kp.total_mblks : num_sms();
  ot_ffn_fused_kernel<<<grid, FF_THREADS, FF_SMEM_BYTES, st>>>(tmA, tmW1, tmW2, tmPre, tmOut, tmNorm, tmRes, tmH, kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int ffn_fwd_impl(const ot_ffn_params* p, cudaStream_t st) { return ffn_launch(p, st, false); }
int ffn_bwd_impl(const ot_ffn_params* p, cudaStream_t st) { return ffn_launch(p, st, true); }

}  // namespace ot
