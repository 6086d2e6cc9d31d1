// ot_attn.cuh — shared pieces of the attention kernels (forward, dQ, dK/dV).
//
// Geometry (OT/model.py:101-114 with pyramid tail queries):
//   queries  = the last Lq token positions of the sequence, keys/values = all Lk positions;
//   query i (0-based inside the tail) sits at absolute position off+i, off = Lk-Lq, and may attend
//   keys 0..off+i  (tf.linalg.band_part(ones,-1,0) aligned to the sequence tail, OT/model.py:64,109);
//   masked scores contribute exactly 0 probability (the reference replaces them by -1e9, whose
//   exp underflows to 0 in fp32).
// Memory: token-major rows, element (l, b, h, e) of a [L*B, ld] buffer at (l*B + b)*ld + h*DH + e.
// Tiles: 128 queries x 128 keys; every operand tile is TMA-loaded once as [128 rows x DH] in the
// swizzled slab layout and then used as a K-major or an MN-major UMMA operand as the product needs.
#pragma once
#include "ot_common.cuh"

namespace ot {

template <int DH, int SWB>
struct AttnTile {
  static constexpr int SLABC = SWB / 2;                // columns (of DH) per slab
  static constexpr int NSLAB = DH / SLABC;
  static constexpr int SLAB_BYTES = 128 * SWB;         // 128 rows x SWB bytes
  static constexpr int TILE_BYTES = NSLAB * SLAB_BYTES;  // = 128*DH*2
  static constexpr int KSTEPS_PER_SLAB = SLABC / 16;
  static_assert(DH % SLABC == 0, "head dim must be a whole number of slabs");
};

// descriptors for a [128 x DH] tile used K-major (contraction over DH), K-step kk (16 elements)
template <int DH, int SWB>
__device__ __forceinline__ uint64_t tile_desc_kmajor(uint32_t tile_addr, int kk) {
  using T = AttnTile<DH, SWB>;
  const int slab = kk / T::KSTEPS_PER_SLAB;
  const int within = kk - slab * T::KSTEPS_PER_SLAB;
  return make_smem_desc<SWB>(tile_addr + slab * T::SLAB_BYTES, 16) + 2 * within;
}
// descriptor for a [128 x DH] tile used MN-major (contraction over the 128 rows), K-step kk (16 rows)
template <int DH, int SWB>
__device__ __forceinline__ uint64_t tile_desc_mnmajor(uint32_t tile_addr, int kk) {
  using T = AttnTile<DH, SWB>;
  return make_smem_desc<SWB>(tile_addr, T::SLAB_BYTES) + static_cast<uint64_t>((16 * SWB) >> 4) * kk;
}
// [128 x 128] bf16 probability-like tile, always 128-byte swizzle: two slabs of 64 columns.
static constexpr int PT_SLAB_BYTES = 128 * 128;
static constexpr int PT_BYTES = 2 * PT_SLAB_BYTES;
// K-major use (contraction over the 128 columns), K-step kk of 16 columns
__device__ __forceinline__ uint64_t ptile_desc_kmajor(uint32_t addr, int kk) {
  return make_smem_desc<128>(addr + (kk >> 2) * PT_SLAB_BYTES, 16) + 2 * (kk & 3);
}
// MN-major use (contraction over the 128 rows; the 128 columns are M), K-step kk of 16 rows
__device__ __forceinline__ uint64_t ptile_desc_mnmajor(uint32_t addr, int kk) {
  return make_smem_desc<128>(addr, PT_SLAB_BYTES) + static_cast<uint64_t>((16 * 128) >> 4) * kk;
}
// write 32 consecutive columns [c0, c0+32) of row `row` of a P tile from fp32 registers
__device__ __forceinline__ void ptile_store32(uint8_t* tile, int row, int c0, const float (&p)[32]) {
  uint8_t* slab = tile + (c0 >> 6) * PT_SLAB_BYTES;
  const int ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 q;
    q.x = pack_bf16x2(p[ch * 8 + 0], p[ch * 8 + 1]);
    q.y = pack_bf16x2(p[ch * 8 + 2], p[ch * 8 + 3]);
    q.z = pack_bf16x2(p[ch * 8 + 4], p[ch * 8 + 5]);
    q.w = pack_bf16x2(p[ch * 8 + 6], p[ch * 8 + 7]);
    *reinterpret_cast<uint4*>(slab + swz_off<128>(row, ch0 + ch)) = q;
  }
}

// TMA-load one [128 rows x DH] tile: rows = token positions l0.., fixed sample b, head h.
template <int DH, int SWB>
__device__ __forceinline__ void load_head_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int h, int b, int l0) {
  using T = AttnTile<DH, SWB>;
#pragma unroll
  for (int s = 0; s < T::NSLAB; ++s) tma_load_3d(dst + s * T::SLAB_BYTES, tm, bar, h * DH + s * T::SLABC, b, l0);
}

}  // namespace ot
