// ot_gemm_common.cuh — pieces shared by the grouped-GEMM kernels (ot_gemm.cu, ot_ffn_fused.cu): the device copy of the
// position-segment table (which rows use which weight group, OT/model.py:67-74 as data) and the tile decoder.
#pragma once
#include "ot_common.cuh"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int BM = 128;               // rows per tile (UMMA M)

struct GemmSegDev {
  int row_start, n_units, rows_per_unit, group_start, group_stride, a_row_start;
  int mblk_start, mblk_per_unit;
  uint32_t mpu_rcp;        // ceil(2^32 / mblk_per_unit): unit = umulhi(local, mpu_rcp), exact for local < 2^32 / mblk_per_unit
};

// host: fill a GemmSegDev table from the C-ABI segments; returns the number of 128-row M blocks
inline int fill_seg_table(GemmSegDev* dst, const ot_gemm_seg* segs, int n_segs) {
  int mb = 0;
  for (int s = 0; s < n_segs; ++s) {
    const ot_gemm_seg& sg = segs[s];
    GemmSegDev& d = dst[s];
    d.row_start = sg.row_start; d.n_units = sg.n_units; d.rows_per_unit = sg.rows_per_unit;
    d.group_start = sg.group_start; d.group_stride = sg.group_stride; d.a_row_start = sg.a_row_start;
    d.mblk_start = mb; d.mblk_per_unit = (sg.rows_per_unit + BM - 1) / BM;
    d.mpu_rcp = (uint32_t)(((1ull << 32) + d.mblk_per_unit - 1) / d.mblk_per_unit);
    mb += d.mblk_per_unit * sg.n_units;
  }
  return mb;
}

struct TileInfo {
  int row0, valid, group, a_c1, a_c2;
};

__device__ __forceinline__ int div_rcp(int n, uint32_t rcp, int d) {
  return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), rcp));
}

// Every epilogue warp decodes every tile, so this runs ~16 x tiles times per CTA: no divisions (host-made reciprocals)
// and no dynamic indexing of the kernel parameters (that would copy the segment table to local memory).
template <typename KP>
__device__ __forceinline__ TileInfo decode_tile(const KP& p, int mblk) {
  const bool s2 = p.n_segs > 2 && mblk >= p.segs[2].mblk_start;
  const bool s1 = !s2 && p.n_segs > 1 && mblk >= p.segs[1].mblk_start;
#define OT_SEG(f) (s2 ? p.segs[2].f : s1 ? p.segs[1].f : p.segs[0].f)
  const int mblk_per_unit = OT_SEG(mblk_per_unit);
  const int rows_per_unit = OT_SEG(rows_per_unit);
  const int local = mblk - OT_SEG(mblk_start);
  const int unit = div_rcp(local, OT_SEG(mpu_rcp), mblk_per_unit);
  const int sub = local - unit * mblk_per_unit;
  const int riu = sub * BM;
  TileInfo t;
  t.row0 = OT_SEG(row_start) + unit * rows_per_unit + riu;
  t.valid = min(BM, rows_per_unit - riu);
  t.group = OT_SEG(group_start) + unit * OT_SEG(group_stride);
  if (!p.a_transposed) {
    t.a_c1 = OT_SEG(a_row_start) + unit * rows_per_unit + riu;
    t.a_c2 = 0;
  } else {
    t.a_c1 = OT_SEG(a_row_start) + unit;
    t.a_c2 = riu;
  }
#undef OT_SEG
  return t;
}

}  // namespace ot
