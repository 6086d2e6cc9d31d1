// ot_attn_fwd_common.cuh — pieces shared by the head_dim-64 attention forward kernels v3 / v4 (ot_attn_fwd_v3.cu, ot_attn_fwd_v4.cu):
// shared-memory-address mbarrier helpers, the TMA tile store, and the software-pipelined softmax chunk.
#pragma once
#include "ot_attn.cuh"

namespace ot {

static constexpr float F3_TAU = 8.0f;                                  // lazy rescale threshold, log2 units
#ifndef OT_F3_POLY
#define OT_F3_POLY 0      // pairs (of the four in a group of eight exponentials) evaluated by the FMA-pipe polynomial: 0 .. 4 (experiment;
                          // measured slower.  The kernels now pass -inf for hidden scores, which only ex2.approx maps to 0: keep it 0)
#endif

// OT_HANG_DEBUG=1 (compile time, debugging only): a wait that runs into the hang guard records who waited for what in a device-side
// record (first writer wins) and then carries on as if the wait had succeeded, so that the launch ends and the record can be read
// (ot_debug_hang_read, exported by ot_attn_fwd_v4.cu) instead of the context dying in a trap.
#ifndef OT_HANG_DEBUG
#define OT_HANG_DEBUG 0
#endif
#if OT_HANG_DEBUG
static __device__ unsigned int ot_hang_rec[64];     // [0] claiming block + 1, [1] entries, then (thread | what << 16, a, b) per entry of that block
__device__ __forceinline__ void ot_hang_note(uint32_t what, uint32_t a, uint32_t b) {
  if ((threadIdx.x & 31u) != 0u && what == 1u && __activemask() != 1u) return;   // one record per warp
  const unsigned int me = blockIdx.x + 1u;
  const unsigned int owner = atomicCAS(&ot_hang_rec[0], 0u, me);
  if (owner == 0u || owner == me) {
    const unsigned int i = atomicAdd(&ot_hang_rec[1], 1u);
    if (i < 20u) { ot_hang_rec[2 + 3 * i] = threadIdx.x | (what << 16); ot_hang_rec[3 + 3 * i] = a; ot_hang_rec[4 + 3 * i] = b; }
  }
}
#endif

// ---- shared-memory-address forms of the mbarrier helpers (no generic-to-shared conversion in the inner loops) ----
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (ok) return;
  uint32_t spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
#if OT_HANG_DEBUG
    if (++spins > (1u << 12)) { ot_hang_note(1u, bar, parity); return; }
#elif OT_HANG_GUARD
    if (++spins > (1u << 17)) __trap();
#endif
  } while (!ok);
}
__device__ __forceinline__ bool mbar_test_s(uint32_t bar, uint32_t parity) {      // non-blocking
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// One 32-column chunk of a score row -> probabilities (packed bf16 pairs) and partial row sums, software-pipelined in groups of
// eight (arguments of group g+1 and sums / packs of group g-1 around the exponentials of group g).  The mask-free form works on
// packed fp32 pairs (FFMA2 for the arguments, FADD2 for the row sums: half the issue slots of the scalar form).  Since the callers
// overwrite the hidden scores of a chunk cut by the diagonal with -inf before the row maximum, they use the mask-free form for every
// chunk; the MASK form is kept for reference.
template <bool MASK>
__device__ __forceinline__ void f3_softmax_chunk(const uint32_t (&s)[32], uint32_t (&pk)[16], float scale_log2, float mb, int lim_rel,
                                                 f32x2 (&rs)[2]) {
  float a[32];
  const f32x2 c2 = pk2(scale_log2), nmb2 = pk2(-mb);
#pragma unroll
  for (int i = 0; i < 4; ++i) upk2(fma2(pk2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, nmb2), a[2 * i], a[2 * i + 1]);
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    if (g + 1 < 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        upk2(fma2(pk2(__uint_as_float(s[(g + 1) * 8 + 2 * i]), __uint_as_float(s[(g + 1) * 8 + 2 * i + 1])), c2, nmb2), a[(g + 1) * 8 + 2 * i],
             a[(g + 1) * 8 + 2 * i + 1]);
    }
    if (g < 4) {
      // exponentials of group g: OT_F3_POLY of its four pairs go to the FMA-pipe polynomial, the rest to the XU pipe
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= 4 - OT_F3_POLY) {
          upk2(ex2_poly2(pk2(a[g * 8 + 2 * i], a[g * 8 + 2 * i + 1])), a[g * 8 + 2 * i], a[g * 8 + 2 * i + 1]);
        } else {
          a[g * 8 + 2 * i] = ex2_approx(a[g * 8 + 2 * i]);
          a[g * 8 + 2 * i + 1] = ex2_approx(a[g * 8 + 2 * i + 1]);
        }
      }
    }
    if (g >= 1) {
      const int b0 = (g - 1) * 8;
      if (MASK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[b0 + i] = (b0 + i <= lim_rel) ? a[b0 + i] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i & 1] = add2(rs[i & 1], pk2(a[b0 + 2 * i], a[b0 + 2 * i + 1]));
        pk[(b0 >> 1) + i] = pack_bf16x2(a[b0 + 2 * i], a[b0 + 2 * i + 1]);
      }
    }
  }
}

}  // namespace ot
