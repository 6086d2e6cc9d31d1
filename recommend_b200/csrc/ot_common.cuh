// ot_common.cuh — sm_100a building blocks shared by every OneTrans kernel:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), UMMA shared-memory
// and instruction descriptors.  Everything here is inline PTX; nothing is borrowed from a library.
//
// Conventions
//   * all tiles that feed tcgen05.mma live in shared memory in the canonical swizzled layouts that a
//     TMA box load with CU_TENSOR_MAP_SWIZZLE_{64B,128B} produces ("slabs": rows x SWB bytes, 8-row
//     swizzle atoms of 8*SWB bytes, slab base aligned to 1024 B);
//   * "K-major"  operand: the slab row is an M/N index, the bytes along the row run along K;
//   * "MN-major" operand: the slab row is a K index,  the bytes along the row run along M/N.
//     The same bytes can therefore serve as a K-major operand of one product and as an MN-major
//     operand of the transposed product (used by attention and by the weight-gradient kernel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef OT_HANG_GUARD
#define OT_HANG_GUARD 1
#endif

namespace ot {

static constexpr int kNumSMsB200 = 148;

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Same packing on the integer pipe (round half away from zero instead of half to even; finite inputs only): two adds
// and one byte permute instead of one F2FP.  For inner loops whose conversion pipe is the busy one.
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
  return __byte_perm(a, b, 0x7632);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware may park the thread for up to `ns` nanoseconds and wakes it when the phase
// completes, so a waiting warp issues (almost) no instructions.  Round-2 finding (profiles/README.md): spin / nanosleep poll
// loops of waiting warps were 22 % of all issued instructions of the fused FFN kernel.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
#ifndef OT_WAIT_HINT_NS
#define OT_WAIT_HINT_NS 20000     // 0 = plain try_wait spin (and nanosleep back-off where asked), the round-1 behaviour
#endif
// Blocking wait.  With OT_HANG_GUARD a wait that never completes traps (the launch fails loudly)
// instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if OT_WAIT_HINT_NS > 0
  if (mbar_try_wait(bar, parity)) return;       // the common fast path: already complete
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, OT_WAIT_HINT_NS)) {
#if OT_HANG_GUARD
    if (++spins > (1u << 17)) { __trap(); }      // ~2.6 s of full-length suspensions: a hang, not a wait
#endif
  }
#elif OT_HANG_GUARD
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { __trap(); }
  }
#else
  while (!mbar_try_wait(bar, parity)) {}
#endif
}

// Same wait for roles that are far off the critical path (GEMM producer / MMA issuer / epilogue waiting for an
// accumulator).  Round 1 slept between polls (__nanosleep) so that the spinning warp stopped competing for issue slots with the
// math warps on its scheduler; with the suspend-time hint the hardware does the parking and wakes the warp on completion.
#ifndef OT_BACKOFF
#define OT_BACKOFF 1      // 0: experiment build, every back-off wait becomes a plain try_wait spin (profiles/README.md, round 2)
#endif
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns = 64) {
#if OT_WAIT_HINT_NS > 0 || !OT_BACKOFF
  (void)ns;
  mbar_wait(bar, parity);
#elif OT_HANG_GUARD
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++spins > (1u << 24)) { __trap(); }
  }
#else
  while (!mbar_try_wait(bar, parity)) { __nanosleep(ns); }
#endif
}

// generic-proxy writes to shared memory -> visible to the async proxy (TMA store, tcgen05.mma reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barrier among a subset of warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA loads (global -> shared, completion on an mbarrier)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA store (shared -> global, bulk async-group completion).  The smem tile must have been written by
// generic-proxy stores followed by fence_proxy_async_smem() and a barrier before the issuing thread gets here.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have completed entirely (before the CTA exits)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: tensor memory allocation
// ----------------------------------------------------------------------------------------------
// Executed by ONE full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: MMA issue / commit
// ----------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (lane = row of A, 32-bit column c holds the bf16
// pair K = 2c, 2c+1; one K = 16 step is 8 columns), always K-major.  Measured (profiles/exp_mma_rate.cu): a 128 x 64 x 16 product
// costs ~36 clk in this form against ~52 clk with both operands in shared memory, whose 128 B/clk operand reads are the limit.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation.
//   bit 4-5 D format (1 = f32), 7-9 A format (1 = bf16), 10-12 B format (1 = bf16),
//   15 A major (0 = K, 1 = MN), 16 B major, 17-22 N>>3, 24-28 M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// Shared-memory matrix descriptor (sm_100 format, version field = 1).
//   SWB = swizzle width in bytes (128 or 64).  layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
//   K-major : SBO = 8*SWB (next 8 rows of M/N), LBO unused.
//   MN-major: LBO = byte distance between consecutive slabs along M/N (slab = SWB bytes wide),
//             SBO = 8*SWB (next 8 rows of K).
template <int SWB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  static_assert(SWB == 128 || SWB == 64, "swizzle width");
  constexpr uint64_t layout = (SWB == 128) ? 2ull : 4ull;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(((8u * SWB) >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld: TMEM -> registers, shape 32x32b (thread i of the warp reads lane base+i, N columns)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.st: registers -> TMEM, shape 32x32b (thread i of the warp writes lane base+i, N consecutive columns)
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address = (lane << 16) | column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) {
  return base + (lane << 16) + col;
}

// ----------------------------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------------------------
// GELU (exact-erf form of the reference, Keras activation='gelu') in the GEMM epilogues.
// erf(x/sqrt2) is evaluated as tanh(x*(c1 + c3 x^2 + c5 x^4)) with minimax coefficients: max abs error
// 3.7e-5 against erf (fit over [0,6], see tests/test_host.py::test_gelu_approximation_constants), i.e. gelu is within
// 5.5e-5 and its derivative within 1.4e-4 of the erf form — far below the bf16 resolution of the stored results.
// One MUFU.TANH and ~8 FMA-pipe instructions per element: the libm erff (and an exp+rcp formulation) made the
// FFN epilogues issue-bound instead of HBM-bound.  |x| is clamped to 8 (erf(8/sqrt2) == 1 in fp32; the quintic
// turns over near |x| = 8.7).
__device__ __forceinline__ float tanh_approx(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return t;
}
static constexpr float kGeluC1 = 7.97717834e-01f, kGeluC3 = 3.67982560e-02f, kGeluC5 = -3.15807047e-04f;
// The argument is bounded by clamping t = x^2 at 64: there the odd polynomial equals 1.86 x, |u| > 14.9, tanh = +-1.
__device__ __forceinline__ float gelu_erf(float x) {
  const float x2 = fminf(x * x, 64.0f);
  const float t = tanh_approx(x * fmaf(fmaf(kGeluC5, x2, kGeluC3), x2, kGeluC1));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// d/dx gelu(x) = Phi(x) + x*phi(x), from the same approximation: Phi = (1+t)/2, phi = (1-t^2)/2 * u'(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float x2 = fminf(x * x, 64.0f);
  const float t = tanh_approx(x * fmaf(fmaf(kGeluC5, x2, kGeluC3), x2, kGeluC1));
  const float du = fmaf(fmaf(5.0f * kGeluC5, x2, 3.0f * kGeluC3), x2, kGeluC1);
  const float half_sech2 = fmaf(-0.5f * t, t, 0.5f);
  return fmaf(x * half_sech2, du, fmaf(0.5f, t, 0.5f));
}

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes) ----------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(f32x2 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// The two scalar steps of the packed GELU (there is no packed tanh / min): one asm block that works on the halves of the 64-bit
// register pair in place, so that ptxas does not shuffle the halves through IMAD.MOV to re-form an aligned pair (12 % of the
// instructions of the fused FFN kernel were such moves, profiles/README.md round 2).
__device__ __forceinline__ f32x2 tanh2(f32x2 u) {
  asm("{\n\t.reg .f32 a, b;\n\tmov.b64 {a, b}, %0;\n\ttanh.approx.f32 a, a;\n\ttanh.approx.f32 b, b;\n\tmov.b64 %0, {a, b};\n\t}" : "+l"(u));
  return u;
}
__device__ __forceinline__ f32x2 min2(f32x2 v, float m) {
  asm("{\n\t.reg .f32 a, b;\n\tmov.b64 {a, b}, %0;\n\tmin.f32 a, a, %1;\n\tmin.f32 b, b, %1;\n\tmov.b64 %0, {a, b};\n\t}" : "+l"(v) : "f"(m));
  return v;
}
// the two functions above on a pair: 5 / 8 issue slots per element instead of 9 / 14
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  const f32x2 x2 = min2(mul2(x, x), 64.0f);
  const f32x2 t = tanh2(mul2(x, fma2(fma2(pk2(kGeluC5), x2, pk2(kGeluC3)), x2, pk2(kGeluC1))));
  const f32x2 hx = mul2(x, pk2(0.5f));
  return fma2(hx, t, hx);
}
__device__ __forceinline__ f32x2 gelu_erf_grad2(f32x2 x) {
  const f32x2 x2 = min2(mul2(x, x), 64.0f);
  const f32x2 t = tanh2(mul2(x, fma2(fma2(pk2(kGeluC5), x2, pk2(kGeluC3)), x2, pk2(kGeluC1))));
  const f32x2 du = fma2(fma2(pk2(5.0f * kGeluC5), x2, pk2(3.0f * kGeluC3)), x2, pk2(kGeluC1));
  const f32x2 half_sech2 = fma2(mul2(t, pk2(-0.5f)), t, pk2(0.5f));
  return fma2(mul2(x, half_sech2), du, fma2(t, pk2(0.5f), pk2(0.5f)));
}

// 2^x on the MUFU pipe, one instruction (exp2f() adds range fix-ups that matter in the attention inner loops)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 magic add, degree-3 minimax polynomial
// for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, i.e. 1/50 of a bf16 half-ulp; same bf16 value as the exact exp2 for 99.1 % of
// inputs), exponent injected by an integer add.  8 FMA/ALU-pipe instructions against one instruction on the 16-lane XU pipe: meant
// for a FRACTION of the elements of the attention softmax loops, whose element-wise phases are XU-bound (profiles/README.md, known
// gap 1).  Arguments below -126 give 2^-126 (~1e-38, zero for a softmax).  Wired into the attention softmax loops behind the compile-time
// switch OT_EX2_POLY_MODE (ex2_mixed below; default 0 = MUFU only).  Round-2 measurement: the loops are issue-bound, not XU-bound
// (XU 29 %), so the extra FMA-pipe instructions make them SLOWER (forward 1.12 -> 1.64 ms with mode 1) - the switch stays off.  The
// constants and the arithmetic are pinned by a bit-level fp32 emulation in tests/test_host.py::test_polynomial_exp2_building_block.
#define OT_EX2_POLY_C0 0.9999280571937561f
#define OT_EX2_POLY_C1 0.6932609677314758f
#define OT_EX2_POLY_C2 0.2426111251115799f
#define OT_EX2_POLY_C3 0.0551716648042202f
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                 // low mantissa bits of t now hold round(x) in two's complement
  const float f = x - (t - 12582912.0f);           // in [-0.5, 0.5]
  float p = fmaf(f, OT_EX2_POLY_C3, OT_EX2_POLY_C2);
  p = fmaf(p, f, OT_EX2_POLY_C1);
  p = fmaf(p, f, OT_EX2_POLY_C0);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // p * 2^round(x)
}

// ex2_poly on a packed pair: 2 FMNMX + 6 packed FMA-pipe instructions + 2 integer ops for two results (the scalar form takes 8
// per result).  Same constants and arithmetic per lane as ex2_poly (pinned by tests/test_host.py).  Used by the attention forward
// v3 for a fraction of the exponentials: its two softmax warps per scheduler run their exponentials at the same time (lockstep
// tiles) and queue for the 16-lane XU pipe, while the FMA pipe is a quarter busy (profiles/README.md, round 2 second session).
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 x) {
  float x0, x1;
  upk2(x, x0, x1);
  x = pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const f32x2 t = add2(x, pk2(12582912.0f));
  const f32x2 f = fma2(add2(t, pk2(-12582912.0f)), pk2(-1.0f), x);     // x - round(x), in [-0.5, 0.5]
  f32x2 p = fma2(f, pk2(OT_EX2_POLY_C3), pk2(OT_EX2_POLY_C2));
  p = fma2(p, f, pk2(OT_EX2_POLY_C1));
  p = fma2(p, f, pk2(OT_EX2_POLY_C0));
  float p0, p1, t0, t1;
  upk2(p, p0, p1);
  upk2(t, t0, t1);
  return pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)), __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}

// The i-th exponential of an unrolled softmax loop.  OT_EX2_POLY_MODE (compile time, default 0 = every exponential on the MUFU pipe,
// the verified build): 1 sends every second one to ex2_poly, 2 every fourth.  An experiment switch for round 2
// (OT_NVCC_EXTRA="-DOT_EX2_POLY_MODE=1" recommend_b200/csrc/build.sh); `i` is a compile-time constant after unrolling.
#ifndef OT_EX2_POLY_MODE
#define OT_EX2_POLY_MODE 0
#endif
__device__ __forceinline__ float ex2_mixed(float x, int i) {
#if OT_EX2_POLY_MODE == 1
  return (i & 1) ? ex2_poly(x) : ex2_approx(x);
#elif OT_EX2_POLY_MODE == 2
  return ((i & 3) == 3) ? ex2_poly(x) : ex2_approx(x);
#else
  (void)i;
  return ex2_approx(x);
#endif
}

// Counter-based dropout mask (Keras inverted dropout, OT/model.py:184,193,198): one 32-bit hash decides the two
// elements (row, col) and (row, col+1), col even; an element is kept iff its 16-bit lane >= thr16 = round(rate*65536).
// The forward epilogue and the backward mask kernel evaluate the same function, so no mask is ever stored.
__device__ __forceinline__ uint32_t ot_hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t dropout_bits(uint32_t seed, uint32_t row, uint32_t col_even, uint32_t n_cols) {
  return ot_hash32(((row * n_cols + col_even) >> 1) * 2654435761u + seed);
}

// Byte offset of 16-byte chunk `chunk` of row `row` inside a SWB-byte-wide swizzled slab.
template <int SWB>
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t chunk) {
  if (SWB == 128) return row * 128u + ((chunk ^ (row & 7u)) << 4);
  return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}

}  // namespace ot
