// ot_optimizer.cu — the parameter update of the OneTrans train step as two HBM-streaming kernels
// (OT/train.py:131-138: per-tensor tf.clip_by_norm, then RMSprop.apply_gradients; optimizer built at
// OT/train.py:65-70 with rho 0.9 / epsilon 1e-7 defaults and lr / momentum from OT/config.py:39-52).
//
// Layout: gradients and optimizer state live in flat fp32 buffers in which every tensor starts on a
// 1024-element boundary (OT_OPT_CHUNK), so a 1024-element chunk never straddles two tensors; the fp32
// masters stay where the framework allocated them and are reached through a pointer table.
//   pass 1  sqnorm[v] = sum(g_v^2) per Keras variable v (clip slot)      4 B / parameter
//   pass 2  g' = g * clip/max(||g||, clip);  rms = rho rms + (1-rho) g'^2;
//           inc = lr g' rsqrt(rms + eps);  mom = momentum mom + inc;  w -= mom      28-32 B / parameter
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int OPT_CHUNK = 1024;   // elements per chunk == 256 threads x float4
static_assert(OPT_CHUNK == OT_OPT_CHUNK, "header / kernel chunk mismatch");

__device__ __forceinline__ int seg_of(const long long* __restrict__ seg_off, int n_seg, long long idx) {
  int lo = 0, hi = n_seg - 1;              // last s with seg_off[s] <= idx
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_off[mid] <= idx) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Clip slot of element `local` of tensor `seg` (see ot_rmsprop_params.seg_slot): which Keras variable the element belongs to.
__device__ __forceinline__ int slot_of(const long long* __restrict__ seg_slot, int seg, long long local) {
  if (seg_slot == nullptr) return seg;
  const long long base = seg_slot[4 * seg], outer = seg_slot[4 * seg + 1], row = seg_slot[4 * seg + 2], part = seg_slot[4 * seg + 3];
  return (int)(base + (local / outer) * (row / part) + (local % row) / part);
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                          // red[] may still be read by the previous flush
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;                                 // valid in thread 0
}

// Each CTA owns a contiguous run of chunks and flushes one atomic per tensor it touches.
__global__ void __launch_bounds__(256)
opt_sqnorm_kernel(const float* __restrict__ grad, const long long* __restrict__ seg_off, int n_seg, long long n_chunks,
                  float grad_scale, float* __restrict__ sqnorm) {
  __shared__ float red[8];
  const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const long long c0 = (long long)blockIdx.x * per;
  const long long c1 = (c0 + per < n_chunks) ? c0 + per : n_chunks;
  if (c0 >= c1) return;
  int seg = seg_of(seg_off, n_seg, c0 * OPT_CHUNK);
  long long seg_end = seg_off[seg + 1];
  float acc = 0.0f;
  for (long long c = c0; c < c1; ++c) {
    const long long base = c * OPT_CHUNK;
    if (base >= seg_end) {                  // block-uniform: next tensor starts here
      const float t = block_sum_256(acc, red);
      if (threadIdx.x == 0) atomicAdd(&sqnorm[seg], t);
      acc = 0.0f;
      seg = seg_of(seg_off, n_seg, base);
      seg_end = seg_off[seg + 1];
    }
    const float4 g = __ldg(reinterpret_cast<const float4*>(grad + base) + threadIdx.x);
    const float a = g.x * grad_scale, b = g.y * grad_scale, cc = g.z * grad_scale, d = g.w * grad_scale;
    acc += a * a + b * b + cc * cc + d * d;
  }
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(&sqnorm[seg], t);
}

// Same pass with clip slots finer than tensors (one per Keras variable): a warp covers 128 consecutive elements, which lie
// in one slot whenever slot_part % 128 == 0 (every production shape: d, F multiples of 128) - then the warp keeps a running
// (slot, sum) pair over the CTA's chunks and flushes one atomic when the slot changes; otherwise every lane adds its own.
// Alignment padding holds zeros (the flat buffer is zero-initialised and kernels only write real elements).
__global__ void __launch_bounds__(256)
opt_sqnorm_slots_kernel(const float* __restrict__ grad, const long long* __restrict__ seg_off, const long long* __restrict__ seg_numel,
                        const long long* __restrict__ seg_slot, int n_seg, long long n_chunks, float grad_scale,
                        float* __restrict__ sqnorm) {
  const long long per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const long long c0 = (long long)blockIdx.x * per;
  const long long c1 = (c0 + per < n_chunks) ? c0 + per : n_chunks;
  if (c0 >= c1) return;
  int seg = seg_of(seg_off, n_seg, c0 * OPT_CHUNK);
  long long seg_end = seg_off[seg + 1];
  int run_slot = -1;
  float run = 0.0f;
  for (long long c = c0; c < c1; ++c) {
    const long long base = c * OPT_CHUNK;
    if (base >= seg_end) { seg = seg_of(seg_off, n_seg, base); seg_end = seg_off[seg + 1]; }
    const long long local = base - seg_off[seg] + threadIdx.x * 4;
    const bool real = local < seg_numel[seg];
    float v = 0.0f;
    int slot = -1;
    if (real) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(grad + base) + threadIdx.x);
      const float a = g.x * grad_scale, b = g.y * grad_scale, cc = g.z * grad_scale, d = g.w * grad_scale;
      v = a * a + b * b + cc * cc + d * d;
      slot = slot_of(seg_slot, seg, local);
    }
    const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
    if (__all_sync(0xffffffffu, slot == slot0)) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (slot0 != run_slot) {
        if ((threadIdx.x & 31) == 0 && run_slot >= 0) atomicAdd(&sqnorm[run_slot], run);
        run_slot = slot0;
        run = 0.0f;
      }
      run += v;
    } else if (real) {
      atomicAdd(&sqnorm[slot], v);
    }
  }
  if ((threadIdx.x & 31) == 0 && run_slot >= 0) atomicAdd(&sqnorm[run_slot], run);
}

template <bool MOMENTUM>
__global__ void __launch_bounds__(256)
opt_rmsprop_kernel(float* const* __restrict__ param_ptrs, const long long* __restrict__ seg_off,
                   const long long* __restrict__ seg_numel, int n_seg, long long n_chunks, float* __restrict__ grad,
                   float* __restrict__ rms, float* __restrict__ mom, const float* __restrict__ sqnorm, float lr, float rho,
                   float momentum, float eps, float clip_norm, float grad_scale, int zero_grad,
                   const long long* __restrict__ seg_slot) {
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const long long base = c * OPT_CHUNK;
    const int seg = seg_of(seg_off, n_seg, base);
    const long long local = base - seg_off[seg] + threadIdx.x * 4;       // element index inside the tensor
    const long long numel = seg_numel[seg];
    if (local >= numel) continue;                                        // alignment padding
    float scale = grad_scale;
    // tf.clip_by_norm per Keras variable; the four elements of a thread share a slot (slot_part % 4 == 0)
    if (clip_norm > 0.0f) scale *= clip_norm / fmaxf(sqrtf(sqnorm[slot_of(seg_slot, seg, local)]), clip_norm);
    float* w = param_ptrs[seg] + local;
    const long long fi = base + threadIdx.x * 4;
    const int n = (numel - local >= 4) ? 4 : (int)(numel - local);
    float g[4], r[4], m[4], p[4];
    if (n == 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(grad + fi), r4 = *reinterpret_cast<const float4*>(rms + fi);
      const float4 p4 = *reinterpret_cast<const float4*>(w);
      g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w; r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
      p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
      if (MOMENTUM) { const float4 m4 = *reinterpret_cast<const float4*>(mom + fi); m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w; }
    } else {
      for (int e = 0; e < n; ++e) { g[e] = grad[fi + e]; r[e] = rms[fi + e]; p[e] = w[e]; if (MOMENTUM) m[e] = mom[fi + e]; }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < n) {
        const float ge = g[e] * scale;
        r[e] = rho * r[e] + (1.0f - rho) * ge * ge;
        const float inc = lr * ge * rsqrtf(r[e] + eps);
        if (MOMENTUM) { m[e] = momentum * m[e] + inc; p[e] -= m[e]; } else { p[e] -= inc; }
      }
    }
    if (n == 4) {
      *reinterpret_cast<float4*>(rms + fi) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(w) = make_float4(p[0], p[1], p[2], p[3]);
      if (MOMENTUM) *reinterpret_cast<float4*>(mom + fi) = make_float4(m[0], m[1], m[2], m[3]);
      if (zero_grad) *reinterpret_cast<float4*>(grad + fi) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int e = 0; e < n; ++e) { rms[fi + e] = r[e]; w[e] = p[e]; if (MOMENTUM) mom[fi + e] = m[e]; if (zero_grad) grad[fi + e] = 0.0f; }
    }
  }
}

int clip_rmsprop_impl(const ot_rmsprop_params* p, cudaStream_t st) {
  if (!p || !p->param_ptrs || !p->seg_off || !p->seg_numel || !p->grad || !p->rms || !p->sqnorm)
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_clip_rmsprop_step: null pointer");
  if (p->momentum != 0.0f && !p->mom) OT_FAIL(OT_ERR_INVALID_ARG, "ot_clip_rmsprop_step: momentum %f needs a mom buffer", (double)p->momentum);
  if (p->n_seg <= 0 || p->n_flat <= 0) return OT_OK;
  if (p->n_flat % OPT_CHUNK) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_clip_rmsprop_step: n_flat=%lld is not a multiple of %d", (long long)p->n_flat, OPT_CHUNK);
  if (!(p->rho >= 0.0f && p->rho <= 1.0f) || !(p->eps >= 0.0f)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_clip_rmsprop_step: rho=%f eps=%g", (double)p->rho, (double)p->eps);
  const long long n_chunks = p->n_flat / OPT_CHUNK;
  const long long cap = (long long)num_sms() * 8;
  const int grid = (int)(n_chunks < cap ? n_chunks : cap);
  const long long* seg_slot = (const long long*)p->seg_slot;
  if (seg_slot != nullptr && p->n_slots <= 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_clip_rmsprop_step: seg_slot without n_slots");
  if (p->clip_norm > 0.0f) {
    OT_CUDA_CHECK(cudaMemsetAsync(p->sqnorm, 0, sizeof(float) * (seg_slot ? p->n_slots : p->n_seg), st));
    if (seg_slot)
      opt_sqnorm_slots_kernel<<<grid, 256, 0, st>>>(p->grad, (const long long*)p->seg_off, (const long long*)p->seg_numel, seg_slot, p->n_seg,
                                                    n_chunks, p->grad_scale, p->sqnorm);
    else
      opt_sqnorm_kernel<<<grid, 256, 0, st>>>(p->grad, (const long long*)p->seg_off, p->n_seg, n_chunks, p->grad_scale, p->sqnorm);
    OT_CUDA_CHECK(cudaGetLastError());
  }
  if (p->momentum != 0.0f)
    opt_rmsprop_kernel<true><<<grid, 256, 0, st>>>(p->param_ptrs, (const long long*)p->seg_off, (const long long*)p->seg_numel, p->n_seg, n_chunks,
                                                    p->grad, p->rms, p->mom, p->sqnorm, p->lr, p->rho, p->momentum, p->eps, p->clip_norm,
                                                    p->grad_scale, p->zero_grad, seg_slot);
  else
    opt_rmsprop_kernel<false><<<grid, 256, 0, st>>>(p->param_ptrs, (const long long*)p->seg_off, (const long long*)p->seg_numel, p->n_seg, n_chunks,
                                                     p->grad, p->rms, nullptr, p->sqnorm, p->lr, p->rho, 0.0f, p->eps, p->clip_norm,
                                                     p->grad_scale, p->zero_grad, seg_slot);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
