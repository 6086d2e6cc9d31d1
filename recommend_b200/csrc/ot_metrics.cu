// ot_metrics.cu — streaming evaluation metrics of the OneTrans trainer / evaluator as HBM-streaming integer kernels
// (OT/train.py:95-109 and OT/evaluate.py:39-56 create, per binary task, Keras AUC / BinaryAccuracy / Precision / Recall
// (+ F1Score / BinaryCrossentropy in the evaluator); update_state calls at OT/train.py:141-150, 178-187 and
// OT/evaluate.py:91-99; result() at OT/train.py:248-249 and OT/evaluate.py:109-111).
//
// Keras 2.12 semantics restated (oracle/metrics_oracle.py cites the same rules):
//   AUC()              200 evenly spaced thresholds -> the bucketed update: b = max(ceil(clip(p, 0, 1) * (NT - 1)) - 1, 0) in
//                      fp32, per-bucket label sums, reverse cumulative sums = tp / fp at every threshold, ROC with
//                      'interpolation' (trapezoid) summation
//   BinaryAccuracy()   mean(y == (p > 0.5));  Precision() / Recall(): confusion counts at p > 0.5, div_no_nan
//   BinaryCrossentropy metric: mean of -(y log(p' + 1e-7) + (1 - y) log(1 - p' + 1e-7)), p' = clip(p, 1e-7, 1 - 1e-7)
// All counts are integers (int64, exact where Keras' fp32 accumulators stop being exact at 2^24 samples per bucket).
//
// Also here: the exact, tie-aware ROC-AUC used for the "AUC delta <= 1e-4" criterion and the per-user
// AUC: keys (segment, order-preserving bits of p, label) are packed by ot_auc_pack_keys, sorted by the caller, and
// ot_auc_ranksum accumulates twice the mid-rank of every positive (integers, bit-exact).
//
// Traffic: update 8 B / (sample, task); pack 8-12 B read + 8 B written; ranksum 8 B read per sample.
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int MET_THREADS = 256;
static constexpr int MET_WARPS = MET_THREADS / 32;

__device__ __forceinline__ void atomic_add_i64(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

// block-wide sums of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum_u32(unsigned (&v)[NV], unsigned* red /* [MET_WARPS * NV] */) {
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) red[(threadIdx.x >> 5) * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      unsigned t = 0;
      for (int w = 0; w < MET_WARPS; ++w) t += red[w * NV + k];
      v[k] = t;
    }
}

// one sample into the per-warp histogram and the per-thread counters
__device__ __forceinline__ void metrics_sample(float pi, float yi, unsigned* __restrict__ my, int NT, float scale, float threshold,
                                               unsigned (&c)[6], float& bce) {
  if (!(pi == pi) || !(yi == 0.0f || yi == 1.0f)) { c[5]++; return; }       // NaN prediction or non-binary label
  const bool pos = yi != 0.0f;
  const float pc = fminf(fmaxf(pi, 0.0f), 1.0f);
  int b = (int)ceilf(pc * scale) - 1;
  b = b < 0 ? 0 : b;
  atomicAdd(&my[(pos ? 0 : NT) + b], 1u);
  const bool pp = pi > threshold;
  c[0] += (pos && pp); c[1] += (!pos && pp); c[2] += (!pos && !pp); c[3] += (pos && !pp); c[4]++;
  const float q = fminf(fmaxf(pi, 1e-7f), 1.0f - 1e-7f);
  bce -= logf((pos ? q : 1.0f - q) + 1e-7f);      // == -(y log(q + eps) + (1 - y) log(1 - q + eps)) for y in {0, 1}
}

// grid (x = sample chunks, y = task).  Per-warp private histograms in shared memory (predictions of a trained model pile up
// in a few buckets; eight copies keep the shared atomics apart), one int64 atomic per non-empty bucket and CTA at the end.
// Loads are 16-byte and two deep per thread (32 B of predictions + 32 B of labels in flight per thread): with one 4-byte
// load per thread the kernel sat at 23 % of the HBM peak, stalled on the long scoreboard (profiles/README.md).
__global__ void __launch_bounds__(MET_THREADS)
metrics_update_kernel(const float* __restrict__ probs, const float* __restrict__ labels, long long ld, long long B, int NT,
                      float threshold, long long* __restrict__ state, long long state_stride) {
  extern __shared__ unsigned hist[];                  // [MET_WARPS][2][NT]
  __shared__ unsigned red[MET_WARPS * 6];
  __shared__ double red_d[MET_WARPS];
  const int task = blockIdx.y;
  const float* p = probs + (long long)task * ld;
  const float* y = labels + (long long)task * ld;
  long long* st = state + (long long)task * state_stride;
  for (int i = threadIdx.x; i < MET_WARPS * 2 * NT; i += MET_THREADS) hist[i] = 0u;
  __syncthreads();
  unsigned* my = hist + (threadIdx.x >> 5) * 2 * NT;
  const float scale = (float)(NT - 1);
  unsigned c[6] = {0u, 0u, 0u, 0u, 0u, 0u};           // tp, fp, tn, fn, count, rejected
  float bce = 0.0f;
  // each CTA walks a contiguous run of samples (a multiple of 4, so that runs start 16-byte aligned when the rows do)
  const long long per = (((B + gridDim.x - 1) / gridDim.x) + 3) & ~3ll;
  const long long i0 = (long long)blockIdx.x * per;
  const long long i1 = (i0 + per < B) ? i0 + per : B;
  long long done = i0;
  if (i0 < i1 && ((reinterpret_cast<uintptr_t>(p + i0) | reinterpret_cast<uintptr_t>(y + i0)) & 15) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(p + i0);
    const float4* y4 = reinterpret_cast<const float4*>(y + i0);
    const long long n4 = (i1 - i0) >> 2;
    for (long long j = threadIdx.x; j < n4; j += 2 * MET_THREADS) {
      const bool two = j + MET_THREADS < n4;
      const float4 pa = __ldg(p4 + j), ya = __ldg(y4 + j);
      float4 pb = make_float4(0.f, 0.f, 0.f, 0.f), yb = pb;
      if (two) { pb = __ldg(p4 + j + MET_THREADS); yb = __ldg(y4 + j + MET_THREADS); }
      metrics_sample(pa.x, ya.x, my, NT, scale, threshold, c, bce);
      metrics_sample(pa.y, ya.y, my, NT, scale, threshold, c, bce);
      metrics_sample(pa.z, ya.z, my, NT, scale, threshold, c, bce);
      metrics_sample(pa.w, ya.w, my, NT, scale, threshold, c, bce);
      if (two) {
        metrics_sample(pb.x, yb.x, my, NT, scale, threshold, c, bce);
        metrics_sample(pb.y, yb.y, my, NT, scale, threshold, c, bce);
        metrics_sample(pb.z, yb.z, my, NT, scale, threshold, c, bce);
        metrics_sample(pb.w, yb.w, my, NT, scale, threshold, c, bce);
      }
    }
    done = i0 + (n4 << 2);
  }
  for (long long i = done + threadIdx.x; i < i1; i += MET_THREADS)
    metrics_sample(__ldg(p + i), __ldg(y + i), my, NT, scale, threshold, c, bce);
  __syncthreads();
  for (int b = threadIdx.x; b < 2 * NT; b += MET_THREADS) {
    unsigned t = 0;
#pragma unroll
    for (int w = 0; w < MET_WARPS; ++w) t += hist[w * 2 * NT + b];
    if (t) atomic_add_i64(st + b, (long long)t);
  }
  block_sum_u32<6>(c, red);
  double bd = (double)bce;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bd += __shfl_xor_sync(0xffffffffu, bd, o);
  if ((threadIdx.x & 31) == 0) red_d[threadIdx.x >> 5] = bd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < MET_WARPS; ++w) t += red_d[w];
    long long* tail = st + 2 * NT;
#pragma unroll
    for (int k = 0; k < 6; ++k) if (c[k]) atomic_add_i64(tail + k, (long long)c[k]);
    if (c[4]) atomicAdd(reinterpret_cast<double*>(tail + 6), t);
  }
}

__device__ __forceinline__ float div_no_nan(float a, float b) { return b == 0.0f ? 0.0f : a / b; }

// one CTA per task; the table is NT <= 512 entries, thread 0 walks it (fp32 element-wise arithmetic as Keras' result())
__global__ void __launch_bounds__(32)
metrics_result_kernel(const long long* __restrict__ state, long long state_stride, int NT, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  const long long* st = state + (long long)blockIdx.x * state_stride;
  const long long* tail = st + 2 * NT;
  double* o = out + (long long)blockIdx.x * OT_METRICS_RESULT_WORDS;
  long long P = 0, N = 0;
  for (int b = 0; b < NT; ++b) { P += st[b]; N += st[NT + b]; }
  // AUC: threshold i counts the buckets >= i  (reverse cumulative sum);  x = fp rate, y = recall
  long long tp = P, fp = N;
  float x_prev = div_no_nan((float)fp, (float)fp + (float)(N - fp)), y_prev = div_no_nan((float)tp, (float)tp + (float)(P - tp));
  float auc = 0.0f;
  for (int i = 1; i < NT; ++i) {
    tp -= st[i - 1]; fp -= st[NT + i - 1];
    const float x = div_no_nan((float)fp, (float)fp + (float)(N - fp)), y = div_no_nan((float)tp, (float)tp + (float)(P - tp));
    auc += (x_prev - x) * ((y_prev + y) / 2.0f);
    x_prev = x; y_prev = y;
  }
  const float ctp = (float)tail[0], cfp = (float)tail[1], cfn = (float)tail[3];
  const double cnt = (double)tail[4];
  const float prec = div_no_nan(ctp, ctp + cfp), rec = div_no_nan(ctp, ctp + cfn);
  o[0] = auc;
  o[1] = cnt > 0 ? (double)(tail[0] + tail[2]) / cnt : 0.0;
  o[2] = prec;
  o[3] = rec;
  o[4] = div_no_nan(2.0f * prec * rec, prec + rec);
  o[5] = cnt > 0 ? *reinterpret_cast<const double*>(tail + 6) / cnt : 0.0;
  o[6] = cnt;
  o[7] = (double)tail[5];
}

int metrics_update_impl(const ot_metrics_params* p, cudaStream_t st) {
  if (!p || !p->probs || !p->labels || !p->state) OT_FAIL(OT_ERR_INVALID_ARG, "ot_metrics_update: null pointer");
  if (p->num_thresholds < 3 || p->num_thresholds > OT_METRICS_MAX_THRESHOLDS)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_metrics_update: num_thresholds=%d outside [3, %d]", p->num_thresholds, OT_METRICS_MAX_THRESHOLDS);
  if (p->n_tasks <= 0 || p->n_tasks > 65535 || p->B < 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_metrics_update: n_tasks=%d B=%lld", p->n_tasks, (long long)p->B);
  if (p->state_stride < 2 * p->num_thresholds + OT_METRICS_TAIL_WORDS) OT_FAIL(OT_ERR_INVALID_ARG, "ot_metrics_update: state_stride=%lld too small", (long long)p->state_stride);
  if (p->B == 0) return OT_OK;
  const long long want = (p->B + MET_THREADS * 16 - 1) / (MET_THREADS * 16);
  long long cap = (long long)num_sms() * 8 / p->n_tasks;
  if (cap < 1) cap = 1;
  long long gx = want < cap ? want : cap;
  const long long min_gx = (p->B + (1ll << 30) - 1) >> 30;          // u32 per-CTA counters
  if (gx < min_gx) gx = min_gx;
  const size_t smem = sizeof(unsigned) * MET_WARPS * 2 * p->num_thresholds;
  metrics_update_kernel<<<dim3((unsigned)gx, (unsigned)p->n_tasks), MET_THREADS, smem, st>>>(
      p->probs, p->labels, p->ld, p->B, p->num_thresholds, p->threshold, (long long*)p->state, p->state_stride);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int metrics_result_impl(const ot_metrics_params* p, cudaStream_t st) {
  if (!p || !p->state || !p->result) OT_FAIL(OT_ERR_INVALID_ARG, "ot_metrics_result: null pointer");
  if (p->num_thresholds < 3 || p->num_thresholds > OT_METRICS_MAX_THRESHOLDS || p->n_tasks <= 0)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "ot_metrics_result: num_thresholds=%d n_tasks=%d", p->num_thresholds, p->n_tasks);
  metrics_result_kernel<<<p->n_tasks, 32, 0, st>>>((const long long*)p->state, p->state_stride, p->num_thresholds, p->result);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

// ---- exact ROC-AUC / per-segment AUC ---------------------------------------------------------------------------------
static constexpr long long AUC_REJECTED_KEY = 0x7fffffffffffffffll;          // segment field 2^30 - 1: n_segments stays below it
// key = segment << 33 | ordered(p) << 1 | label, ordered(p) = the usual order-preserving map of fp32 onto u32
__global__ void __launch_bounds__(256)
auc_pack_kernel(const float* __restrict__ probs, const float* __restrict__ labels, const int* __restrict__ seg, long long n,
                int n_segments, long long* __restrict__ keys, int* __restrict__ rejected) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = __ldg(probs + i) + 0.0f, y = __ldg(labels + i);           // + 0.0f: -0 -> +0
    const int s = seg ? __ldg(seg + i) : 0;
    if (!(p == p) || !(y == 0.0f || y == 1.0f) || s < 0 || s >= n_segments) {
      if (rejected) atomicAdd(rejected, 1);
      keys[i] = AUC_REJECTED_KEY;                                             // sorts last, skipped by the rank-sum pass
      continue;
    }
    unsigned b = __float_as_uint(p);
    b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;
    keys[i] = ((long long)s << 33) | ((long long)b << 1) | (y != 0.0f ? 1ll : 0ll);
  }
}

// sorted keys -> per segment: cnt, pos, sum2 = sum over positives of (first + last + 1) of their tie group, i.e. twice the
// 1-based mid-rank in the GLOBAL order (the host subtracts the segment start).  A CTA covers AUC_ITEMS * 256 consecutive keys
// and, when they all belong to one segment (the common case), leaves with one atomic triple: with one triple per 256 keys the
// same-address atomics of the single-segment case serialised the kernel at 185 GB/s (profiles/README.md).
static constexpr int AUC_ITEMS = 8;
__global__ void __launch_bounds__(256)
auc_ranksum_kernel(const long long* __restrict__ keys, long long n, long long* __restrict__ seg_cnt, long long* __restrict__ seg_pos,
                   long long* __restrict__ seg_sum2) {
  __shared__ long long red[3][8];
  const long long base = (long long)blockIdx.x * (256 * AUC_ITEMS);
  const long long last = (base + 256 * AUC_ITEMS <= n ? base + 256 * AUC_ITEMS : n) - 1;
  const bool uniform = (keys[base] >> 33) == (keys[last] >> 33);      // base < n by the grid size
  long long cnt = 0, pos = 0, sum2 = 0;
#pragma unroll 2
  for (int it = 0; it < AUC_ITEMS; ++it) {
    const long long i = base + it * 256 + threadIdx.x;
    if (i >= n) break;
    const long long k = keys[i];
    if (k == AUC_REJECTED_KEY) continue;
    const long long v = k >> 1;                     // (segment, value): the tie class
    long long s2 = 0;
    if (k & 1) {
      long long s = i, e = i + 1;                   // tie group [s, e)
      // tie groups are short: gallop outwards from i (1, 2, 4 ... keys), then bisect the last bracket
      if (i > 0 && (keys[i - 1] >> 1) == v) {
        long long hi = i - 1, step = 1, below = -1;                 // keys[hi] == v; keys[below] < v (or below == -1)
        for (;;) {
          const long long probe = hi - step;
          if (probe < 0) break;
          if ((keys[probe] >> 1) == v) { hi = probe; step <<= 1; } else { below = probe; break; }
        }
        long long lo = below + 1;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((keys[mid] >> 1) < v) lo = mid + 1; else hi = mid; }
        s = lo;
      }
      if (i + 1 < n && (keys[i + 1] >> 1) == v) {
        long long lo = i + 1, step = 1, above = n;                  // keys[lo] == v; keys[above] > v (or above == n)
        for (;;) {
          const long long probe = lo + step;
          if (probe >= n) break;
          if ((keys[probe] >> 1) == v) { lo = probe; step <<= 1; } else { above = probe; break; }
        }
        long long a = lo + 1, b = above;                            // first index with a larger value lies in [a, b]
        while (a < b) { const long long mid = (a + b) >> 1; if ((keys[mid] >> 1) <= v) a = mid + 1; else b = mid; }
        e = a;
      }
      s2 = s + e + 1;
    }
    if (uniform) {
      cnt += 1; pos += (k & 1); sum2 += s2;
    } else {
      const int seg = (int)(k >> 33);
      atomic_add_i64(seg_cnt + seg, 1);
      if (k & 1) { atomic_add_i64(seg_pos + seg, 1); atomic_add_i64(seg_sum2 + seg, s2); }
    }
  }
  if (uniform) {                                     // CTA-uniform branch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o); pos += __shfl_xor_sync(0xffffffffu, pos, o); sum2 += __shfl_xor_sync(0xffffffffu, sum2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cnt; red[1][threadIdx.x >> 5] = pos; red[2][threadIdx.x >> 5] = sum2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long a = 0, b = 0, c = 0;
      for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
      const int sg = (int)(keys[base] >> 33);
      if (a) atomic_add_i64(seg_cnt + sg, a);        // a == 0: a CTA of rejected keys only
      if (b) { atomic_add_i64(seg_pos + sg, b); atomic_add_i64(seg_sum2 + sg, c); }
    }
  }
}

int auc_pack_impl(const ot_auc_params* p, cudaStream_t st) {
  if (!p || !p->probs || !p->labels || !p->keys) OT_FAIL(OT_ERR_INVALID_ARG, "ot_auc_pack_keys: null pointer");
  if (p->n < 0 || p->n_segments <= 0 || p->n_segments >= (1 << 30) - 1) OT_FAIL(OT_ERR_INVALID_ARG, "ot_auc_pack_keys: n=%lld n_segments=%d", (long long)p->n, p->n_segments);
  if (p->n == 0) return OT_OK;
  const long long want = (p->n + 255) / 256, cap = (long long)num_sms() * 8;
  auc_pack_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(p->probs, p->labels, p->segment_ids, p->n, p->n_segments,
                                                                      (long long*)p->keys, p->rejected);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int auc_ranksum_impl(const ot_auc_params* p, cudaStream_t st) {
  if (!p || !p->keys || !p->seg_count || !p->seg_pos || !p->seg_sum2) OT_FAIL(OT_ERR_INVALID_ARG, "ot_auc_ranksum: null pointer");
  if (p->n < 0 || p->n > (1ll << 38)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_auc_ranksum: n=%lld", (long long)p->n);
  if (p->n == 0) return OT_OK;
  auc_ranksum_kernel<<<(unsigned)((p->n + 256 * AUC_ITEMS - 1) / (256 * AUC_ITEMS)), 256, 0, st>>>((const long long*)p->keys, p->n, (long long*)p->seg_count,
                                                                    (long long*)p->seg_pos, (long long*)p->seg_sum2);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
