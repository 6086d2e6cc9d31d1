// ot_heads.cu — output RMSNorm + task heads (+ the BCE loss) of the OneTrans model on the last token, in fp32
// (OT/model.py:322-330 heads = Dense(d/2, gelu) -> Dense(1, sigmoid) per task, :384-391 output norm on x[:, -1, :];
// loss OT/train.py:84-87, 124-128: Keras BinaryCrossentropy(from_logits=False), mean over the batch, summed over tasks).
//
// The work is tiny ([B, d] rows, 2 x 67 MFLOP at C2) and precision-critical (it produces the logits the parity bar is
// stated on), so it runs on the fp32 CUDA cores: tensor cores are for the QKV / FFN / attention contractions.  What
// matters here is launch count: these three kernels replace ~110 framework launches per training step.
//   ot_heads_fwd : x -> xn = RMSNorm(x) -> per task  pre = xn W0 + b0, h = gelu(pre), logit = h.w1 + b1, prob = sigmoid
//                  (+ with labels: loss += sum_t mean_b BCE, and g_bce = d loss / d logit)
//   ot_heads_bwd : dlogit -> dW1, db1, db0, dpre (kept for the dW0 pass), dxn -> norm backward -> dx, dgain
//   (second launch inside ot_heads_bwd) dW0[k, j] += sum_b xn[b, k] dpre[b, j], one CTA per 8 rows of W0 and task
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

static constexpr int HD_SPB = 16;        // samples per CTA
static constexpr int HD_THREADS = 256;
static constexpr int HD_MAX_D = 512, HD_MAX_H = 256;

struct HeadsKParams {
  const float* x; long long ldx;
  const float* gain; float eps;
  int B, d, Hd, T;
  const float* W0[OT_MAX_TASKS]; const float* b0[OT_MAX_TASKS]; const float* W1[OT_MAX_TASKS]; const float* b1[OT_MAX_TASKS];
  float* xn; float* rstd; float* pre;      // saved: [B, d], [B], [T, B, Hd]
  float* logits; float* probs;             // [T, B]
  const float* labels; float* loss; float* g_bce;   // optional: [T, B], scalar (accumulated), [T, B]
  // backward
  const float* dlogit;                     // [T, B]
  float* dpre;                             // [T, B, Hd] workspace
  float* dW0[OT_MAX_TASKS]; float* db0[OT_MAX_TASKS]; float* dW1[OT_MAX_TASKS]; float* db1[OT_MAX_TASKS];
  float* dgain; float* dx; long long lddx;
};

__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float v) {
  return 0.5f * (1.0f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v);
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(HD_THREADS)
heads_fwd_kernel(const __grid_constant__ HeadsKParams p) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [SPB][d]   normalised rows
  float* hs = sm + HD_SPB * p.d;           // [SPB][Hd]  gelu outputs of the current task
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * HD_SPB;
  // ---- output norm (OT/model.py:384, RMSNorm.call :19-23): warp w takes samples 2w, 2w+1 ----
  for (int s = 2 * warp; s < 2 * warp + 2; ++s) {
    const int b = b0 + s;
    float ss = 0.0f;
    if (b < p.B)
      for (int k = lane; k < p.d; k += 32) { const float v = p.x[(long long)b * p.ldx + k]; ss += v * v; }
    ss = warp_sum_f(ss);
    const float r = rsqrtf(ss / (float)p.d + p.eps);
    for (int k = lane; k < p.d; k += 32) {
      float v = 0.0f;
      if (b < p.B) {
        v = p.x[(long long)b * p.ldx + k] * r * p.gain[k];
        p.xn[(long long)b * p.d + k] = v;
      }
      xs[s * p.d + k] = v;
    }
    if (b < p.B && lane == 0) p.rstd[b] = r;
  }
  __syncthreads();
  const int ng = HD_THREADS / p.Hd;        // thread groups over the hidden columns (1 or 2 ...)
  const int spg = HD_SPB / ng;             // samples per group (ng divides 16: Hd in {16, 32, 64, 128, 256} or ng == 1)
  const int g = tid / p.Hd, j = tid - g * p.Hd;
  for (int t = 0; t < p.T; ++t) {
    if (g < ng) {
      float acc[HD_SPB];
#pragma unroll
      for (int s = 0; s < HD_SPB; ++s) acc[s] = 0.0f;
      const float* w = p.W0[t] + j;
#pragma unroll 8
      for (int k = 0; k < p.d; ++k) {        // d % 8 == 0: eight independent loads in flight per thread
        const float wk = __ldg(w + (long long)k * p.Hd);
#pragma unroll
        for (int s = 0; s < HD_SPB; ++s)
          if (s < spg) acc[s] = fmaf(xs[(g * spg + s) * p.d + k], wk, acc[s]);
      }
      const float bj = p.b0[t][j];
#pragma unroll
      for (int s = 0; s < HD_SPB; ++s) {
        if (s < spg) {
          const int sl = g * spg + s, b = b0 + sl;
          const float pre = acc[s] + bj;
          if (b < p.B) p.pre[((long long)t * p.B + b) * p.Hd + j] = pre;
          hs[sl * p.Hd + j] = gelu_exact(pre);
        }
      }
    }
    __syncthreads();
    // ---- Dense(1, sigmoid) (OT/model.py:329) and the loss (OT/train.py:84-87) ----
    for (int s = 2 * warp; s < 2 * warp + 2; ++s) {
      const int b = b0 + s;
      float a = 0.0f;
      for (int jj = lane; jj < p.Hd; jj += 32) a += hs[s * p.Hd + jj] * p.W1[t][jj];
      a = warp_sum_f(a);
      if (lane == 0 && b < p.B) {
        const float logit = a + p.b1[t][0];
        const float prob = 1.0f / (1.0f + __expf(-logit));
        p.logits[(long long)t * p.B + b] = logit;
        p.probs[(long long)t * p.B + b] = prob;
        if (p.labels != nullptr) {
          const float y = p.labels[(long long)t * p.B + b];
          const float e = 1e-7f;
          const float pc = fminf(fmaxf(prob, e), 1.0f - e);
          const float l = -(y * logf(pc + e) + (1.0f - y) * logf(1.0f - pc + e));
          atomicAdd(p.loss, l / (float)p.B);
          const bool inside = prob > e && prob < 1.0f - e;            // clip passes no gradient outside
          const float dl_dp = inside ? -(y / (pc + e) - (1.0f - y) / (1.0f - pc + e)) : 0.0f;
          p.g_bce[(long long)t * p.B + b] = dl_dp * prob * (1.0f - prob) / (float)p.B;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(HD_THREADS)
heads_bwd_kernel(const __grid_constant__ HeadsKParams p) {
  extern __shared__ float sm[];
  float* xs = sm;                            // [SPB][d]   xh = x * rstd
  float* dxn = xs + HD_SPB * p.d;            // [SPB][d]   gradient w.r.t. the normalised rows, summed over tasks
  float* dps = dxn + HD_SPB * p.d;           // [SPB][Hd]  dpre of the current task
  float* dls = dps + HD_SPB * p.Hd;          // [SPB]      dlogit of the current task
  float* rs = dls + HD_SPB;                  // [SPB]      rstd
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * HD_SPB;
  for (int i = tid; i < HD_SPB * p.d; i += HD_THREADS) {
    const int s = i / p.d, k = i - s * p.d, b = b0 + s;
    const float r = b < p.B ? p.rstd[b] : 0.0f;
    xs[i] = b < p.B ? p.x[(long long)b * p.ldx + k] * r : 0.0f;
    dxn[i] = 0.0f;
    if (k == 0) rs[s] = r;
  }
  __syncthreads();
  const int ng = HD_THREADS / p.Hd, spg = HD_SPB / ng;
  const int g = tid / p.Hd, j = tid - g * p.Hd;
  for (int t = 0; t < p.T; ++t) {
    if (tid < HD_SPB) dls[tid] = (b0 + tid < p.B) ? p.dlogit[(long long)t * p.B + b0 + tid] : 0.0f;
    __syncthreads();
    if (g < ng) {
      const float w1 = p.W1[t][j];
      float a_w1 = 0.0f, a_b0 = 0.0f;
      for (int s = 0; s < spg; ++s) {
        const int sl = g * spg + s, b = b0 + sl;
        float dp = 0.0f;
        if (b < p.B) {
          const float pre = p.pre[((long long)t * p.B + b) * p.Hd + j];
          const float dl = dls[sl];
          a_w1 = fmaf(dl, gelu_exact(pre), a_w1);
          dp = dl * w1 * gelu_exact_grad(pre);
          a_b0 += dp;
          p.dpre[((long long)t * p.B + b) * p.Hd + j] = dp;
        }
        dps[sl * p.Hd + j] = dp;
      }
      atomicAdd(&p.dW1[t][j], a_w1);
      atomicAdd(&p.db0[t][j], a_b0);
    }
    if (warp == 0) {
      float v = lane < HD_SPB ? dls[lane] : 0.0f;
      v = warp_sum_f(v);
      if (lane == 0) atomicAdd(&p.db1[t][0], v);
    }
    __syncthreads();
    // dxn[s][k] += sum_j dpre[s][j] * W0[k][j]   (thread k walks its own row of W0: 512 contiguous bytes, L1-resident)
    for (int k = tid; k < p.d; k += HD_THREADS) {
      float acc[HD_SPB];
#pragma unroll
      for (int s = 0; s < HD_SPB; ++s) acc[s] = 0.0f;
      const float4* wrow = reinterpret_cast<const float4*>(p.W0[t] + (long long)k * p.Hd);
#pragma unroll 4
      for (int j4 = 0; j4 < p.Hd / 4; ++j4) {
        const float4 w = __ldg(wrow + j4);
#pragma unroll
        for (int s = 0; s < HD_SPB; ++s) {
          const float4 d4 = *reinterpret_cast<const float4*>(dps + s * p.Hd + 4 * j4);
          acc[s] += d4.x * w.x + d4.y * w.y + d4.z * w.z + d4.w * w.w;
        }
      }
#pragma unroll
      for (int s = 0; s < HD_SPB; ++s) dxn[s * p.d + k] += acc[s];
    }
    __syncthreads();
  }
  // ---- output-norm backward (RMSNorm: y = xh * g): dxh = dxn * g, dx = rstd (dxh - xh mean(dxh xh)), dg += dxn xh ----
  for (int s = 2 * warp; s < 2 * warp + 2; ++s) {
    const int b = b0 + s;
    float dot = 0.0f;
    for (int k = lane; k < p.d; k += 32) dot += dxn[s * p.d + k] * p.gain[k] * xs[s * p.d + k];
    dot = warp_sum_f(dot) / (float)p.d;
    if (b < p.B)
      for (int k = lane; k < p.d; k += 32)
        p.dx[(long long)b * p.lddx + k] = rs[s] * (dxn[s * p.d + k] * p.gain[k] - xs[s * p.d + k] * dot);
  }
  for (int k = tid; k < p.d; k += HD_THREADS) {
    float a = 0.0f;
#pragma unroll
    for (int s = 0; s < HD_SPB; ++s) a += dxn[s * p.d + k] * xs[s * p.d + k];
    atomicAdd(&p.dgain[k], a);
  }
}

// dW0[t][k, j] += sum_b xn[b, k] * dpre[t][b, j]: CTA (kb, t, z) owns rows 8 kb .. 8 kb + 7 of W0[t] and walks its slice of the batch.
__global__ void __launch_bounds__(HD_THREADS)
heads_dw0_kernel(const __grid_constant__ HeadsKParams p) {
  const int t = blockIdx.y, k0 = blockIdx.x * 8;
  const int ng = HD_THREADS / p.Hd;
  const int g = threadIdx.x / p.Hd, j = threadIdx.x - g * p.Hd;
  if (g >= ng) return;
  const int n_slices = ng * gridDim.z;          // the batch is cut over the thread groups and over blockIdx.z
  const int per = (p.B + n_slices - 1) / n_slices;
  const int bb = (blockIdx.z * ng + g) * per, be = min(p.B, bb + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* dp = p.dpre + (long long)t * p.B * p.Hd + j;
#pragma unroll 4
  for (int b = bb; b < be; ++b) {
    const float d = dp[(long long)b * p.Hd];
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.xn + (long long)b * p.d + k0));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p.xn + (long long)b * p.d + k0) + 1);
    acc[0] = fmaf(a.x, d, acc[0]); acc[1] = fmaf(a.y, d, acc[1]); acc[2] = fmaf(a.z, d, acc[2]); acc[3] = fmaf(a.w, d, acc[3]);
    acc[4] = fmaf(c.x, d, acc[4]); acc[5] = fmaf(c.y, d, acc[5]); acc[6] = fmaf(c.z, d, acc[6]); acc[7] = fmaf(c.w, d, acc[7]);
  }
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) atomicAdd(&p.dW0[t][(long long)(k0 + kk) * p.Hd + j], acc[kk]);
}

static int fill(const ot_heads_params* p, HeadsKParams& kp, const char* who) {
  if (!p || !p->x || !p->gain || !p->xn || !p->rstd || !p->pre) OT_FAIL(OT_ERR_INVALID_ARG, "%s: null pointer", who);
  if (p->n_tasks < 1 || p->n_tasks > OT_MAX_TASKS) OT_FAIL(OT_ERR_INVALID_ARG, "%s: n_tasks=%d", who, p->n_tasks);
  const int ng_chk = p->hidden > 0 ? HD_THREADS / p->hidden : 0;     // thread groups over the hidden columns
  if (p->d <= 0 || p->d > HD_MAX_D || p->d % 8 || p->hidden <= 0 || p->hidden > HD_MAX_H || p->hidden % 4 || ng_chk < 1 || HD_SPB % ng_chk)
    OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "%s: d=%d hidden=%d (d %% 8 == 0, d <= %d; hidden %% 4 == 0, <= %d, 256 / hidden in {1, 2, 4, 8, 16})", who,
            p->d, p->hidden, HD_MAX_D, HD_MAX_H);
  memset(&kp, 0, sizeof(kp));
  kp.x = p->x; kp.ldx = p->ldx; kp.gain = p->gain; kp.eps = p->eps; kp.B = p->B; kp.d = p->d; kp.Hd = p->hidden; kp.T = p->n_tasks;
  for (int t = 0; t < p->n_tasks; ++t) {
    if (!p->W0[t] || !p->b0[t] || !p->W1[t] || !p->b1[t]) OT_FAIL(OT_ERR_INVALID_ARG, "%s: task %d has a null parameter", who, t);
    if (reinterpret_cast<uintptr_t>(p->W0[t]) & 15) OT_FAIL(OT_ERR_INVALID_ARG, "%s: W0 must be 16-byte aligned", who);
    kp.W0[t] = p->W0[t]; kp.b0[t] = p->b0[t]; kp.W1[t] = p->W1[t]; kp.b1[t] = p->b1[t];
  }
  kp.xn = p->xn; kp.rstd = p->rstd; kp.pre = p->pre;
  return OT_OK;
}

int heads_fwd_impl(const ot_heads_params* p, cudaStream_t st) {
  HeadsKParams kp;
  int rc = fill(p, kp, "ot_heads_fwd");
  if (rc) return rc;
  if (!p->logits || !p->probs) OT_FAIL(OT_ERR_INVALID_ARG, "ot_heads_fwd: logits / probs");
  if (p->labels && (!p->loss || !p->g_bce)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_heads_fwd: labels need loss and g_bce");
  if (p->B <= 0) return OT_OK;
  kp.logits = p->logits; kp.probs = p->probs; kp.labels = p->labels; kp.loss = p->loss; kp.g_bce = p->g_bce;
  const size_t smem = sizeof(float) * HD_SPB * (p->d + p->hidden);
  heads_fwd_kernel<<<(p->B + HD_SPB - 1) / HD_SPB, HD_THREADS, smem, st>>>(kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int heads_bwd_impl(const ot_heads_params* p, cudaStream_t st) {
  HeadsKParams kp;
  int rc = fill(p, kp, "ot_heads_bwd");
  if (rc) return rc;
  if (!p->dlogit || !p->dpre || !p->dgain || !p->dx) OT_FAIL(OT_ERR_INVALID_ARG, "ot_heads_bwd: dlogit / dpre / dgain / dx");
  for (int t = 0; t < p->n_tasks; ++t) {
    if (!p->dW0[t] || !p->db0[t] || !p->dW1[t] || !p->db1[t]) OT_FAIL(OT_ERR_INVALID_ARG, "ot_heads_bwd: task %d has a null gradient", t);
    kp.dW0[t] = p->dW0[t]; kp.db0[t] = p->db0[t]; kp.dW1[t] = p->dW1[t]; kp.db1[t] = p->db1[t];
  }
  if (p->B <= 0) return OT_OK;
  kp.dlogit = p->dlogit; kp.dpre = p->dpre; kp.dgain = p->dgain; kp.dx = p->dx; kp.lddx = p->lddx;
  const size_t smem = sizeof(float) * (HD_SPB * (2 * p->d + p->hidden) + 2 * HD_SPB);
  static bool attr_done = false;
  if (!attr_done) {
    OT_CUDA_CHECK(cudaFuncSetAttribute(heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  heads_bwd_kernel<<<(p->B + HD_SPB - 1) / HD_SPB, HD_THREADS, smem, st>>>(kp);
  OT_CUDA_CHECK(cudaGetLastError());
  const int bsplit = p->B >= 1024 ? 8 : (p->B >= 128 ? 2 : 1);   // enough CTAs to fill the device; each adds its partial with atomics
  heads_dw0_kernel<<<dim3(p->d / 8, p->n_tasks, bsplit), HD_THREADS, 0, st>>>(kp);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
