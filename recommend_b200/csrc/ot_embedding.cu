// ot_embedding.cu — ID front end of the behaviour-sequence tokenizer (north_star item 1).
//
// The reference feeds the tokenizer pre-embedded 64-d events (OT/model.py:217-219, 262-265; OT/data_loader.py:126-154
// draws them at random), the paper builds them from ID embeddings (PAPER:89-109) and the repository's own idiom for that
// is "several Embedding lookups -> concat" (recall/bert_like/kuaiformer/practice/model.py:58-94).  This file is that
// idiom as three HBM-streaming kernels:
//   gather   event[e, f*EF .. (f+1)*EF) = bf16(table[field_off[f] + id[e, f]])       (a copy: bit-exact by construction)
//   scatter  grad[field_off[f] + id[e, f]] += d_event[e, f*EF ..]                     (fp32 vector reductions)
//   adagrad  every touched row once: acc += g^2 ; w -= lr * g * rsqrt(acc + eps) ; g = 0   (Keras Adagrad on the
//            summed IndexedSlices gradient; sparse_optimizer 'adagrad', sparse_lr 0.1 in OT/config.py:39-47)
// One thread moves one 16-byte output chunk (8 bf16): consecutive threads write consecutive chunks of an event row
// (full 128-byte lines) and read consecutive 32-byte sectors of one fp32 table row.
#include "ot_common.cuh"
#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

__global__ void __launch_bounds__(256)
embed_gather_kernel(const float* __restrict__ table, const long long* __restrict__ field_off, const int* __restrict__ ids,
                    __nv_bfloat16* __restrict__ out, long long ld_out, long long n_events, int n_fields, int ef,
                    const long long* __restrict__ field_rows, int* __restrict__ bad) {
  const int cpf = ef >> 3;                         // 16-byte output chunks per field
  const int cpe = cpf * n_fields;                  // ... per event
  const long long total = n_events * cpe;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / cpe;
    const int c = (int)(idx - e * cpe);
    const int f = c / cpf, k = c - f * cpf;
    const int id = ids[e * n_fields + f];
    uint4 q = make_uint4(0, 0, 0, 0);
    if (id >= 0 && id < field_rows[f]) {
      const float4* src = reinterpret_cast<const float4*>(table + (field_off[f] + id) * ef + k * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      q.x = pack_bf16x2(a.x, a.y); q.y = pack_bf16x2(a.z, a.w); q.z = pack_bf16x2(b.x, b.y); q.w = pack_bf16x2(b.z, b.w);
    } else if (bad != nullptr && k == 0) {
      atomicAdd(bad, 1);                           // out-of-vocabulary id: the row stays zero and the caller is told
    }
    *reinterpret_cast<uint4*>(out + e * ld_out + f * ef + k * 8) = q;
  }
}

__global__ void __launch_bounds__(256)
embed_scatter_kernel(const __nv_bfloat16* __restrict__ d_out, long long ld, const long long* __restrict__ field_off,
                     const int* __restrict__ ids, float* __restrict__ grad, long long n_events, int n_fields, int ef,
                     const long long* __restrict__ field_rows) {
  const int cpf = ef >> 3;
  const int cpe = cpf * n_fields;
  const long long total = n_events * cpe;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / cpe;
    const int c = (int)(idx - e * cpe);
    const int f = c / cpf, k = c - f * cpf;
    const int id = ids[e * n_fields + f];
    if (id < 0 || id >= field_rows[f]) continue;
    const uint4 q = *reinterpret_cast<const uint4*>(d_out + e * ld + f * ef + k * 8);
    float* dst = grad + (field_off[f] + id) * ef + k * 8;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(bf16lo(q.x)), "f"(bf16hi(q.x)), "f"(bf16lo(q.y)),
                 "f"(bf16hi(q.y)) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(bf16lo(q.z)), "f"(bf16hi(q.z)), "f"(bf16lo(q.w)),
                 "f"(bf16hi(q.w)) : "memory");
  }
}

// One thread per (event, field).  The first thread that meets a row in this step (atomicExch on its stamp) owns it.
__global__ void __launch_bounds__(256)
embed_adagrad_kernel(float* __restrict__ table, float* __restrict__ acc, float* __restrict__ grad, int* __restrict__ stamp,
                     const long long* __restrict__ field_off, const int* __restrict__ ids, long long n_events, int n_fields, int ef,
                     const long long* __restrict__ field_rows, int step_id, float lr, float eps) {
  const long long total = n_events * n_fields;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(idx % n_fields);
    const int id = ids[idx];
    if (id < 0 || id >= field_rows[f]) continue;
    const long long row = field_off[f] + id;
    if (atomicExch(&stamp[row], step_id) == step_id) continue;        // somebody else has this row
    float4* g4 = reinterpret_cast<float4*>(grad + row * ef);
    float4* a4 = reinterpret_cast<float4*>(acc + row * ef);
    float4* w4 = reinterpret_cast<float4*>(table + row * ef);
    for (int j = 0; j < ef / 4; ++j) {
      const float4 g = g4[j];
      float4 a = a4[j], w = w4[j];
      a.x += g.x * g.x; a.y += g.y * g.y; a.z += g.z * g.z; a.w += g.w * g.w;
      // Keras 2.12 Adagrad.update_step: variable -= lr * grad / sqrt(accumulator + epsilon)   (epsilon INSIDE the root)
      w.x -= lr * g.x * rsqrtf(a.x + eps); w.y -= lr * g.y * rsqrtf(a.y + eps);
      w.z -= lr * g.z * rsqrtf(a.z + eps); w.w -= lr * g.w * rsqrtf(a.w + eps);
      a4[j] = a; w4[j] = w;
      g4[j] = make_float4(0.f, 0.f, 0.f, 0.f);                        // the gradient table is all-zero again after the step
    }
  }
}

static int check_embed(const ot_embed_params* p, const char* who) {
  if (!p || !p->table || !p->field_off || !p->field_rows || !p->ids) OT_FAIL(OT_ERR_INVALID_ARG, "%s: null pointer", who);
  if (p->n_fields <= 0 || p->ef <= 0 || p->ef % 8) OT_FAIL(OT_ERR_UNSUPPORTED_SHAPE, "%s: n_fields=%d ef=%d (ef %% 8 == 0)", who, p->n_fields, p->ef);
  if (reinterpret_cast<uintptr_t>(p->table) & 15) OT_FAIL(OT_ERR_INVALID_ARG, "%s: table must be 16-byte aligned", who);
  return OT_OK;
}

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  return (int)(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
}

int embed_gather_impl(const ot_embed_params* p, cudaStream_t st) {
  int rc = check_embed(p, "ot_embed_gather_fwd");
  if (rc) return rc;
  if (!p->events || (p->ld_events % 8)) OT_FAIL(OT_ERR_INVALID_ARG, "ot_embed_gather_fwd: events / ld_events");
  if (p->n_events <= 0) return OT_OK;
  embed_gather_kernel<<<grid_for(p->n_events * p->n_fields * (p->ef / 8)), 256, 0, st>>>(
      p->table, (const long long*)p->field_off, p->ids, (__nv_bfloat16*)p->events, p->ld_events, p->n_events, p->n_fields, p->ef,
      (const long long*)p->field_rows, p->bad_ids);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int embed_scatter_impl(const ot_embed_params* p, cudaStream_t st) {
  int rc = check_embed(p, "ot_embed_scatter_bwd");
  if (rc) return rc;
  if (!p->events || (p->ld_events % 8) || !p->grad || (reinterpret_cast<uintptr_t>(p->grad) & 15))
    OT_FAIL(OT_ERR_INVALID_ARG, "ot_embed_scatter_bwd: events (the event gradients) / grad");
  if (p->n_events <= 0) return OT_OK;
  embed_scatter_kernel<<<grid_for(p->n_events * p->n_fields * (p->ef / 8)), 256, 0, st>>>(
      (const __nv_bfloat16*)p->events, p->ld_events, (const long long*)p->field_off, p->ids, p->grad, p->n_events, p->n_fields, p->ef,
      (const long long*)p->field_rows);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

int embed_adagrad_impl(const ot_embed_params* p, cudaStream_t st) {
  int rc = check_embed(p, "ot_embed_adagrad_step");
  if (rc) return rc;
  if (!p->grad || !p->acc || !p->stamp) OT_FAIL(OT_ERR_INVALID_ARG, "ot_embed_adagrad_step: grad / acc / stamp");
  if (p->step_id == 0) OT_FAIL(OT_ERR_INVALID_ARG, "ot_embed_adagrad_step: step_id 0 is the stamp table's initial value");
  if (p->n_events <= 0) return OT_OK;
  embed_adagrad_kernel<<<grid_for(p->n_events * p->n_fields), 256, 0, st>>>(
      const_cast<float*>(p->table), p->acc, p->grad, p->stamp, (const long long*)p->field_off, p->ids, p->n_events, p->n_fields, p->ef,
      (const long long*)p->field_rows, p->step_id, p->lr, p->eps);
  OT_CUDA_CHECK(cudaGetLastError());
  return OT_OK;
}

}  // namespace ot
