// probe_attn.cu — standalone GPU probe for ot_attn_fwd / ot_attn_bwd against a CPU fp64 reference.
// Usage: probe_attn <test-id>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/onetrans_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 777u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
struct HostMat {
  std::vector<float> f; std::vector<__nv_bfloat16> h; void* d = nullptr;
  void init(size_t n, float scale) {
    f.resize(n); h.resize(n);
    for (size_t i = 0; i < n; ++i) { h[i] = __float2bfloat16(frand() * scale); f[i] = __bfloat162float(h[i]); }
    CK(cudaMalloc(&d, n * 2)); CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  }
};
static std::vector<float> fetch_bf16(const void* d, size_t n) {
  std::vector<__nv_bfloat16> h(n); CK(cudaMemcpy(h.data(), d, n * 2, cudaMemcpyDeviceToHost));
  std::vector<float> f(n); for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(h[i]); return f;
}

struct Cmp { double max_err = 0; long bad = 0; size_t first = 0; double got = 0, ref = 0; };
static void cmp_one(Cmp& c, size_t idx, double got, double ref, double tol) {
  const double err = fabs(got - ref) / (1.0 + fabs(ref));
  if (!(err <= tol)) { if (!c.bad) { c.first = idx; c.got = got; c.ref = ref; } ++c.bad; }
  if (err > c.max_err || err != err) c.max_err = err;
}

static int test_attn(const char* name, int B, int H, int DH, int Lq, int Lk, int swizzle, bool do_bwd, float qscale) {
  const int d = H * DH;
  const int ldq = d, ldkv = 2 * d;  // K and V live in one [Lk*B, 2d] buffer like the model uses them
  HostMat Q, KV, dO;
  Q.init((size_t)Lq * B * ldq, qscale);
  KV.init((size_t)Lk * B * ldkv, qscale);
  dO.init((size_t)Lq * B * d, 1.0f);
  __nv_bfloat16 *o, *dq, *dkv; float *lse, *delta;
  CK(cudaMalloc(&o, (size_t)Lq * B * d * 2)); CK(cudaMemset(o, 0xFF, (size_t)Lq * B * d * 2));
  CK(cudaMalloc(&dq, (size_t)Lq * B * d * 2)); CK(cudaMemset(dq, 0xFF, (size_t)Lq * B * d * 2));
  CK(cudaMalloc(&dkv, (size_t)Lk * B * ldkv * 2)); CK(cudaMemset(dkv, 0xFF, (size_t)Lk * B * ldkv * 2));
  CK(cudaMalloc(&lse, (size_t)B * H * Lq * 4)); CK(cudaMalloc(&delta, (size_t)B * H * Lq * 4));
  ot_attn_params p; memset(&p, 0, sizeof(p));
  p.q = Q.d; p.ldq = ldq; p.k = KV.d; p.ldk = ldkv; p.v = (const __nv_bfloat16*)KV.d + d; p.ldv = ldkv;
  p.o = o; p.ldo = d; p.lse = lse; p.d_o = dO.d; p.lddo = d; p.dq = dq; p.lddq = d;
  p.dk = dkv; p.lddk = ldkv; p.dv = dkv + d; p.lddv = ldkv; p.delta = delta;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.head_dim = DH; p.swizzle = swizzle;
  int rc = ot_attn_fwd(&p, nullptr);
  if (rc) { printf("FAIL %s: fwd rc=%d (%s)\n", name, rc, ot_last_error_string()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL %s: fwd kernel error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> o_h = fetch_bf16(o, (size_t)Lq * B * d);
  std::vector<float> lse_h((size_t)B * H * Lq); CK(cudaMemcpy(lse_h.data(), lse, lse_h.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dq_h, dkv_h;
  if (do_bwd) {
    rc = ot_attn_bwd(&p, nullptr);
    if (rc) { printf("FAIL %s: bwd rc=%d (%s)\n", name, rc, ot_last_error_string()); return 1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAIL %s: bwd kernel error %s\n", name, cudaGetErrorString(e)); return 1; }
    dq_h = fetch_bf16(dq, (size_t)Lq * B * d);
    dkv_h = fetch_bf16(dkv, (size_t)Lk * B * ldkv);
  }
  // ---- CPU reference (fp64) ----
  const int off = Lk - Lq;
  const double scale = 1.0 / sqrt((double)DH);
  Cmp c_o, c_lse, c_dq, c_dk, c_dv;
  std::vector<double> s(Lk), pr(Lk), dP(Lk);
  std::vector<double> dK_ref, dV_ref;
  for (int b = 0; b < B; ++b) for (int h = 0; h < H; ++h) {
    if (do_bwd) { dK_ref.assign((size_t)Lk * DH, 0.0); dV_ref.assign((size_t)Lk * DH, 0.0); }
    for (int i = 0; i < Lq; ++i) {
      const float* q = &Q.f[((size_t)i * B + b) * ldq + h * DH];
      const int nk = off + i + 1;
      double m = -1e300;
      for (int j = 0; j < nk; ++j) {
        const float* k = &KV.f[((size_t)j * B + b) * ldkv + h * DH];
        double a = 0; for (int t = 0; t < DH; ++t) a += (double)q[t] * k[t];
        s[j] = a * scale; if (s[j] > m) m = s[j];
      }
      double l = 0; for (int j = 0; j < nk; ++j) { pr[j] = exp(s[j] - m); l += pr[j]; }
      for (int j = 0; j < nk; ++j) pr[j] /= l;
      double oref[128];
      for (int t = 0; t < DH; ++t) oref[t] = 0;
      for (int j = 0; j < nk; ++j) { const float* v = &KV.f[((size_t)j * B + b) * ldkv + d + h * DH]; for (int t = 0; t < DH; ++t) oref[t] += pr[j] * v[t]; }
      for (int t = 0; t < DH; ++t) cmp_one(c_o, ((size_t)i * B + b) * d + h * DH + t, o_h[((size_t)i * B + b) * d + h * DH + t], oref[t], 2e-2);
      cmp_one(c_lse, ((size_t)b * H + h) * Lq + i, lse_h[((size_t)b * H + h) * Lq + i], m + log(l), 2e-3);
      if (do_bwd) {
        const float* g = &dO.f[((size_t)i * B + b) * d + h * DH];
        double delta_r = 0; for (int t = 0; t < DH; ++t) delta_r += oref[t] * g[t];
        double dqref[128]; for (int t = 0; t < DH; ++t) dqref[t] = 0;
        for (int j = 0; j < nk; ++j) {
          const float* v = &KV.f[((size_t)j * B + b) * ldkv + d + h * DH];
          const float* k = &KV.f[((size_t)j * B + b) * ldkv + h * DH];
          double dp = 0; for (int t = 0; t < DH; ++t) dp += (double)g[t] * v[t];
          const double ds = pr[j] * (dp - delta_r) * scale;
          for (int t = 0; t < DH; ++t) { dqref[t] += ds * k[t]; dK_ref[(size_t)j * DH + t] += ds * q[t]; dV_ref[(size_t)j * DH + t] += pr[j] * g[t]; }
        }
        for (int t = 0; t < DH; ++t) cmp_one(c_dq, ((size_t)i * B + b) * d + h * DH + t, dq_h[((size_t)i * B + b) * d + h * DH + t], dqref[t], 3e-2);
      }
    }
    if (do_bwd) for (int j = 0; j < Lk; ++j) for (int t = 0; t < DH; ++t) {
      cmp_one(c_dk, (size_t)j, dkv_h[((size_t)j * B + b) * ldkv + h * DH + t], dK_ref[(size_t)j * DH + t], 3e-2);
      cmp_one(c_dv, (size_t)j, dkv_h[((size_t)j * B + b) * ldkv + d + h * DH + t], dV_ref[(size_t)j * DH + t], 3e-2);
    }
  }
  const bool ok = !c_o.bad && !c_lse.bad && !c_dq.bad && !c_dk.bad && !c_dv.bad;
  printf("%s %s: B=%d H=%d DH=%d Lq=%d Lk=%d swz=%d | O err=%.2e bad=%ld | lse err=%.2e bad=%ld", ok ? "PASS" : "FAIL", name, B, H, DH, Lq, Lk, swizzle,
         c_o.max_err, c_o.bad, c_lse.max_err, c_lse.bad);
  if (do_bwd) printf(" | dQ err=%.2e bad=%ld | dK err=%.2e bad=%ld | dV err=%.2e bad=%ld", c_dq.max_err, c_dq.bad, c_dk.max_err, c_dk.bad, c_dv.max_err, c_dv.bad);
  printf("\n");
  if (c_o.bad) printf("   O first bad idx=%zu (row=%zu col=%zu) got=%g ref=%g\n", c_o.first, c_o.first / d, c_o.first % d, c_o.got, c_o.ref);
  if (c_lse.bad) printf("   lse first bad idx=%zu got=%g ref=%g\n", c_lse.first, c_lse.got, c_lse.ref);
  if (c_dq.bad) printf("   dQ first bad idx=%zu (row=%zu col=%zu) got=%g ref=%g\n", c_dq.first, c_dq.first / d, c_dq.first % d, c_dq.got, c_dq.ref);
  if (c_dk.bad) printf("   dK first bad key=%zu got=%g ref=%g\n", c_dk.first, c_dk.got, c_dk.ref);
  if (c_dv.bad) printf("   dV first bad key=%zu got=%g ref=%g\n", c_dv.first, c_dv.got, c_dv.ref);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const int id = argc > 1 ? atoi(argv[1]) : 0;
  switch (id) {
    case 0: return test_attn("fwd_1tile", 1, 1, 64, 128, 128, 0, false, 1.0f);
    case 1: return test_attn("fwd_2blk", 2, 2, 64, 128, 256, 0, false, 1.0f);
    case 2: return test_attn("fwd_ragged", 3, 4, 64, 202, 288, 0, false, 1.0f);
    case 3: return test_attn("fwd_c2_l0", 2, 4, 64, 458, 544, 0, false, 2.0f);
    case 4: return test_attn("fwd_small_tail", 5, 4, 64, 13, 27, 0, false, 1.0f);
    case 5: return test_attn("fwd_dh64_swz64", 2, 2, 64, 150, 300, 64, false, 1.0f);
    case 6: return test_attn("fwd_dh96", 2, 4, 96, 224, 288, 0, false, 1.0f);
    case 7: return test_attn("fwd_many_items", 40, 4, 64, 288, 373, 0, false, 1.0f);
    case 10: return test_attn("bwd_1tile", 1, 1, 64, 128, 128, 0, true, 1.0f);
    case 11: return test_attn("bwd_2blk", 2, 2, 64, 128, 256, 0, true, 1.0f);
    case 12: return test_attn("bwd_ragged", 3, 4, 64, 202, 288, 0, true, 1.0f);
    case 13: return test_attn("bwd_c2_l0", 2, 4, 64, 458, 544, 0, true, 2.0f);
    case 14: return test_attn("bwd_small_tail", 5, 4, 64, 13, 27, 0, true, 1.0f);
    case 15: return test_attn("bwd_dh64_swz64", 2, 2, 64, 150, 300, 64, true, 1.0f);
    case 16: return test_attn("bwd_dh96", 2, 4, 96, 224, 288, 0, true, 1.0f);
    case 17: return test_attn("bwd_many_items", 40, 4, 64, 288, 373, 0, true, 1.0f);
    case 18: return test_attn("bwd_long", 1, 4, 64, 1024, 2048, 0, true, 1.0f);
    default: printf("unknown test %d\n", id); return 3;
  }
}
