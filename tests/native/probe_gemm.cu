// probe_gemm.cu — standalone GPU probe for ot_mixed_gemm / ot_wgrad against a CPU fp32 reference.
// Usage: probe_gemm <test-id>   (each test in its own process so that a device fault cannot poison
// the next one).  Prints "PASS"/"FAIL" lines; exit code 0 only if every check of the test passed.
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

static uint32_t rng_state = 12345u;
static float frand() {  // uniform in [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct HostMat {
  std::vector<float> f;            // bf16-rounded values as float
  std::vector<__nv_bfloat16> h;
  void* d = nullptr;
  void init(size_t n, float scale) {
    f.resize(n); h.resize(n);
    for (size_t i = 0; i < n; ++i) { h[i] = __float2bfloat16(frand() * scale); f[i] = __bfloat162float(h[i]); }
    CK(cudaMalloc(&d, n * 2));
    CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  }
};

static double gelu_ref(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }
static double gelu_grad_ref(double x) {
  return 0.5 * (1.0 + erf(x / sqrt(2.0))) + x * exp(-0.5 * x * x) / sqrt(2.0 * M_PI);
}

struct Seg { int row_start, n_units, rows_per_unit, group_start, group_stride; };

// Generic GEMM test.  A is [rows_total, K] (2-D mode).
static int test_gemm(const char* name, int rows_total, int N, int K, int n_groups, std::vector<Seg> segs, int flags,
                     int block_n, int swizzle, bool dual_out) {
  HostMat A, W, R, X;
  A.init((size_t)rows_total * K, 1.0f);
  W.init((size_t)n_groups * N * K, 0.25f);
  R.init((size_t)rows_total * N, 1.0f);
  X.init((size_t)rows_total * N, 2.0f);
  std::vector<float> bias((size_t)n_groups * N), rscale(rows_total);
  for (auto& b : bias) b = frand();
  for (auto& r : rscale) r = 0.5f + 0.5f * fabsf(frand());
  float *d_bias, *d_rs;
  CK(cudaMalloc(&d_bias, bias.size() * 4)); CK(cudaMemcpy(d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_rs, rscale.size() * 4)); CK(cudaMemcpy(d_rs, rscale.data(), rscale.size() * 4, cudaMemcpyHostToDevice));
  __nv_bfloat16 *d_out, *d_out2;
  CK(cudaMalloc(&d_out, (size_t)rows_total * N * 2)); CK(cudaMemset(d_out, 0xFF, (size_t)rows_total * N * 2));
  CK(cudaMalloc(&d_out2, (size_t)rows_total * N * 2)); CK(cudaMemset(d_out2, 0xFF, (size_t)rows_total * N * 2));

  ot_gemm_params p; memset(&p, 0, sizeof(p));
  p.A = A.d; p.a_dim1 = rows_total; p.a_dim2 = 1; p.a_stride1 = K; p.a_stride2 = 0; p.a_transposed = 0;
  p.n_groups = n_groups; p.W = W.d; p.ldw = K; p.N = N; p.K = K; p.n_segs = (int)segs.size(); p.flags = flags;
  for (size_t s = 0; s < segs.size(); ++s) {
    p.segs[s].row_start = segs[s].row_start; p.segs[s].n_units = segs[s].n_units; p.segs[s].rows_per_unit = segs[s].rows_per_unit;
    p.segs[s].group_start = segs[s].group_start; p.segs[s].group_stride = segs[s].group_stride; p.segs[s].a_row_start = segs[s].row_start;
  }
  p.out = d_out; p.ldo = N; p.out2 = dual_out ? d_out2 : nullptr; p.ldo2 = N;
  p.res = R.d; p.ldr = N; p.aux = X.d; p.ldaux = N; p.bias = d_bias; p.bias_group_stride = N; p.row_scale = d_rs;
  p.block_n = block_n; p.swizzle = swizzle;
  int rc = ot_mixed_gemm(&p, nullptr);
  if (rc) { printf("FAIL %s: launch rc=%d (%s)\n", name, rc, ot_last_error_string()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL %s: kernel error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<__nv_bfloat16> out((size_t)rows_total * N), out2((size_t)rows_total * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out2.data(), d_out2, out2.size() * 2, cudaMemcpyDeviceToHost));

  // reference
  std::vector<int> row_group(rows_total, -1);
  for (auto& s : segs)
    for (int u = 0; u < s.n_units; ++u)
      for (int r = 0; r < s.rows_per_unit; ++r) row_group[s.row_start + u * s.rows_per_unit + r] = s.group_start + u * s.group_stride;
  double max_err = 0, max_err2 = 0; long bad = 0, untouched_bad = 0; int first_bad_r = -1, first_bad_c = -1; double fb_got = 0, fb_ref = 0;
  for (int r = 0; r < rows_total; ++r) {
    const int g = row_group[r];
    for (int n = 0; n < N; ++n) {
      const float got = __bfloat162float(out[(size_t)r * N + n]);
      if (g < 0) {  // row not covered by any segment must be untouched (0xFFFF = NaN pattern)
        uint16_t raw; memcpy(&raw, &out[(size_t)r * N + n], 2);
        if (raw != 0xFFFF) ++untouched_bad;
        continue;
      }
      double acc = 0;
      const float* a = &A.f[(size_t)r * K];
      const float* w = &W.f[((size_t)g * N + n) * K];
      for (int k = 0; k < K; ++k) acc += (double)a[k] * w[k];
      double v = acc;
      if (flags & OT_EPI_ROW_SCALE) v *= rscale[r];
      if (flags & OT_EPI_BIAS) v += bias[(size_t)g * N + n];
      double pre = v;
      if (flags & OT_EPI_GELU) v = gelu_ref(v);
      if (flags & OT_EPI_GELU_GRAD) v *= gelu_grad_ref(X.f[(size_t)r * N + n]);
      if (flags & OT_EPI_RESIDUAL) v += R.f[(size_t)r * N + n];
      const double err = fabs(got - v) / (1.0 + fabs(v));
      if (!(err <= 2e-2)) { if (bad == 0) { first_bad_r = r; first_bad_c = n; fb_got = got; fb_ref = v; } ++bad; }
      if (err > max_err || err != err) max_err = err;
      if (dual_out) {
        const float got2 = __bfloat162float(out2[(size_t)r * N + n]);
        const double err2 = fabs(got2 - pre) / (1.0 + fabs(pre));
        if (!(err2 <= 2e-2)) ++bad;
        if (err2 > max_err2) max_err2 = err2;
      }
    }
  }
  const bool ok = bad == 0 && untouched_bad == 0;
  printf("%s %s: rows=%d N=%d K=%d bn=%d swz=%d flags=%d max_rel_err=%.3e max_rel_err2=%.3e bad=%ld untouched_bad=%ld", ok ? "PASS" : "FAIL",
         name, rows_total, N, K, block_n, swizzle, flags, max_err, max_err2, bad, untouched_bad);
  if (!ok && first_bad_r >= 0) printf(" first_bad=(%d,%d) got=%g ref=%g", first_bad_r, first_bad_c, fb_got, fb_ref);
  printf("\n");
  if (!ok) {  // error map by 8x8 blocks of the first 128x64 region, to see layout mistakes
    printf("  got[0..3][0..7]:");
    for (int r = 0; r < 4 && r < rows_total; ++r) { printf("\n   "); for (int n = 0; n < 8; ++n) printf(" %9.4f", __bfloat162float(out[(size_t)r * N + n])); }
    printf("\n");
  }
  return ok ? 0 : 1;
}

// transposed-A mode: A is events [Bsz, L, K]; output row = (off + l)*Bsz + b
static int test_gemm_transposed(int Bsz, int L, int K, int N, int swizzle) {
  HostMat A, W;
  A.init((size_t)Bsz * L * K, 1.0f);
  W.init((size_t)N * K, 0.25f);
  const int off = 3;
  const int rows_total = (off + L + 2) * Bsz;
  __nv_bfloat16* d_out; CK(cudaMalloc(&d_out, (size_t)rows_total * N * 2)); CK(cudaMemset(d_out, 0xFF, (size_t)rows_total * N * 2));
  std::vector<float> bias(N); for (auto& b : bias) b = frand();
  float* d_bias; CK(cudaMalloc(&d_bias, N * 4)); CK(cudaMemcpy(d_bias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  ot_gemm_params p; memset(&p, 0, sizeof(p));
  p.A = A.d; p.a_dim1 = L; p.a_dim2 = Bsz; p.a_stride1 = K; p.a_stride2 = (int64_t)L * K; p.a_transposed = 1;
  p.n_groups = 1; p.W = W.d; p.ldw = K; p.N = N; p.K = K; p.n_segs = 1; p.flags = OT_EPI_BIAS;
  p.segs[0].row_start = off * Bsz; p.segs[0].n_units = L; p.segs[0].rows_per_unit = Bsz; p.segs[0].group_start = 0; p.segs[0].group_stride = 0; p.segs[0].a_row_start = 0;
  p.out = d_out; p.ldo = N; p.bias = d_bias; p.bias_group_stride = 0; p.swizzle = swizzle;
  int rc = ot_mixed_gemm(&p, nullptr);
  if (rc) { printf("FAIL gemm_transposed: rc=%d (%s)\n", rc, ot_last_error_string()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL gemm_transposed: kernel error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<__nv_bfloat16> out((size_t)rows_total * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0; long bad = 0;
  for (int l = 0; l < L; ++l) for (int b = 0; b < Bsz; ++b) for (int n = 0; n < N; ++n) {
    double acc = bias[n];
    for (int k = 0; k < K; ++k) acc += (double)A.f[((size_t)b * L + l) * K + k] * W.f[(size_t)n * K + k];
    const float got = __bfloat162float(out[((size_t)(off + l) * Bsz + b) * N + n]);
    const double err = fabs(got - acc) / (1.0 + fabs(acc));
    if (!(err <= 2e-2)) ++bad;
    if (err > max_err || err != err) max_err = err;
  }
  long untouched_bad = 0;
  for (int r = 0; r < rows_total; ++r) {
    if (r >= off * Bsz && r < (off + L) * Bsz) continue;
    for (int n = 0; n < N; ++n) { uint16_t raw; memcpy(&raw, &out[(size_t)r * N + n], 2); if (raw != 0xFFFF) ++untouched_bad; }
  }
  const bool ok = bad == 0 && untouched_bad == 0;
  printf("%s gemm_transposed: B=%d L=%d K=%d N=%d swz=%d max_rel_err=%.3e bad=%ld untouched_bad=%ld\n", ok ? "PASS" : "FAIL", Bsz, L, K, N, swizzle, max_err, bad, untouched_bad);
  return ok ? 0 : 1;
}

// wgrad: shared run (units=1 or many with group_stride 0) + optional NS run
static int test_wgrad(const char* name, int Mdim, int Ndim, int shared_units, int shared_rpu, int ns_units, int ns_rpu, int block_n, int swizzle, int target_ctas) {
  const int rows_shared = shared_units * shared_rpu;
  const int rows_total = rows_shared + ns_units * ns_rpu;
  HostMat P, Q;
  P.init((size_t)rows_total * Mdim, 1.0f);
  Q.init((size_t)rows_total * Ndim, 1.0f);
  const int n_groups = 1 + ns_units;
  float* d_C; CK(cudaMalloc(&d_C, (size_t)n_groups * Mdim * Ndim * 4)); CK(cudaMemset(d_C, 0, (size_t)n_groups * Mdim * Ndim * 4));
  ot_wgrad_params p; memset(&p, 0, sizeof(p));
  p.Mdim = Mdim; p.Ndim = Ndim; p.swizzle = swizzle; p.block_n = block_n; p.target_ctas = target_ctas;
  p.C = d_C; p.c_group_stride = (int64_t)Mdim * Ndim; p.c_stride_m = Ndim; p.c_stride_n = 1;
  int ns = 0;
  p.segs[ns].P = P.d; p.segs[ns].p_stride_row = Mdim; p.segs[ns].p_stride_unit = (int64_t)shared_rpu * Mdim;
  p.segs[ns].Q = Q.d; p.segs[ns].q_stride_row = Ndim; p.segs[ns].q_stride_unit = (int64_t)shared_rpu * Ndim;
  p.segs[ns].n_units = shared_units; p.segs[ns].rows_per_unit = shared_rpu; p.segs[ns].group_start = 0; p.segs[ns].group_stride = 0; ++ns;
  if (ns_units > 0) {
    p.segs[ns].P = (const __nv_bfloat16*)P.d + (size_t)rows_shared * Mdim; p.segs[ns].p_stride_row = Mdim; p.segs[ns].p_stride_unit = (int64_t)ns_rpu * Mdim;
    p.segs[ns].Q = (const __nv_bfloat16*)Q.d + (size_t)rows_shared * Ndim; p.segs[ns].q_stride_row = Ndim; p.segs[ns].q_stride_unit = (int64_t)ns_rpu * Ndim;
    p.segs[ns].n_units = ns_units; p.segs[ns].rows_per_unit = ns_rpu; p.segs[ns].group_start = 1; p.segs[ns].group_stride = 1; ++ns;
  }
  p.n_segs = ns;
  int rc = ot_wgrad(&p, nullptr);
  if (rc) { printf("FAIL %s: rc=%d (%s)\n", name, rc, ot_last_error_string()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL %s: kernel error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> C((size_t)n_groups * Mdim * Ndim);
  CK(cudaMemcpy(C.data(), d_C, C.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<double> ref(C.size(), 0.0);
  for (int r = 0; r < rows_total; ++r) {
    const int g = r < rows_shared ? 0 : 1 + (r - rows_shared) / ns_rpu;
    double* cg = &ref[(size_t)g * Mdim * Ndim];
    for (int m = 0; m < Mdim; ++m) {
      const double pv = P.f[(size_t)r * Mdim + m];
      const float* q = &Q.f[(size_t)r * Ndim];
      double* row = cg + (size_t)m * Ndim;
      for (int n = 0; n < Ndim; ++n) row[n] += pv * q[n];
    }
  }
  double max_err = 0; long bad = 0; size_t fb = 0;
  for (size_t i = 0; i < C.size(); ++i) {
    const double err = fabs(C[i] - ref[i]) / (1.0 + fabs(ref[i]));
    if (!(err <= 5e-3)) { if (!bad) fb = i; ++bad; }
    if (err > max_err || err != err) max_err = err;
  }
  const bool ok = bad == 0;
  printf("%s %s: Mdim=%d Ndim=%d shared=%dx%d ns=%dx%d bn=%d swz=%d max_rel_err=%.3e bad=%ld", ok ? "PASS" : "FAIL", name, Mdim, Ndim, shared_units, shared_rpu, ns_units, ns_rpu, block_n, swizzle, max_err, bad);
  if (!ok) printf(" first_bad idx=%zu (g=%zu m=%zu n=%zu) got=%g ref=%g", fb, fb / ((size_t)Mdim * Ndim), (fb / Ndim) % Mdim, fb % Ndim, C[fb], ref[fb]);
  printf("\n");
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const int id = argc > 1 ? atoi(argv[1]) : 0;
  const int ALL = OT_EPI_ROW_SCALE | OT_EPI_BIAS;
  switch (id) {
    case 0: return test_gemm("gemm_min", 128, 64, 64, 1, {{0, 1, 128, 0, 0}}, 0, 64, 0, false);
    case 1: return test_gemm("gemm_k256", 256, 64, 256, 1, {{0, 1, 256, 0, 0}}, 0, 64, 0, false);
    case 2: return test_gemm("gemm_bn128", 384, 128, 128, 1, {{0, 1, 384, 0, 0}}, 0, 128, 0, false);
    case 3: return test_gemm("gemm_bn256_partial", 1000, 256, 256, 1, {{0, 1, 1000, 0, 0}}, 0, 256, 0, false);
    case 4: return test_gemm("gemm_many_tiles", 148 * 128 * 2 + 77, 512, 256, 1, {{0, 1, 148 * 128 * 2 + 77, 0, 0}}, 0, 256, 0, false);
    case 5: return test_gemm("gemm_grouped", 3 * 160 + 5 * 160, 768, 256, 6, {{0, 1, 480, 0, 0}, {480, 5, 160, 1, 1}}, 0, 256, 0, false);
    case 6: return test_gemm("gemm_grouped_small_b", 10 * 32 + 4 * 32, 256, 256, 9, {{0, 1, 320, 0, 0}, {320, 4, 32, 5, 1}}, ALL, 0, 0, false);
    case 7: return test_gemm("gemm_gelu_dual", 500, 1024, 256, 1, {{0, 1, 500, 0, 0}}, OT_EPI_BIAS | OT_EPI_GELU, 0, 0, true);
    case 8: return test_gemm("gemm_residual", 500, 256, 1024, 1, {{0, 1, 500, 0, 0}}, OT_EPI_BIAS | OT_EPI_RESIDUAL, 0, 0, false);
    case 9: return test_gemm("gemm_gelugrad", 300, 1024, 256, 1, {{0, 1, 300, 0, 0}}, OT_EPI_GELU_GRAD, 0, 0, false);
    case 10: return test_gemm("gemm_res_rowscale", 300, 256, 256, 1, {{0, 1, 300, 0, 0}}, OT_EPI_RESIDUAL | OT_EPI_ROW_SCALE, 0, 0, false);
    case 11: return test_gemm("gemm_hole", 640, 128, 64, 2, {{0, 1, 200, 0, 0}, {384, 2, 100, 0, 1}}, 0, 0, 0, false);
    case 12: return test_gemm("gemm_swz64", 300, 256, 96, 1, {{0, 1, 300, 0, 0}}, OT_EPI_BIAS, 0, 64, false);
    case 13: return test_gemm_transposed(200, 7, 64, 256, 0);
    case 14: return test_gemm_transposed(64, 5, 64, 128, 64);
    case 20: return test_wgrad("wgrad_min", 128, 64, 1, 64, 0, 0, 64, 0, 1);
    case 21: return test_wgrad("wgrad_k1024", 128, 64, 1, 1024, 0, 0, 64, 0, 1);
    case 22: return test_wgrad("wgrad_bn256", 256, 768, 1, 4096, 0, 0, 256, 0, 0);
    case 23: return test_wgrad("wgrad_grouped", 256, 256, 1, 3000, 4, 200, 0, 0, 0);
    case 24: return test_wgrad("wgrad_units_shared", 128, 256, 7, 100, 0, 0, 0, 0, 0);
    case 25: return test_wgrad("wgrad_swz64", 128, 96 + 32, 1, 500, 2, 70, 0, 64, 0);
    case 26: return test_wgrad("wgrad_big", 1024, 256, 1, 20000, 3, 512, 0, 0, 0);
    default: printf("unknown test %d\n", id); return 3;
  }
}
