// ot_api.cu — the C-ABI surface of libonetrans_sm100.so (include/onetrans_b200.h): thin extern "C"
// wrappers that validate, build tensor maps and enqueue the sm_100a kernels on the caller's stream.
// No C++ exception crosses this boundary, nothing here allocates device memory or synchronises.
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "ot_host.h"
#include "../../include/onetrans_b200.h"

namespace ot {

char* error_slot() {
  static thread_local char buf[512] = {0};
  return buf;
}

PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
    }
  }
  return fn;
}

// Work counters of the persistent kernels (dynamic tile / item scheduling): a per-device ring of 1024 ints, allocated once
// on first use (the only device allocation the library ever makes, 4 KB); every launch takes the next slot and zeroes it
// on its stream.  OT_STATIC_SCHED=1 (or a failed allocation) returns NULL = static round-robin schedule.
int* sched_slot(cudaStream_t st) {
  static int* base[64] = {nullptr};
  static std::atomic<unsigned> next[64];
  static std::mutex mu;
  static const bool disabled = [] { const char* e = getenv("OT_STATIC_SCHED"); return e && e[0] && e[0] != '0'; }();
  if (disabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (base[dev] == nullptr) {
    std::lock_guard<std::mutex> lk(mu);
    if (base[dev] == nullptr) {
      int* ptr = nullptr;
      if (cudaMalloc(&ptr, 1024 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      base[dev] = ptr;
    }
  }
  int* slot = base[dev] + (next[dev].fetch_add(1) % 1024u);
  if (cudaMemsetAsync(slot, 0, sizeof(int), st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return slot;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

int mixed_gemm_impl(const ot_gemm_params* p, cudaStream_t st);
int wgrad_impl(const ot_wgrad_params* p, cudaStream_t st);
int ffn_fwd_impl(const ot_ffn_params* p, cudaStream_t st);
int ffn_bwd_impl(const ot_ffn_params* p, cudaStream_t st);
int attn_fwd_impl(const ot_attn_params* p, cudaStream_t st);
int attn_bwd_impl(const ot_attn_params* p, cudaStream_t st);
int attn_cached_impl(const ot_attn_cached_params* p, cudaStream_t st);
int rmsnorm_fwd_impl(const ot_rmsnorm_params* p, cudaStream_t st);
int rmsnorm_bwd_impl(const ot_rmsnorm_params* p, cudaStream_t st);
int ns_tokenizer_fwd_impl(const ot_ns_tokenizer_params* p, cudaStream_t st);
int ns_tokenizer_bwd_impl(const ot_ns_tokenizer_params* p, cudaStream_t st);
int fill_rows_impl(const float* vec, void* out, long long ldo, long long row0, long long n_rows, int d, cudaStream_t st);
int colsum_impl(const ot_colsum_params* p, cudaStream_t st);
int dropout_mask_impl(const void* in, long long ld_in, void* out, long long ld_out, long long rows, int cols, uint32_t seed,
                      float rate, cudaStream_t st);
int clip_rmsprop_impl(const ot_rmsprop_params* p, cudaStream_t st);
int embed_gather_impl(const ot_embed_params* p, cudaStream_t st);
int embed_scatter_impl(const ot_embed_params* p, cudaStream_t st);
int embed_adagrad_impl(const ot_embed_params* p, cudaStream_t st);
int heads_fwd_impl(const ot_heads_params* p, cudaStream_t st);
int heads_bwd_impl(const ot_heads_params* p, cudaStream_t st);
int metrics_update_impl(const ot_metrics_params* p, cudaStream_t st);
int metrics_result_impl(const ot_metrics_params* p, cudaStream_t st);
int auc_pack_impl(const ot_auc_params* p, cudaStream_t st);
int auc_ranksum_impl(const ot_auc_params* p, cudaStream_t st);

}  // namespace ot

extern "C" {

int ot_version(void) { return OT_ABI_VERSION; }
const char* ot_last_error_string(void) { return ot::error_slot(); }
int ot_num_sms(void) { return ot::num_sms(); }

int ot_mixed_gemm(const ot_gemm_params* p, void* stream) {
  return ot::mixed_gemm_impl(p, static_cast<cudaStream_t>(stream));
}
int ot_ffn_fwd(const ot_ffn_params* p, void* stream) { return ot::ffn_fwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_ffn_bwd(const ot_ffn_params* p, void* stream) { return ot::ffn_bwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_wgrad(const ot_wgrad_params* p, void* stream) {
  return ot::wgrad_impl(p, static_cast<cudaStream_t>(stream));
}

int ot_attn_fwd(const ot_attn_params* p, void* stream) {
  return ot::attn_fwd_impl(p, static_cast<cudaStream_t>(stream));
}
int ot_attn_bwd(const ot_attn_params* p, void* stream) {
  return ot::attn_bwd_impl(p, static_cast<cudaStream_t>(stream));
}

int ot_attn_ns_cached_fwd(const ot_attn_cached_params* p, void* stream) {
  return ot::attn_cached_impl(p, static_cast<cudaStream_t>(stream));
}
int ot_rmsnorm_fwd(const ot_rmsnorm_params* p, void* stream) { return ot::rmsnorm_fwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_rmsnorm_bwd(const ot_rmsnorm_params* p, void* stream) { return ot::rmsnorm_bwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_ns_tokenizer_fwd(const ot_ns_tokenizer_params* p, void* stream) { return ot::ns_tokenizer_fwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_ns_tokenizer_bwd(const ot_ns_tokenizer_params* p, void* stream) { return ot::ns_tokenizer_bwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_fill_rows(const float* vec, void* out, int64_t ldo, int64_t row0, int64_t n_rows, int32_t d, void* stream) {
  return ot::fill_rows_impl(vec, out, ldo, row0, n_rows, d, static_cast<cudaStream_t>(stream));
}
int ot_colsum(const ot_colsum_params* p, void* stream) { return ot::colsum_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_dropout_mask(const void* in, int64_t ld_in, void* out, int64_t ld_out, int64_t rows, int32_t cols, uint32_t seed, float rate,
                    void* stream) {
  return ot::dropout_mask_impl(in, ld_in, out, ld_out, rows, cols, seed, rate, static_cast<cudaStream_t>(stream));
}
int ot_clip_rmsprop_step(const ot_rmsprop_params* p, void* stream) { return ot::clip_rmsprop_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_embed_gather_fwd(const ot_embed_params* p, void* stream) { return ot::embed_gather_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_embed_scatter_bwd(const ot_embed_params* p, void* stream) { return ot::embed_scatter_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_embed_adagrad_step(const ot_embed_params* p, void* stream) { return ot::embed_adagrad_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_heads_fwd(const ot_heads_params* p, void* stream) { return ot::heads_fwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_heads_bwd(const ot_heads_params* p, void* stream) { return ot::heads_bwd_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_metrics_update(const ot_metrics_params* p, void* stream) { return ot::metrics_update_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_metrics_result(const ot_metrics_params* p, void* stream) { return ot::metrics_result_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_auc_pack_keys(const ot_auc_params* p, void* stream) { return ot::auc_pack_impl(p, static_cast<cudaStream_t>(stream)); }
int ot_auc_ranksum(const ot_auc_params* p, void* stream) { return ot::auc_ranksum_impl(p, static_cast<cudaStream_t>(stream)); }

}  // extern "C"
