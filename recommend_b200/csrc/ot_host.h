// ot_host.h — host-side helpers shared by the launchers: error slot, driver entry point for
// cuTensorMapEncodeTiled (resolved through the runtime, so the library does not link libcuda),
// tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace ot {

// ---- error reporting (C-ABI: return codes + ot_last_error_string) -------------------------------
enum : int {
  OT_OK = 0,
  OT_ERR_INVALID_ARG = -1,
  OT_ERR_UNSUPPORTED_SHAPE = -2,
  OT_ERR_DRIVER = -3,
  OT_ERR_CUDA_BASE = -1000  // -(1000 + cudaError_t)
};

char* error_slot();  // thread-local buffer, defined in ot_api.cu
#define OT_FAIL(code, ...)                                   \
  do {                                                       \
    snprintf(::ot::error_slot(), 512, __VA_ARGS__);          \
    return (code);                                           \
  } while (0)
#define OT_CUDA_CHECK(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      snprintf(::ot::error_slot(), 512, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
               __FILE__, __LINE__);                                                                  \
      return ::ot::OT_ERR_CUDA_BASE - static_cast<int>(_e);                                          \
    }                                                                                                \
  } while (0)

// ---- tensor maps ---------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled get_encode_fn();  // defined in ot_api.cu

// bf16 tensor map of rank <= 3.  dims[0] is the contiguous dimension; strides_bytes[i] is the byte
// stride of dims[i+1].  swizzle_bytes: 0, 64 or 128.  Returns 0 on success.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  PFN_cuTensorMapEncodeTiled fn = get_encode_fn();
  if (!fn) OT_FAIL(OT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[3] = {1, 1, 1};
  cuuint64_t gstr[2] = {0, 0};
  cuuint32_t bx[3] = {1, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    OT_FAIL(OT_ERR_DRIVER,
            "cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=[%llu,%llu,%llu] strides=[%llu,%llu] "
            "box=[%u,%u,%u] swz=%d",
            static_cast<int>(r), base, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
            (unsigned long long)gdim[2], (unsigned long long)gstr[0], (unsigned long long)gstr[1], bx[0], bx[1],
            bx[2], swizzle_bytes);
  }
  return OT_OK;
}

int num_sms();  // cached SM count of the current device (ot_api.cu)
int* sched_slot(cudaStream_t st);  // zeroed work counter for one persistent-kernel launch, or NULL (static schedule)

}  // namespace ot
