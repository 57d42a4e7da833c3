// Host-side internals shared by the translation units of libgenima_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>

#include "../../include/genima_b200.h"

struct gn_handle {
  int device = 0;
  int num_sms = 148;
  char err[512] = {0};
  void* workspace = nullptr;
  int64_t workspace_bytes = 0;
  int force_block_n = 0;
  int force_splits = 0;
  int32_t last_cfg[4] = {0, 0, 0, 0};
  int64_t launches = 0;
  bool gemm_attr_set = false;
  void* stats_scratch = nullptr;  // GroupNorm (sum, sumsq) accumulators, owned by the handle
  int64_t stats_scratch_bytes = 0;
  bool attn_attr_set = false;
  // cuTensorMapEncodeTiled resolved at runtime through cudaGetDriverEntryPoint (the library must load on a
  // CPU-only box, so libcuda is never linked directly).
  void* encode_fn = nullptr;
  // tensor-map cache: descriptor encode costs ~1 us, but identical maps recur every denoise step.
  std::unordered_map<std::string, CUtensorMap> tmap_cache;
};

namespace gn {

int set_error(gn_handle* h, int code, const char* fmt, ...);

#define GN_CHECK_ARG(h, cond, ...)                                  \
  do {                                                              \
    if (!(cond)) return gn::set_error((h), GN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define GN_CHECK_CUDA(h, expr)                                                                        \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return gn::set_error((h), GN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                       \
  } while (0)

#define GN_CHECK_LAUNCH(h)                                                                                   \
  do {                                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                                     \
    if (_e != cudaSuccess)                                                                                   \
      return gn::set_error((h), GN_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                           __FILE__, __LINE__);                                                              \
    (h)->launches++;                                                                                         \
  } while (0)

// fp16 tiled tensor map, SWIZZLE_128B, zero OOB fill.  dims/strides innermost first; strides[i] (bytes) is the stride
// of dim i+1.  Returns 0 or a negative gn_status.
int make_tmap_f16(gn_handle* h, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace gn
