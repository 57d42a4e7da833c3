// Host-side internals shared by the translation units of libgenima_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <array>
#include <utility>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/genima_b200.h"

struct gn_prof_rec {
  cudaEvent_t a, b;
  int cls;
  double flops, bytes;
};

struct gn_handle {
  int device = 0;
  // per-call CUDA-event profiling (gn_profile_begin / gn_profile_end); never active during graph capture
  bool profiling = false;
  std::vector<gn_prof_rec> prof;
  std::vector<cudaEvent_t> event_pool;
  int num_sms = 148;
  char err[512] = {0};
  void* workspace = nullptr;
  int64_t workspace_bytes = 0;
  void* gemm_trace = nullptr;  // device uint64[8]: phase timestamps of CTA (0,0,0) of the next GEMM launches
  int force_block_n = 0;
  int force_splits = 0;
  int force_occupancy = 0;
  int pair_mode = 1;    // CTA pairs (cta_group::2, M = 256 MMAs): 0 never, 1 a candidate of the tile search, 2 wherever possible
  int last_pair = 0;    // the last GEMM-class launch used CTA pairs
  // measured tile configurations per problem shape (gn_set_autotune): key -> {block_n, splits, stages, tmem_cols}
  bool autotune = false;
  std::unordered_map<std::string, std::array<int, 4>> tune_cache;
  cudaEvent_t tune_ev[2] = {nullptr, nullptr};
  int32_t last_cfg[4] = {0, 0, 0, 0};
  int last_rowstats_parts = 0;
  int64_t launches = 0;
  bool gemm_attr_set = false;
  void* stats_scratch = nullptr;  // GroupNorm: grid-barrier words (first 256 B, zero-initialised) + per-CTA partials
  int64_t stats_scratch_bytes = 0;
  bool attn_attr_set = false;
  int attn_kv_split = 1;  // 0: never, 1: when it shortens the critical path (default), 2: whenever Tk spans >= 2 blocks
  bool gn_attr_set = false;
  bool gna_attr_set = false;
  int gn_max_ctas = 0;  // 0: one CTA per SM
  bool pdl = true;      // launch the main kernels with programmatic stream serialization (gn_set_pdl)
  bool half_a_box = true;       // 64-row A boxes / 8 KiB A stages when an m-tile has at most 64 real rows
  bool w_prefetch = true;       // GEMM producers request the first W tiles before griddepcontrol.wait (gn_set_pdl(h, 2) = off)
  bool staged_epilogue = true;  // GEMM outputs staged in shared memory and TMA-stored (gn_set_staged_epilogue)
  bool fast_epilogue = true;    // compact per-activation kernel flavours (gn_set_staged_epilogue(h, 2) = staged, generic)
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

// fp16 tiled tensor map, SWIZZLE_128B by default (64 / 32 for narrower inner boxes), zero OOB fill.  dims/strides
// innermost first; strides[i] (bytes) is the stride of dim i+1.  Returns 0 or a negative gn_status.
int make_tmap_f16(gn_handle* h, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

// RAII scope recording a CUDA event pair around one C-ABI call when profiling is on (cls: enum gn_prof_class).
struct ProfScope {
  gn_handle* h;
  cudaStream_t st;
  int idx = -1;
  ProfScope(gn_handle* h_, void* stream, int cls, double flops, double bytes) : h(h_), st((cudaStream_t)stream) {
    if (!h || !h->profiling) return;
    gn_prof_rec r;
    r.cls = cls;
    r.flops = flops;
    r.bytes = bytes;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
      if (!h->event_pool.empty()) {
        *e = h->event_pool.back();
        h->event_pool.pop_back();
      } else if (cudaEventCreate(e) != cudaSuccess) {
        cudaGetLastError();
        return;
      }
    }
    cudaEventRecord(r.a, st);
    h->prof.push_back(r);
    idx = (int)h->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof[idx].b, st);
  }
};

// Launch through cudaLaunchKernelEx with the programmatic-dependent-launch attribute (when enabled on the handle) and an
// optional thread-block cluster along grid.z.  Every kernel launched through this helper calls pdl_wait() before its
// first global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(const gn_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, int cluster_z, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  // cluster_z: low byte = cluster size along grid.z, next byte = along grid.y, third byte = along grid.x (0 / 1 = none)
  const int cz = (cluster_z & 0xff) > 1 ? (cluster_z & 0xff) : 1;
  const int cy = ((cluster_z >> 8) & 0xff) > 1 ? ((cluster_z >> 8) & 0xff) : 1;
  const int cx = ((cluster_z >> 16) & 0xff) > 1 ? ((cluster_z >> 16) & 0xff) : 1;
  if (cz > 1 || cy > 1 || cx > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cx;
    attr[na].val.clusterDim.y = cy;
    attr[na].val.clusterDim.z = cz;
    ++na;
  }
  if (h->pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace gn
