// Handle lifecycle, error reporting and TMA descriptor construction for libgenima_b200.so.
#include <stdarg.h>

#include <algorithm>

#include "common.h"

namespace gn {

int set_error(gn_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_f16(gn_handle* h, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  if (!h->encode_fn) return set_error(h, GN_ERR_NODRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  // cache key: raw bytes of every argument
  std::string key;
  key.reserve(8 + rank * 20);
  key.append(reinterpret_cast<const char*>(&base), sizeof(base));
  key.append(reinterpret_cast<const char*>(&rank), sizeof(rank));
  key.append(reinterpret_cast<const char*>(dims), sizeof(uint64_t) * rank);
  key.append(reinterpret_cast<const char*>(strides_bytes), sizeof(uint64_t) * (rank - 1));
  key.append(reinterpret_cast<const char*>(box), sizeof(uint32_t) * rank);
  key.append(reinterpret_cast<const char*>(&swizzle_bytes), sizeof(swizzle_bytes));
  auto it = h->tmap_cache.find(key);
  if (it != h->tmap_cache.end()) {
    *out = it->second;
    return GN_OK;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(h, GN_ERR_INVALID, "TMA base address %p is not 16-byte aligned", base);
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error(h, GN_ERR_INVALID, "TMA box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0)
      return set_error(h, GN_ERR_INVALID, "TMA stride[%d]=%llu bytes is not a multiple of 16", i,
                       (unsigned long long)strides_bytes[i]);
  }
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (swizzle_bytes == 64) swz = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 32) swz = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes != 128) return set_error(h, GN_ERR_INVALID, "TMA swizzle %d unsupported", swizzle_bytes);
  if ((uint64_t)box[0] * 2 > (uint64_t)swizzle_bytes)
    return set_error(h, GN_ERR_INVALID, "TMA inner box (%u elements) exceeds the %d-byte swizzle span", box[0],
                     swizzle_bytes);
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_fn)(
      out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides,
      gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(h, GN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  if (h->tmap_cache.size() > 65536) h->tmap_cache.clear();
  h->tmap_cache.emplace(std::move(key), *out);
  return GN_OK;
}

}  // namespace gn

extern "C" {

const char* gn_version(void) { return "genima_b200 0.1 (sm_100a)"; }

int gn_create(int device, gn_handle** out) {
  if (!out) return GN_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return GN_ERR_NODRIVER;
  }
  if (device < 0 || device >= count) return GN_ERR_INVALID;
  gn_handle* h = new (std::nothrow) gn_handle();
  if (!h) return GN_ERR_NOMEM;
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete h;
    return GN_ERR_CUDA;
  }
  if (prop.major != 10) {
    delete h;
    return GN_ERR_INVALID;  // sm_100a only: the kernels use tcgen05/TMEM
  }
  h->num_sms = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete h;
    return GN_ERR_CUDA;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    delete h;
    return GN_ERR_NODRIVER;
  }
  h->encode_fn = fn;
  h->stats_scratch_bytes = 256 * 1024;
  if (cudaMalloc(&h->stats_scratch, h->stats_scratch_bytes) != cudaSuccess ||
      cudaMemset(h->stats_scratch, 0, h->stats_scratch_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete h;
    return GN_ERR_NOMEM;
  }
  *out = h;
  return GN_OK;
}

int gn_destroy(gn_handle* h) {
  if (h) {
    for (auto& r : h->prof) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (auto e : h->tune_ev)
      if (e) cudaEventDestroy(e);
  }
  if (h && h->stats_scratch) cudaFree(h->stats_scratch);
  delete h;
  return GN_OK;
}

const char* gn_last_error(const gn_handle* h) { return h ? h->err : "null handle"; }

int gn_set_workspace(gn_handle* h, void* dptr, int64_t bytes) {
  if (!h) return GN_ERR_INVALID;
  h->workspace = dptr;
  h->workspace_bytes = bytes;
  return GN_OK;
}

int gn_set_gemm_tuning(gn_handle* h, int block_n, int splits) {
  if (!h) return GN_ERR_INVALID;
  h->force_block_n = block_n;
  h->force_splits = splits;
  return GN_OK;
}

int gn_set_gemm_pair(gn_handle* h, int mode) {
  if (!h || mode < 0 || mode > 2) return GN_ERR_INVALID;
  h->pair_mode = mode;
  h->tune_cache.clear();
  return GN_OK;
}

int gn_last_gemm_pair(const gn_handle* h) { return h ? h->last_pair : 0; }

int gn_set_gemm_occupancy(gn_handle* h, int ctas_per_sm) {
  if (!h || ctas_per_sm < 0 || ctas_per_sm > 2) return GN_ERR_INVALID;
  h->force_occupancy = ctas_per_sm;
  return GN_OK;
}

int gn_set_attention_kv_split(gn_handle* h, int mode) {
  if (!h || mode < 0 || mode > 2) return GN_ERR_INVALID;
  h->attn_kv_split = mode;
  return GN_OK;
}

int gn_set_pdl(gn_handle* h, int enable) {
  if (!h) return GN_ERR_INVALID;
  h->pdl = enable != 0;
  h->w_prefetch = enable == 1;
  return GN_OK;
}

int gn_set_staged_epilogue(gn_handle* h, int enable) {
  if (!h) return GN_ERR_INVALID;
  h->staged_epilogue = enable != 0;
  h->fast_epilogue = enable == 1;  // 2: staged output through the generic kernel flavour (A/B)
  h->tune_cache.clear();
  return GN_OK;
}

int gn_set_gn_max_ctas(gn_handle* h, int max_ctas) {
  if (!h || max_ctas < 0) return GN_ERR_INVALID;
  h->gn_max_ctas = max_ctas;
  return GN_OK;
}

int gn_set_autotune(gn_handle* h, int enable) {
  if (!h) return GN_ERR_INVALID;
  h->autotune = enable != 0;
  if (enable < 0) h->tune_cache.clear();
  return GN_OK;
}

// Measured tile configurations as text, one "key=block_n,splits,stages,packed" line per problem shape: lets rank 0 tune
// once and every other rank / handle import the same choices, so that all ranks sum in the same order (sharding an
// evaluation must not change an episode's result).  Returns the number of bytes the export needs (call with cap = 0
// to size the buffer); negative on error.
int64_t gn_tune_cache_export(const gn_handle* h, char* buf, int64_t cap) {
  if (!h) return GN_ERR_INVALID;
  std::vector<std::string> lines;
  lines.reserve(h->tune_cache.size());
  for (const auto& kv : h->tune_cache) {
    char tmp[192];
    snprintf(tmp, sizeof(tmp), "%s=%d,%d,%d,%d\n", kv.first.c_str(), kv.second[0], kv.second[1], kv.second[2],
             kv.second[3]);
    lines.emplace_back(tmp);
  }
  std::sort(lines.begin(), lines.end());
  int64_t need = 0;
  for (const auto& l : lines) need += (int64_t)l.size();
  if (buf && cap >= need) {
    char* p = buf;
    for (const auto& l : lines) {
      memcpy(p, l.data(), l.size());
      p += l.size();
    }
  }
  return need;
}

int gn_tune_cache_import(gn_handle* h, const char* buf, int64_t n, int replace) {
  if (!h || (!buf && n > 0) || n < 0) return GN_ERR_INVALID;
  if (replace) h->tune_cache.clear();
  const char* p = buf;
  const char* end = buf + n;
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', end - p));
    const char* le = nl ? nl : end;
    const char* eq = static_cast<const char*>(memchr(p, '=', le - p));
    if (eq) {
      int v[4];
      std::string val(eq + 1, le);
      if (sscanf(val.c_str(), "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) != 4)
        return gn::set_error(h, GN_ERR_INVALID, "gn_tune_cache_import: malformed line");
      h->tune_cache[std::string(p, eq)] = {v[0], v[1], v[2], v[3]};
    }
    p = le + 1;
  }
  return GN_OK;
}

int gn_set_gemm_trace(gn_handle* h, void* dptr_u64x8) {
  if (!h) return GN_ERR_INVALID;
  h->gemm_trace = dptr_u64x8;
  return GN_OK;
}

int gn_get_last_gemm_config(const gn_handle* h, int32_t* out4) {
  if (!h || !out4) return GN_ERR_INVALID;
  for (int i = 0; i < 4; ++i) out4[i] = h->last_cfg[i];
  return GN_OK;
}

int gn_get_last_rowstats_parts(const gn_handle* h) { return h ? h->last_rowstats_parts : -1; }

int64_t gn_launch_count(const gn_handle* h) { return h ? h->launches : -1; }

int gn_profile_begin(gn_handle* h) {
  if (!h) return GN_ERR_INVALID;
  for (auto& r : h->prof) {
    h->event_pool.push_back(r.a);
    h->event_pool.push_back(r.b);
  }
  h->prof.clear();
  h->profiling = true;
  return GN_OK;
}

int gn_profile_end(gn_handle* h, double* ms, int64_t* calls, double* flops, double* bytes) {
  if (!h || !ms || !calls || !flops || !bytes) return GN_ERR_INVALID;
  h->profiling = false;
  for (int c = 0; c < GN_PROF_NUM_CLASSES; ++c) {
    ms[c] = 0.0;
    calls[c] = 0;
    flops[c] = 0.0;
    bytes[c] = 0.0;
  }
  int rc = GN_OK;
  for (auto& r : h->prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) {
      cudaGetLastError();
      rc = gn::set_error(h, GN_ERR_CUDA, "gn_profile_end: event timing failed (profiling inside graph capture?)");
    } else if (r.cls >= 0 && r.cls < GN_PROF_NUM_CLASSES) {
      ms[r.cls] += t;
      calls[r.cls] += 1;
      flops[r.cls] += r.flops;
      bytes[r.cls] += r.bytes;
    }
    h->event_pool.push_back(r.a);
    h->event_pool.push_back(r.b);
  }
  h->prof.clear();
  return rc;
}

}  // extern "C"
