// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[M, N] = epilogue( sum_k A[m, k] * W[n, k] )        fp16 operands, fp32 accumulation in TMEM
//
// One CTA computes a 128 x block_n output tile (block_n chosen per problem, multiple of 16, <= 256):
//   warp 0      TMA producer: A tile [128 rows][64 fp16] and W tile [block_n rows][64 fp16] per k-block,
//               SWIZZLE_128B, multi-stage ring guarded by full/empty mbarriers
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 MMAs per k-block),
//               tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld (one TMEM lane = one output row per thread) -> fused
//               scale/bias/time-embedding/activation/residual/GEGLU -> fp16 stores
//
// The A operand is either a plain row-major matrix (2D tensor map) or an NHWC image addressed through 4D tensor
// maps: k-blocks walk a table of "segments" (tap (dy, dx) x 64-channel blocks); TMA's zero OOB fill implements the
// convolution padding, per-phase tensor maps implement stride 2, and extra 1x1 segments fuse ResnetBlock2D's
// conv_shortcut into the same accumulation.  Split-K (gridDim.z) covers the weight-streaming-bound 8x8/16x16 levels.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace gn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int MAX_SEGS = 52;
constexpr int EPI_COLSPLIT = 2;                        // epilogue warps per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + 128 * EPI_COLSPLIT;  // producer warp + MMA warp + epilogue warps

struct KSeg {
  int8_t map;  // index into tmA
  int8_t dy;
  int8_t dx;
  int8_t _pad;
  int32_t nblk;  // number of 64-channel k-blocks in this segment
};

struct EpiParams {
  __half* out;
  float* out32;
  int64_t ldo;
  const float* scale;
  const float* bias;
  const float* rowvec;
  const __half* residual;
  int64_t ldr;
  int rows_per_batch;
  int act_pre, act_post;
  float alpha, beta;
  int geglu;
  int M, N;  // N = accumulator columns (before GEGLU halving)
  // LayerNorm folded into this GEMM (A = un-normalised x, W pre-multiplied by gamma):
  //   out = rstd[m] * (acc - mean[m] * colsum[n]) + bias[n],  (mean, rstd) from the row partials the producer of x wrote
  const float2* ln_stats;  // [M][ln_parts] (sum, sumsq) partials of each row of x; NULL = no folded LayerNorm
  int ln_parts;
  int ln_dim;              // row length of x (= K)
  float ln_eps;
  // row statistics of THIS GEMM's fp16 output for a LayerNorm folded into its consumer: [M][rs_parts] partials
  float2* rs_out;
  int rs_parts;
};

struct GemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  EpiParams epi;
  int num_kblocks;
  int kb_per_split;
  int splits;
  int mode;  // 0: 2D A, 1: NHWC conv A
  int num_segs;
  int Ho, Wo, Bn;
  int bw, bh, bb;
  int tiles_w, tiles_h;
  int block_n, stages, tmem_cols;
  float* ws;  // split-K partials [splits][M][N] fp32
  int rs_capacity;            // host-side: capacity (partials per row) of epi.rs_out
  unsigned long long* trace;  // optional: %globaltimer stamps of CTA (0,0,0)'s phases (gn_set_gemm_trace)
  KSeg segs[MAX_SEGS];
};

__device__ __forceinline__ void trace_stamp(const GemmParams& p, int slot) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case GN_ACT_SILU: return silu_f(v);
    case GN_ACT_GELU: return gelu_erf_f(v);
    case GN_ACT_RELU: return fmaxf(v, 0.0f);
    case GN_ACT_QUICKGELU: return quick_gelu_f(v);
    default: return v;
  }
}
template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if constexpr (ACT == GN_ACT_SILU) return silu_f(v);
  else if constexpr (ACT == GN_ACT_GELU) return gelu_erf_f(v);
  else if constexpr (ACT == GN_ACT_RELU) return fmaxf(v, 0.0f);
  else if constexpr (ACT == GN_ACT_QUICKGELU) return quick_gelu_f(v);
  else return v;
}

// v = act_pre(acc * scale[n] + bias[n] + rowvec[b, n])   (runtime-dispatched form, used by the split-K reduce pass)
__device__ __forceinline__ float epi_pre(const EpiParams& e, float acc, int n, int b) {
  float v = acc;
  if (e.scale) v *= __ldg(e.scale + n);
  if (e.bias) v += __ldg(e.bias + n);
  if (e.rowvec) v += __ldg(e.rowvec + (int64_t)b * e.N + n);
  return apply_act(v, e.act_pre);
}

// Finalise CH consecutive output columns [nout, nout + CH) of row m from pre-activation values v[]:
// out = act_post(alpha * v + beta * residual).  16-byte vector path when the chunk is full and aligned.
template <int CH>
__device__ __forceinline__ void epi_store(const EpiParams& e, float (&v)[CH], int m, int nout, int n_out_total,
                                          float* rs = nullptr) {
  const bool full = (nout + CH <= n_out_total);
  if (e.residual) {
    const __half* rp = e.residual + (int64_t)m * e.ldr + nout;
    if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < CH; j += 8) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + j));
        const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 f = __half22float2(hp[t]);
          v[j + 2 * t] = fmaf(e.alpha, v[j + 2 * t], e.beta * f.x);
          v[j + 2 * t + 1] = fmaf(e.alpha, v[j + 2 * t + 1], e.beta * f.y);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float r = (nout + j < n_out_total) ? __half2float(rp[j]) : 0.0f;
        v[j] = e.alpha * v[j] + e.beta * r;
      }
    }
  } else if (e.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] *= e.alpha;
  }
  if (e.act_post == GN_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (e.act_post != GN_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = apply_act(v[j], e.act_post);
  }
  if (e.out32) {
    float* op = e.out32 + (int64_t)m * e.ldo + nout;
    if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < CH; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (nout + j < n_out_total) op[j] = v[j];
    }
    return;
  }
  __half* op = e.out + (int64_t)m * e.ldo + nout;
  if (rs) {
    // statistics of the values the consumer will read back: the fp16-rounded outputs
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const float h = (nout + j < n_out_total) ? __half2float(__float2half_rn(v[j])) : 0.f;
      rs[0] += h;
      rs[1] = fmaf(h, h, rs[1]);
    }
  }
  if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < CH; j += 8) {
      uint4 q;
      q.x = pack_half2(v[j], v[j + 1]);
      q.y = pack_half2(v[j + 2], v[j + 3]);
      q.z = pack_half2(v[j + 4], v[j + 5]);
      q.w = pack_half2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(op + j) = q;
    }
  } else {
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (nout + j < n_out_total) op[j] = __float2half_rn(v[j]);
  }
}

// (rstd, -mean * rstd) of row m of the GEMM's A operand from the row partials its producer wrote (fixed summation order).
__device__ __forceinline__ void ln_row(const EpiParams& e, int m, float& rstd, float& nmr) {
  const float2* sp = e.ln_stats + (int64_t)m * e.ln_parts;
  float s1 = 0.f, s2 = 0.f;
  int i = 0;
  for (; i + 8 <= e.ln_parts; i += 8) {  // eight independent loads in flight, summed in index order
    float2 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = __ldg(sp + i + u);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s1 += t[u].x;
      s2 += t[u].y;
    }
  }
  for (; i < e.ln_parts; ++i) {
    const float2 t = __ldg(sp + i);
    s1 += t.x;
    s2 += t.y;
  }
  const float inv = 1.0f / (float)e.ln_dim;
  const float mean = s1 * inv;
  const float var = fmaxf(s2 * inv - mean * mean, 0.f);
  rstd = rsqrtf(var + e.ln_eps);
  nmr = -mean * rstd;
}

// Epilogue of one output row over this warp's share of the tile's 16-column chunks (chunk index cw, cw + EPI_COLSPLIT, ...).
// s_scale / s_bias: per-column fp32 vectors of the tile staged in shared memory (scale = 1 / bias = 0 when absent).
// Rolled loop over chunks with a compile-time activation: the body stays small enough for the instruction cache
// (a fully unrolled, runtime-dispatched epilogue measured ~60 instructions per element and was fetch-bound).
template <int ACT, bool CLUSTER>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, uint32_t taddr, int n0, int m, int b, bool valid,
                                              int cw, const float* s_scale, const float* s_bias, const float* stage,
                                              int row, float ln_rstd, float ln_nmr) {
  const EpiParams& e = p.epi;
  const int nchunks = p.block_n >> 4;
  // CLUSTER (split-K): this CTA finishes the chunks rank, rank + S, ... of the tile, summing the fp32 partials that all
  // S CTAs of the cluster staged in their shared memory (read through DSMEM in rank order: deterministic).
  const int first = CLUSTER ? (int)cg::this_cluster().block_rank() + p.splits * cw : cw;
  const int step = CLUSTER ? p.splits * EPI_COLSPLIT : EPI_COLSPLIT;
  float rs[2] = {0.f, 0.f};
#pragma unroll 1
  for (int ch = first; ch < nchunks; ch += step) {
    const int c = ch << 4;
    float v[16];
    if constexpr (CLUSTER) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
      for (int sr = 0; sr < p.splits; ++sr) {
        const float* peer = cg::this_cluster().map_shared_rank(stage, sr) + (size_t)c * BLOCK_M + row;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += peer[j * BLOCK_M];
      }
    } else {
      uint32_t r[16];
      tmem_ld_x16(taddr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    }
    const int n = n0 + c;
    if (!valid || n >= e.N) continue;
    if (e.ln_stats) {
      // s_scale holds colsum[n] = sum_k gamma[k] W[n, k], s_bias holds bias[n] + sum_k beta[k] W[n, k]
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 cs = *reinterpret_cast<const float4*>(s_scale + c + j);
        const float4 bi = *reinterpret_cast<const float4*>(s_bias + c + j);
        v[j] = fmaf(v[j], ln_rstd, fmaf(ln_nmr, cs.x, bi.x));
        v[j + 1] = fmaf(v[j + 1], ln_rstd, fmaf(ln_nmr, cs.y, bi.y));
        v[j + 2] = fmaf(v[j + 2], ln_rstd, fmaf(ln_nmr, cs.z, bi.z));
        v[j + 3] = fmaf(v[j + 3], ln_rstd, fmaf(ln_nmr, cs.w, bi.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c + j);
        const float4 bi = *reinterpret_cast<const float4*>(s_bias + c + j);
        v[j] = fmaf(v[j], sc.x, bi.x);
        v[j + 1] = fmaf(v[j + 1], sc.y, bi.y);
        v[j + 2] = fmaf(v[j + 2], sc.z, bi.z);
        v[j + 3] = fmaf(v[j + 3], sc.w, bi.w);
      }
    }
    if (e.rowvec) {
      const float* rv = e.rowvec + (int64_t)b * e.N + n;
      if (n + 16 <= e.N && ((reinterpret_cast<uintptr_t>(rv) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(rv + j));
          v[j] += t.x;
          v[j + 1] += t.y;
          v[j + 2] += t.z;
          v[j + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < e.N) v[j] += __ldg(rv + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_ct<ACT>(v[j]);
    epi_store<16>(e, v, m, n, e.N, e.rs_out ? rs : nullptr);
  }
  if (e.rs_out && valid) {
    // one partial per (n-tile, K-split rank, column share); the consumer sums them in index order
    const int rank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int part = ((int)blockIdx.x * p.splits + rank) * EPI_COLSPLIT + cw;
    e.rs_out[(int64_t)m * e.rs_parts + part] = make_float2(rs[0], rs[1]);
  }
}

template <bool CLUSTER>
__device__ __forceinline__ void epilogue_dispatch(const GemmParams& p, uint32_t taddr, int n0, int m, int b, bool valid,
                                                  int cw, const float* s_scale, const float* s_bias,
                                                  const float* stage, int row, float ln_rstd, float ln_nmr) {
  switch (p.epi.act_pre) {
    case GN_ACT_SILU:
      epilogue_rows<GN_ACT_SILU, CLUSTER>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, stage, row, ln_rstd,
                                          ln_nmr);
      break;
    case GN_ACT_GELU:
      epilogue_rows<GN_ACT_GELU, CLUSTER>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, stage, row, ln_rstd,
                                          ln_nmr);
      break;
    case GN_ACT_RELU:
      epilogue_rows<GN_ACT_RELU, CLUSTER>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, stage, row, ln_rstd,
                                          ln_nmr);
      break;
    case GN_ACT_QUICKGELU:
      epilogue_rows<GN_ACT_QUICKGELU, CLUSTER>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, stage, row, ln_rstd,
                                          ln_nmr);
      break;
    default:
      epilogue_rows<GN_ACT_NONE, CLUSTER>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, stage, row, ln_rstd,
                                          ln_nmr);
      break;
  }
}

// GEGLU epilogue: accumulator columns come in 128-wide groups [64 values | 64 gates]; out[m, j] = value * gelu(gate).
__device__ __forceinline__ void epilogue_rows_geglu(const GemmParams& p, uint32_t taddr, int n0, int m, bool valid,
                                                    int cw, const float* s_scale, const float* s_bias, float ln_rstd,
                                                    float ln_nmr) {
  const EpiParams& e = p.epi;
  const int n_out_total = e.N >> 1;
  const int nchunks = p.block_n >> 5;  // 16-wide value chunks: 4 per 128-column group
#pragma unroll 1
  for (int ch = cw; ch < nchunks; ch += EPI_COLSPLIT) {
    const int c = ((ch >> 2) << 7) + ((ch & 3) << 4);  // tile column of the value chunk
    uint32_t rv[16], rg[16];
    tmem_ld_x16(taddr + c, rv);
    tmem_ld_x16(taddr + c + 64, rg);
    tmem_ld_wait();
    const int n = n0 + c;
    if (!valid || n >= e.N) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // folded LayerNorm: acc * rstd - mean * rstd * colsum[n] + bias'[n]  (identity when rstd = 1, nmr = 0)
      const float a = fmaf(__uint_as_float(rv[j]), ln_rstd, fmaf(ln_nmr, s_scale[c + j], s_bias[c + j]));
      const float g = fmaf(__uint_as_float(rg[j]), ln_rstd, fmaf(ln_nmr, s_scale[c + 64 + j], s_bias[c + 64 + j]));
      v[j] = a * gelu_erf_f(g);
    }
    epi_store<16>(e, v, m, ((n0 + ((ch >> 2) << 7)) >> 1) + ((ch & 3) << 4), n_out_total);
  }
}

// Split-K partial: raw fp32 accumulators of this CTA's K-slice into its own shared memory, column-major
// [block_n][128 rows] (a warp writes 32 consecutive rows of one column: conflict-free), for the cluster reduction.
__device__ __forceinline__ void epilogue_rows_stage(const GemmParams& p, uint32_t taddr, int cw, float* stage, int row) {
  const int nchunks = p.block_n >> 4;
#pragma unroll 1
  for (int ch = cw; ch < nchunks; ch += EPI_COLSPLIT) {
    const int c = ch << 4;
    uint32_t r[16];
    tmem_ld_x16(taddr + c, r);
    tmem_ld_wait();
    float* dst = stage + (size_t)c * BLOCK_M + row;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j * BLOCK_M] = __uint_as_float(r[j]);
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int block_n = p.block_n;
  const int b_stage_bytes = block_n * BLOCK_K * 2;

  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * A_STAGE_BYTES;
  // the fp32 staging tile of the split-K cluster reduction aliases the operand ring and may be larger than it
  int ring_bytes = stages * (A_STAGE_BYTES + b_stage_bytes);
  if (p.splits > 1 && block_n * BLOCK_M * 4 > ring_bytes) ring_bytes = block_n * BLOCK_M * 4;
  float* s_scale = reinterpret_cast<float*>(smem + ring_bytes);
  float* s_bias = s_scale + 256;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int n0 = blockIdx.x * block_n;
  if (threadIdx.x == 0) trace_stamp(p, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Everything above touched only this CTA's shared / tensor memory.  Let the next kernel's CTAs be scheduled, then wait
  // for the previous kernel in the stream before the first global-memory access (programmatic dependent launch).
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(p, 1);

  const int mt = blockIdx.y;
  int m0 = 0, x0 = 0, y0 = 0, b0 = 0;
  if (p.mode == 0) {
    m0 = mt * BLOCK_M;
  } else {
    const int tw = mt % p.tiles_w;
    const int th = (mt / p.tiles_w) % p.tiles_h;
    const int tb = mt / (p.tiles_w * p.tiles_h);
    x0 = tw * p.bw;
    y0 = th * p.bh;
    b0 = tb * p.bb;
  }
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.num_kblocks, kb_begin + p.kb_per_split);
  const int num_it = kb_end - kb_begin;
  float ln_rstd = 1.f, ln_nmr = 0.f;  // folded LayerNorm of this thread's row (epilogue warps)

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int seg = 0, seg_start = 0;
      if (p.mode == 1) {
        while (kb_begin >= seg_start + p.segs[seg].nblk) {
          seg_start += p.segs[seg].nblk;
          ++seg;
        }
      }
      int cb = kb_begin - seg_start;
      const uint32_t tx_bytes = A_STAGE_BYTES + b_stage_bytes;
      for (int it = 0; it < num_it; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        const int kb = kb_begin + it;
        if (p.mode == 0) {
          tma_load_2d(smem_a + s * A_STAGE_BYTES, &p.tmA[0], &full_bar[s], kb * BLOCK_K, m0);
        } else {
          const KSeg sg = p.segs[seg];
          tma_load_4d(smem_a + s * A_STAGE_BYTES, &p.tmA[sg.map], &full_bar[s], cb * BLOCK_K, x0 + sg.dx, y0 + sg.dy,
                      b0);
          if (++cb == sg.nblk) {
            cb = 0;
            ++seg;
          }
        }
        tma_load_2d(smem_b + s * b_stage_bytes, &p.tmB, &full_bar[s], kb * BLOCK_K, n0);
        if (it == 0) trace_stamp(p, 2);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = umma_idesc_f16(block_n, 0, 0);
      for (int it = 0; it < num_it; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (it == 0) trace_stamp(p, 3);
        const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + s * A_STAGE_BYTES), 1024, 0);
        const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + s * b_stage_bytes), 1024, 0);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
      trace_stamp(p, 4);
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps (2 .. 2 + 4 * EPI_COLSPLIT)
    const int q = warp & 3;                // TMEM lane quadrant this warp may access
    const int cw = (warp - 2) >> 2;        // which share of the column chunks
    const int row = q * 32 + lane;
    int m;
    bool valid;
    if (p.mode == 0) {
      m = m0 + row;
      valid = m < p.epi.M;
    } else {
      const int x = row % p.bw;
      const int y = (row / p.bw) % p.bh;
      const int bb = row / (p.bw * p.bh);
      const int gx = x0 + x, gy = y0 + y, gb = b0 + bb;
      valid = (gx < p.Wo) && (gy < p.Ho) && (gb < p.Bn);
      m = (gb * p.Ho + gy) * p.Wo + gx;
    }
    const int b = (p.epi.rowvec && valid) ? (m / p.epi.rows_per_batch) : 0;
    // stage the tile's per-column scale / bias once (identity when absent) — read back as broadcast float4s
    for (int i = threadIdx.x - 64; i < block_n; i += GEMM_THREADS - 64) {
      const int n = n0 + i;
      s_scale[i] = (p.epi.scale && n < p.epi.N) ? __ldg(p.epi.scale + n) : 1.0f;
      s_bias[i] = (p.epi.bias && n < p.epi.N) ? __ldg(p.epi.bias + n) : 0.0f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GEMM_THREADS - 64) : "memory");  // epilogue warps only
    // folded LayerNorm: (rstd, -mean * rstd) of this row from the producer's partials, fetched while the MMAs run
    if (p.epi.ln_stats && valid) ln_row(p.epi, m, ln_rstd, ln_nmr);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(p, 5);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (p.splits > 1) {
      // all MMAs have completed (tmem_full), so the operand ring is free: reuse it as the fp32 staging tile
      epilogue_rows_stage(p, taddr, cw, reinterpret_cast<float*>(smem), row);
    } else if (p.epi.geglu) {
      epilogue_rows_geglu(p, taddr, n0, m, valid, cw, s_scale, s_bias, ln_rstd, ln_nmr);
    } else {
      epilogue_dispatch<false>(p, taddr, n0, m, b, valid, cw, s_scale, s_bias, nullptr, row, ln_rstd, ln_nmr);
    }
    tc_fence_before();
    if (threadIdx.x == 64) trace_stamp(p, 6);
  }
  if (p.splits > 1) {
    // ---- split-K reduction across the cluster (gridDim.z == cluster size): no workspace, no second kernel
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (warp >= 2) {
      const int q = warp & 3;
      const int cw = (warp - 2) >> 2;
      const int row = q * 32 + lane;
      int m;
      bool valid;
      if (p.mode == 0) {
        m = m0 + row;
        valid = m < p.epi.M;
      } else {
        const int x = row % p.bw;
        const int y = (row / p.bw) % p.bh;
        const int bb = row / (p.bw * p.bh);
        const int gx = x0 + x, gy = y0 + y, gb = b0 + bb;
        valid = (gx < p.Wo) && (gy < p.Ho) && (gb < p.Bn);
        m = (gb * p.Ho + gy) * p.Wo + gx;
      }
      const int b = (p.epi.rowvec && valid) ? (m / p.epi.rows_per_batch) : 0;
      epilogue_dispatch<true>(p, 0, n0, m, b, valid, cw, s_scale, s_bias, reinterpret_cast<const float*>(smem), row,
                              ln_rstd, ln_nmr);
    }
    cluster.sync();  // peers may still be reading this CTA's staging tile
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) trace_stamp(p, 7);
}

// ------------------------------------------------------------------------------------------------ host side

struct TileChoice {
  int block_n, splits, stages, tmem_cols;
};

static int smem_bytes_for(int block_n, int stages, int splits) {
  int ring = stages * (A_STAGE_BYTES + block_n * BLOCK_K * 2);
  const int stage_tile = block_n * BLOCK_M * 4;  // fp32 staging tile of the cluster split-K reduction (aliases the ring)
  if (splits > 1 && stage_tile > ring) ring = stage_tile;
  return ring + 2 * 256 * 4 + (2 * stages + 1) * 8 + 16 + 1024;
}

constexpr int MAX_CLUSTER_SPLITS = 8;  // portable cluster size limit

constexpr int SMEM_OCC1 = 200 * 1024;  // one CTA per SM: deep operand ring
constexpr int SMEM_OCC2 = 112 * 1024;  // two CTAs per SM: one CTA's epilogue / set-up overlaps the other's main loop

static int stages_for(int block_n, int splits, int kb_per, int budget) {
  int st = 8;
  while (st > 2 && smem_bytes_for(block_n, st, splits) > budget) --st;
  if (st > kb_per) st = kb_per < 2 ? 2 : kb_per;
  return st;
}

struct Candidate {
  TileChoice tc;
  double cost;
};

static int candidate_list(const gn_handle* h, int tiles_m, int N, int num_kblocks, bool geglu, bool allow_split,
                          Candidate* out, int max_out, int rs_capacity = 0) {
  static const int kCand[] = {256, 224, 192, 160, 128, 96, 80, 64, 48, 32, 16};
  const int sms = h->num_sms;
  std::vector<Candidate> all;
  for (int bn : kCand) {
    if (geglu && (bn % 128) != 0) continue;
    if (h->force_block_n && bn != h->force_block_n) continue;
    if (!h->force_block_n && bn > gn::round_up(N, 16)) continue;
    const int tiles_n = gn::ceil_div(N, bn);
    int max_splits = 1;
    if (allow_split && !geglu) {
      max_splits = num_kblocks / 4;  // keep >= 4 k-blocks per split
      if (max_splits < 1) max_splits = 1;
      if (max_splits > MAX_CLUSTER_SPLITS) max_splits = MAX_CLUSTER_SPLITS;
    }
    for (int sp = 1; sp <= max_splits; ++sp) {
      if (h->force_splits && sp != h->force_splits && !(h->force_splits > max_splits && sp == max_splits)) continue;
      const int kb_per = gn::ceil_div(num_kblocks, sp);
      if ((sp - 1) * kb_per >= num_kblocks) continue;  // an empty split
      if (rs_capacity > 0 && tiles_n * sp * EPI_COLSPLIT > rs_capacity) continue;  // row-statistics partials must fit
      const int64_t ctas = (int64_t)tiles_m * tiles_n * sp;
      for (int occ = 1; occ <= 2; ++occ) {
        if (h->force_occupancy && occ != h->force_occupancy) continue;
        if (occ == 2 && smem_bytes_for(bn, 2, sp) > SMEM_OCC2) continue;
        const double t_mma = 2.0 * bn;
        const double t_ld = (128.0 + bn) * 128.0 / 46.0;
        const double t_kb = t_mma > t_ld ? t_mma : t_ld;
        const double per_sm = (double)((ctas + sms - 1) / sms);                // main loops an SM runs back to back
        const double rounds = (double)((ctas + occ * sms - 1) / (occ * sms));  // exposed per-CTA latencies
        double lat = 25.0 * bn + 3000.0;
        if (sp > 1) lat += 2500.0 + 8.0 * bn;   // two cluster barriers + DSMEM reduction
        if (sp == 5 || sp == 7) lat += 2000.0;  // cluster sizes that pack the 18-SM GPCs badly
        const int st = stages_for(bn, sp, kb_per, occ == 2 ? SMEM_OCC2 : SMEM_OCC1);
        double mainloop = per_sm * kb_per * t_kb;
        if (occ * st < 4) mainloop *= 1.3;  // too few loads in flight to cover the TMA round trip
        Candidate c;
        c.tc.block_n = bn;
        c.tc.splits = sp;
        c.tc.stages = st;
        int tm = 32;
        while (tm < bn) tm <<= 1;
        c.tc.tmem_cols = tm;
        c.cost = mainloop + rounds * lat;
        bool dup = false;  // occupancy 1 / 2 sizing may give the same stage count
        for (const Candidate& o : all)
          if (o.tc.block_n == bn && o.tc.splits == sp && o.tc.stages == st) dup = true;
        if (!dup) all.push_back(c);
      }
    }
  }
  std::sort(all.begin(), all.end(), [](const Candidate& a, const Candidate& b) { return a.cost < b.cost; });
  // best by the model first; keep the list diverse (at most 3 entries per tile width) so that a model error on one
  // axis cannot hide the real optimum from the measurement
  int n_out = 0;
  for (const Candidate& c : all) {
    if (n_out >= max_out) break;
    int same_bn = 0;
    for (int i = 0; i < n_out; ++i) same_bn += out[i].tc.block_n == c.tc.block_n;
    if (same_bn >= 3) continue;
    out[n_out++] = c;
  }
  return n_out;
}

// Pick (block_n, splits, CTAs per SM) minimising a simple model of the kernel time (cycles), calibrated on
// gn_set_gemm_trace timelines:
//   * operands reach an SM through TMA at ~46 B/clk, so a 64-deep k-block costs max(2 * bn [tensor issue],
//     (128 + bn) * 128 / 46 [operand feed]) cycles of that SM, whichever CTA it belongs to;
//   * every CTA pays ~3000 clk of latency (set-up, first TMA round trip, exit) plus ~25 clk per accumulator column of
//     epilogue; with two co-resident CTAs that latency overlaps the neighbour's main loop.
static TileChoice choose_tiles(const gn_handle* h, int tiles_m, int N, int num_kblocks, bool geglu, bool allow_split,
                               int rs_capacity) {
  Candidate c[1];
  if (candidate_list(h, tiles_m, N, num_kblocks, geglu, allow_split, c, 1, rs_capacity) < 1)
    return TileChoice{128, 1, 4, 128};
  return c[0].tc;
}

static int fill_epilogue(gn_handle* h, EpiParams& e, const gn_epilogue* epi, void* out, int64_t ldo, int M, int N,
                         int default_rows_per_batch) {
  memset(&e, 0, sizeof(e));
  e.M = M;
  e.N = N;
  e.ldo = ldo;
  e.alpha = 1.0f;
  e.beta = 1.0f;
  e.rows_per_batch = default_rows_per_batch > 0 ? default_rows_per_batch : M;
  e.out = static_cast<__half*>(out);
  if (epi) {
    e.scale = epi->scale;
    e.bias = epi->bias;
    e.rowvec = epi->rowvec;
    e.residual = static_cast<const __half*>(epi->residual);
    e.ldr = epi->ldr;
    if (epi->rows_per_batch > 0) e.rows_per_batch = epi->rows_per_batch;
    e.act_pre = epi->act_pre;
    e.act_post = epi->act_post;
    e.alpha = epi->alpha;
    e.beta = epi->beta;
    e.geglu = epi->geglu;
    if (epi->out_fp32) {
      e.out32 = static_cast<float*>(out);
      e.out = nullptr;
    }
    if (epi->ln_stats) {
      GN_CHECK_ARG(h, epi->ln_colsum && epi->ln_parts > 0 && !epi->scale,
                   "folded LayerNorm needs ln_colsum, ln_parts > 0 and no scale vector");
      e.ln_stats = static_cast<const float2*>(epi->ln_stats);
      e.ln_parts = epi->ln_parts;
      e.ln_eps = epi->ln_eps;
      e.scale = epi->ln_colsum;  // staged in the per-column "scale" slot of the epilogue
    }
    if (epi->rowstats_out) {
      GN_CHECK_ARG(h, !epi->geglu && !epi->out_fp32 && epi->rowstats_capacity > 0,
                   "rowstats_out needs a plain fp16 output and a capacity");
      e.rs_out = static_cast<float2*>(epi->rowstats_out);
      e.rs_parts = epi->rowstats_capacity;  // replaced by the real partial count once the tile config is known
    }
    GN_CHECK_ARG(h, !(epi->geglu && (N % 128) != 0), "GEGLU needs N %% 128 == 0 (got %d)", N);
    GN_CHECK_ARG(h, !(epi->residual && epi->ldr <= 0), "residual given without ldr");
  }
  return GN_OK;
}

static int launch_config(gn_handle* h, GemmParams& p, const TileChoice& tc, int tiles_m, const void* W, int64_t ktot,
                         cudaStream_t stream) {
  const int N = p.epi.N;
  p.block_n = tc.block_n;
  p.splits = tc.splits;
  p.stages = tc.stages;
  p.tmem_cols = tc.tmem_cols;
  p.kb_per_split = gn::ceil_div(p.num_kblocks, tc.splits);
  p.ws = nullptr;
  p.trace = static_cast<unsigned long long*>(h->gemm_trace);
  if (p.epi.rs_out) {
    const int parts = gn::ceil_div(N, tc.block_n) * tc.splits * EPI_COLSPLIT;
    GN_CHECK_ARG(h, parts <= p.rs_capacity, "row-statistics buffer too small: %d partials, capacity %d", parts,
                 p.rs_capacity);
    p.epi.rs_parts = parts;
    h->last_rowstats_parts = parts;
  }
  // weight tensor map: [N rows][ktot] fp16, box {64, block_n}
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {BLOCK_K, (uint32_t)tc.block_n};
    int rc = make_tmap_f16(h, &p.tmB, W, 2, dims, strides, box);
    if (rc) return rc;
  }
  const int smem = smem_bytes_for(tc.block_n, tc.stages, tc.splits);
  if (!h->gemm_attr_set) {
    GN_CHECK_CUDA(h, cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->gemm_attr_set = true;
  }
  dim3 grid(gn::ceil_div(N, tc.block_n), tiles_m, tc.splits);
  // the K-splits of one output tile form a thread-block cluster (grid.z == cluster size)
  GN_CHECK_CUDA(h, launch_ex(h, gemm_tc_kernel, grid, dim3(GEMM_THREADS, 1, 1), smem, stream, tc.splits, p));
  h->last_cfg[0] = tc.block_n;
  h->last_cfg[1] = tc.splits;
  h->last_cfg[2] = tc.stages;
  h->last_cfg[3] = (int)(grid.x * grid.y * grid.z);
  return GN_OK;
}

// Tile configuration: heuristic model, or (gn_set_autotune) the fastest of the model's best candidates, measured once
// per problem shape with CUDA events on the caller's stream and cached in the handle.  Measurement never happens while
// the stream is being captured into a graph (the cached or modelled choice is used there), and re-running a launch is
// safe because GEMM outputs never alias their inputs.
static int launch_gemm(gn_handle* h, GemmParams& p, int tiles_m, const void* W, int64_t ktot, bool allow_split,
                       cudaStream_t stream) {
  const int N = p.epi.N;
  const bool geglu = p.epi.geglu != 0;
  const bool forced = h->force_block_n || h->force_splits || h->force_occupancy;
  char keybuf[96];
  const int rs_capacity = p.epi.rs_out ? p.epi.rs_parts : 0;
  p.rs_capacity = rs_capacity;
  snprintf(keybuf, sizeof(keybuf), "%d:%d:%d:%d:%d:%d:%d:%d:%d", p.mode, tiles_m, N, p.num_kblocks, geglu ? 1 : 0,
           p.epi.out32 ? 1 : 0, p.epi.residual ? 1 : 0, p.epi.ln_stats ? 1 : 0, rs_capacity);
  const std::string key(keybuf);
  if (!forced) {
    auto it = h->tune_cache.find(key);
    if (it != h->tune_cache.end()) {
      TileChoice tc{it->second[0], it->second[1], it->second[2], it->second[3]};
      int rc = launch_config(h, p, tc, tiles_m, W, ktot, stream);
      if (rc == GN_OK) h->launches++;
      return rc;
    }
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cap);
  if (h->autotune && !forced && !h->profiling && cap == cudaStreamCaptureStatusNone) {
    Candidate cand[10];
    const int nc = candidate_list(h, tiles_m, N, p.num_kblocks, geglu, allow_split, cand, 10, rs_capacity);
    if (nc > 1) {
      if (!h->tune_ev[0]) {
        GN_CHECK_CUDA(h, cudaEventCreate(&h->tune_ev[0]));
        GN_CHECK_CUDA(h, cudaEventCreate(&h->tune_ev[1]));
      }
      // In the real step the weights of one layer are long evicted from the 126 MB L2 when the layer runs again
      // (2.6 GB of weights stream through per denoise iteration), so weight-heavy shapes are timed L2-cold: the
      // workspace (> L2) is overwritten before every timed launch.
      const bool cold = (double)N * (double)ktot * 2.0 >= 4.0e6 && h->workspace && h->workspace_bytes >= (140 << 20);
      int best = 0;
      float best_ms = 1e30f;
      for (int i = 0; i < nc; ++i) {
        int rc = launch_config(h, p, cand[i].tc, tiles_m, W, ktot, stream);  // warm-up (tensor maps, smem carve-out)
        if (rc) return rc;
        float total = 0.f;
        const int reps = cold ? 2 : 1;
        for (int r = 0; r < reps; ++r) {
          if (cold) GN_CHECK_CUDA(h, cudaMemsetAsync(h->workspace, r, (size_t)h->workspace_bytes, stream));
          GN_CHECK_CUDA(h, cudaEventRecord(h->tune_ev[0], stream));
          for (int q = 0; q < (cold ? 1 : 3); ++q) {
            rc = launch_config(h, p, cand[i].tc, tiles_m, W, ktot, stream);
            if (rc) return rc;
          }
          GN_CHECK_CUDA(h, cudaEventRecord(h->tune_ev[1], stream));
          GN_CHECK_CUDA(h, cudaEventSynchronize(h->tune_ev[1]));
          float ms = 0.f;
          GN_CHECK_CUDA(h, cudaEventElapsedTime(&ms, h->tune_ev[0], h->tune_ev[1]));
          total += ms;
        }
        if (total < best_ms) {
          best_ms = total;
          best = i;
        }
      }
      const TileChoice& tc = cand[best].tc;
      h->tune_cache[key] = {tc.block_n, tc.splits, tc.stages, tc.tmem_cols};
      int rc = launch_config(h, p, tc, tiles_m, W, ktot, stream);
      if (rc == GN_OK) h->launches++;
      return rc;
    }
  }
  TileChoice tc = choose_tiles(h, tiles_m, N, p.num_kblocks, geglu, allow_split, rs_capacity);
  int rc = launch_config(h, p, tc, tiles_m, W, ktot, stream);
  if (rc == GN_OK) h->launches++;
  return rc;
}

}  // namespace gn

using namespace gn;

extern "C" int gn_linear(gn_handle* h, const void* A, int64_t lda, int M, int K, const void* W, int N, void* out,
                         int64_t ldo, const gn_epilogue* epi, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, A && W && out, "gn_linear: null pointer");
  GN_CHECK_ARG(h, M > 0 && N > 0 && K > 0, "gn_linear: bad shape M=%d N=%d K=%d", M, N, K);
  GN_CHECK_ARG(h, (K % 8) == 0 && (lda % 8) == 0, "gn_linear: K (%d) and lda (%lld) must be multiples of 8", K,
               (long long)lda);
  ProfScope prof(h, stream, GN_PROF_LINEAR, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + (double)M * N * ((epi && epi->out_fp32) ? 4 : 2));
  static thread_local GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_epilogue(h, p.epi, epi, out, ldo, M, N, M);
  if (rc) return rc;
  p.epi.ln_dim = K;
  p.mode = 0;
  p.num_kblocks = ceil_div(K, BLOCK_K);
  p.num_segs = 1;
  p.segs[0].nblk = p.num_kblocks;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BLOCK_K, BLOCK_M};
    rc = make_tmap_f16(h, &p.tmA[0], A, 2, dims, strides, box);
    if (rc) return rc;
  }
  return launch_gemm(h, p, ceil_div(M, BLOCK_M), W, K, /*allow_split=*/true, static_cast<cudaStream_t>(stream));
}

extern "C" int gn_conv2d(gn_handle* h, const void* x, int B, int H, int W, int C, const void* w, int Cout, int KH,
                         int KW, int stride, int pad, const void* ex0, int C_ex0, const void* ex1, int C_ex1,
                         void* out, int64_t ldo, const gn_epilogue* epi, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, x && w && out, "gn_conv2d: null pointer");
  GN_CHECK_ARG(h, B > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, "gn_conv2d: bad shape");
  GN_CHECK_ARG(h, (C % 8) == 0, "gn_conv2d: C (%d) must be a multiple of 8", C);
  GN_CHECK_ARG(h, stride == 1 || stride == 2, "gn_conv2d: stride %d unsupported", stride);
  GN_CHECK_ARG(h, KH * KW + 2 <= MAX_SEGS, "gn_conv2d: %dx%d kernel too large", KH, KW);
  GN_CHECK_ARG(h, stride == 1 || ((H % 2) == 0 && (W % 2) == 0), "gn_conv2d: stride 2 needs even H, W");
  const int Ho = (H + 2 * pad - KH) / stride + 1;
  const int Wo = (W + 2 * pad - KW) / stride + 1;
  GN_CHECK_ARG(h, Ho > 0 && Wo > 0, "gn_conv2d: empty output");
  const int M = B * Ho * Wo;
  const double kreal = (double)KH * KW * C + (ex0 ? C_ex0 : 0) + (ex1 ? C_ex1 : 0);
  ProfScope prof(h, stream, GN_PROF_CONV, 2.0 * M * Cout * kreal,
                 2.0 * ((double)B * H * W * C + (double)M * ((ex0 ? C_ex0 : 0) + (ex1 ? C_ex1 : 0)) + Cout * kreal +
                        (double)M * Cout));

  static thread_local GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_epilogue(h, p.epi, epi, out, ldo, M, Cout, Ho * Wo);
  if (rc) return rc;
  p.mode = 1;
  p.Ho = Ho;
  p.Wo = Wo;
  p.Bn = B;
  // 128 output pixels per tile as a (bw x bh x bb) box of powers of two; parts of the box beyond the image are
  // zero-filled by TMA and masked in the epilogue.
  auto pow2_ceil = [](int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
  };
  int bw = pow2_ceil(Wo);
  if (bw > 128) bw = 128;
  int bh = pow2_ceil(Ho);
  if (bh > 128 / bw) bh = 128 / bw;
  const int bb = 128 / (bw * bh);
  p.bw = bw;
  p.bh = bh;
  p.bb = bb;
  p.tiles_w = ceil_div(Wo, bw);
  p.tiles_h = ceil_div(Ho, bh);
  const int tiles_b = ceil_div(B, bb);
  const int tiles_m = p.tiles_w * p.tiles_h * tiles_b;

  const int Cp = round_up(C, BLOCK_K);
  const int cblk = Cp / BLOCK_K;
  int nseg = 0;
  int nmaps = 0;
  if (stride == 1) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {BLOCK_K, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    rc = make_tmap_f16(h, &p.tmA[0], x, 4, dims, strides, box);
    if (rc) return rc;
    nmaps = 1;
    for (int ky = 0; ky < KH; ++ky)
      for (int kx = 0; kx < KW; ++kx) {
        p.segs[nseg].map = 0;
        p.segs[nseg].dy = (int8_t)(ky - pad);
        p.segs[nseg].dx = (int8_t)(kx - pad);
        p.segs[nseg].nblk = cblk;
        ++nseg;
      }
  } else {
    // stride 2: input pixel (2i + ky - pad, 2j + kx - pad) lives in phase plane (py, px) at (i + oy, j + ox)
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const __half* base = static_cast<const __half*>(x) + ((int64_t)py * W + px) * C;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)B};
        uint64_t strides[3] = {(uint64_t)2 * C * 2, (uint64_t)2 * W * C * 2, (uint64_t)H * W * C * 2};
        uint32_t box[4] = {BLOCK_K, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
        rc = make_tmap_f16(h, &p.tmA[py * 2 + px], base, 4, dims, strides, box);
        if (rc) return rc;
      }
    nmaps = 4;
    for (int ky = 0; ky < KH; ++ky)
      for (int kx = 0; kx < KW; ++kx) {
        const int oy = ky - pad, ox = kx - pad;
        const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
        p.segs[nseg].map = (int8_t)(py * 2 + px);
        p.segs[nseg].dy = (int8_t)((oy - py) / 2);
        p.segs[nseg].dx = (int8_t)((ox - px) / 2);
        p.segs[nseg].nblk = cblk;
        ++nseg;
      }
  }
  int64_t ktot = (int64_t)KH * KW * Cp;
  const void* exs[2] = {ex0, ex1};
  const int exc[2] = {C_ex0, C_ex1};
  for (int i = 0; i < 2; ++i) {
    if (!exs[i]) continue;
    GN_CHECK_ARG(h, nmaps < 4, "gn_conv2d: extra 1x1 sources are not supported together with stride 2");
    GN_CHECK_ARG(h, exc[i] > 0 && (exc[i] % 8) == 0, "gn_conv2d: extra source channels must be a multiple of 8");
    uint64_t dims[4] = {(uint64_t)exc[i], (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)exc[i] * 2, (uint64_t)Wo * exc[i] * 2, (uint64_t)Ho * Wo * exc[i] * 2};
    uint32_t box[4] = {BLOCK_K, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
    rc = make_tmap_f16(h, &p.tmA[nmaps], exs[i], 4, dims, strides, box);
    if (rc) return rc;
    const int ecp = round_up(exc[i], BLOCK_K);
    p.segs[nseg].map = (int8_t)nmaps;
    p.segs[nseg].dy = 0;
    p.segs[nseg].dx = 0;
    p.segs[nseg].nblk = ecp / BLOCK_K;
    ++nseg;
    ++nmaps;
    ktot += ecp;
  }
  p.num_segs = nseg;
  p.num_kblocks = (int)(ktot / BLOCK_K);
  return launch_gemm(h, p, tiles_m, w, ktot, /*allow_split=*/true, static_cast<cudaStream_t>(stream));
}
