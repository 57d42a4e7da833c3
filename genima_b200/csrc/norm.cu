// GroupNorm (NHWC, optional channel-concat of two sources, optional fused SiLU), LayerNorm and row softmax.
// All HBM/L2-bound: 16-byte vector loads, fp32 statistics; GroupNorm is a single launch with a grid barrier between
// the statistics and the normalisation phase (one read + one write when the per-CTA slice fits in shared memory).
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace gn {

struct GNParams {
  const __half* x0;
  const __half* x1;
  int C0, C1, C, B, HW, G, cg, NV;
  int ctas_per_b, pix_per_cta;
  int k;      // pixel lanes per CTA: thread t owns channel vector t % NV and pixels pix0 + t / NV + i * k
  int cache;  // 1: the CTA's slice of x is kept in shared memory between the two phases
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  float* partial;     // [B][ctas_per_b][G][2] shifted (sum, sumsq) of each CTA's slice
  unsigned int* bar;  // grid barrier: [0] arrival count, [1] generation
  unsigned long long* trace;  // optional phase timestamps of CTA 0 (gn_set_gemm_trace)
  __half* y;
};

__device__ __forceinline__ void gn_stamp(const GNParams& p, int slot) {
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

__device__ __forceinline__ const __half* gn_src(const GNParams& p, int b, int pix, int c) {
  // pointer to channel c (multiple of 8; C0 % 8 == 0) of pixel pix in the virtual concat tensor
  if (c < p.C0) return p.x0 + ((int64_t)b * p.HW + pix) * p.C0 + c;
  return p.x1 + ((int64_t)b * p.HW + pix) * p.C1 + (c - p.C0);
}
// shift value of group g: first element of the group at pixel 0 (keeps E[(x-K)^2] - E[x-K]^2 well conditioned)
__device__ __forceinline__ float gn_shift(const GNParams& p, int b, int g) {
  const int c = g * p.cg;
  const __half* s = (c < p.C0) ? p.x0 + (int64_t)b * p.HW * p.C0 + c : p.x1 + (int64_t)b * p.HW * p.C1 + (c - p.C0);
  return __half2float(__ldg(s));
}

// Group index of each of the 8 consecutive channels c0 .. c0 + 7 (one integer division when cg >= 8).
__device__ __forceinline__ void gn_groups8(int c0, int cg, int (&g)[8]) {
  const int g0 = c0 / cg;
  if (cg >= 8) {
    const int rem = c0 - g0 * cg;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = g0 + ((rem + j >= cg) ? 1 : 0);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = (c0 + j) / cg;
  }
}

__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// GroupNorm in ONE launch, one HBM read + one write, bit-reproducible:
//   phase 1  every CTA reads its pixel slice once (kept in shared memory when it fits), accumulates shifted sums per
//            thread, reduces them per channel then per group in a FIXED order and publishes [G][2] partials;
//   barrier  grid-wide (all CTAs are co-resident: grid <= #SMs, one CTA per SM), sense-reversing, self-resetting so a
//            captured CUDA graph can replay it;
//   phase 2  every CTA sums the partials of its image in CTA order, normalises its slice and writes y (+ SiLU).
// Dynamic smem: [red: T*16 floats][chan: C*2 floats][gstat: G*2 floats][slice cache: uint4 x pix_per_cta*NV].
__global__ void __launch_bounds__(1024, 1) gn_fused_kernel(const GNParams p) {
  extern __shared__ __align__(16) uint8_t gn_smem[];
  const int T = p.NV * p.k;  // active threads
  float* red = reinterpret_cast<float*>(gn_smem);
  float* chan = red + (size_t)T * 16;
  float* gstat = chan + (size_t)p.C * 2;
  float* kshift = gstat + ((p.G * 2 + 3) & ~3);  // (spare)
  uint4* slice = reinterpret_cast<uint4*>(kshift + ((p.G + 3) & ~3));
  __shared__ unsigned int s_gen;

  const int t = threadIdx.x;
  const int b = blockIdx.x / p.ctas_per_b;
  const int ci = blockIdx.x % p.ctas_per_b;
  const int pix0 = ci * p.pix_per_cta;
  const int pix1 = min(p.HW, pix0 + p.pix_per_cta);
  const bool active = t < T;
  const int cv = active ? t % p.NV : 0;
  const int pl = active ? t / p.NV : 0;
  const int c0 = cv * 8;
  pdl_trigger();
  pdl_wait();  // x comes from the previous kernel; the barrier generation word may still be moving until it is done
  gn_stamp(p, 0);
  if (t == 0) s_gen = *reinterpret_cast<volatile unsigned int*>(p.bar + 1);

  // ---- phase 1: shifted sums of this thread's channel vector over its pixels
  int gj[8];
  gn_groups8(c0, p.cg, gj);
  float kj[8];
  {
    const float k_first = gn_shift(p, b, gj[0]);
    const float k_last = (gj[7] != gj[0] && p.cg >= 8) ? gn_shift(p, b, gj[7]) : k_first;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      kj[j] = (p.cg >= 8) ? (gj[j] == gj[0] ? k_first : k_last) : ((j == 0) ? k_first : gn_shift(p, b, gj[j]));
  }
  if (active) {
    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
    for (int pix = pix0 + pl; pix < pix1; pix += p.k) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(gn_src(p, b, pix, c0)));
      if (p.cache) slice[(size_t)(pix - pix0) * p.NV + cv] = q;
      const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __half22float2(hp[u]);
        const float d0 = f.x - kj[2 * u], d1 = f.y - kj[2 * u + 1];
        a1[2 * u] += d0;
        a2[2 * u] = fmaf(d0, d0, a2[2 * u]);
        a1[2 * u + 1] += d1;
        a2[2 * u + 1] = fmaf(d1, d1, a2[2 * u + 1]);
      }
    }
    float4* r4 = reinterpret_cast<float4*>(red + (size_t)t * 16);
    r4[0] = make_float4(a1[0], a1[1], a1[2], a1[3]);
    r4[1] = make_float4(a1[4], a1[5], a1[6], a1[7]);
    r4[2] = make_float4(a2[0], a2[1], a2[2], a2[3]);
    r4[3] = make_float4(a2[4], a2[5], a2[6], a2[7]);
  }
  __syncthreads();
  gn_stamp(p, 1);
  // per channel: fixed-order sum over the k pixel lanes
  for (int c = t; c < p.C; c += blockDim.x) {
    const int v = c >> 3, j = c & 7;
    float s1 = 0.f, s2 = 0.f;
    for (int l = 0; l < p.k; ++l) {
      const float* r = red + ((size_t)l * p.NV + v) * 16;
      s1 += r[j];
      s2 += r[8 + j];
    }
    chan[2 * c] = s1;
    chan[2 * c + 1] = s2;
  }
  __syncthreads();
  // per group: one warp per group, lane-strided sums then a fixed shuffle tree; publish this CTA's partial
  {
    const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < p.G; g += nwarps) {
      float s1 = 0.f, s2 = 0.f;
      for (int c = g * p.cg + lane; c < (g + 1) * p.cg; c += 32) {
        s1 += chan[2 * c];
        s2 += chan[2 * c + 1];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) {
        float* dst = p.partial + (((size_t)b * p.ctas_per_b + ci) * p.G + g) * 2;
        dst[0] = s1;
        dst[1] = s2;
      }
    }
  }
  // ---- grid barrier; the LAST CTA to arrive reduces the partials of every image to final (mean, rstd) per group in
  // a fixed order and publishes them before it releases the others (one reader instead of #CTA readers hammering the
  // same few L2 lines).
  __shared__ unsigned int s_last;
  __syncthreads();
  gn_stamp(p, 2);
  if (t == 0) {
    // release: this CTA's partials (ordered before by the __syncthreads above) become visible with the arrival
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.bar) : "memory");
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  float* final_stats = p.partial + (size_t)p.B * p.ctas_per_b * p.G * 2;  // [B][G][2] = (mean, rstd)
  if (s_last) {
    const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    for (int bg = warp; bg < p.B * p.G; bg += nwarps) {
      const int bb = bg / p.G, g = bg - bb * p.G;
      const float* src = p.partial + ((size_t)bb * p.ctas_per_b * p.G + g) * 2;
      float2 v[5];  // up to 160 partials: five independent loads per lane in flight, summed in index order
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int i = lane + 32 * u;
        v[u] = (i < p.ctas_per_b) ? __ldcg(reinterpret_cast<const float2*>(src + (size_t)i * p.G * 2))
                                  : make_float2(0.f, 0.f);
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        s1 += v[u].x;
        s2 += v[u].y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) {
        const float inv_n = 1.0f / ((float)p.cg * (float)p.HW);
        const float m1 = s1 * inv_n, m2 = s2 * inv_n;
        final_stats[2 * bg] = gn_shift(p, bb, g) + m1;
        final_stats[2 * bg + 1] = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + p.eps);
      }
    }
    __syncthreads();
    if (t == 0) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.bar), "r"(0u) : "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar + 1) : "memory");
    }
  } else if (t == 0) {
    const long long t0 = clock64();
    unsigned int g;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(p.bar + 1) : "memory");
      if (g == s_gen && clock64() - t0 > 4000000000LL) __trap();  // a scheduling bug must never hang the GPU
    } while (g == s_gen);
  }
  __syncthreads();
  gn_stamp(p, 3);
  // ---- phase 2: normalisation with the published statistics
  if (t < 2 * p.G) gstat[t] = __ldcg(final_stats + (size_t)b * p.G * 2 + t);
  __syncthreads();
  gn_stamp(p, 4);
  if (!active) return;
  float sc[8], sh[8];
  {
    const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + c0));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + 4));
    const float4 ba = __ldg(reinterpret_cast<const float4*>(p.beta + c0));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + 4));
    const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mean = gstat[2 * gj[j]], rstd = gstat[2 * gj[j] + 1];
      sc[j] = rstd * gam[j];
      sh[j] = bet[j] - mean * sc[j];
    }
  }
  for (int pix = pix0 + pl; pix < pix1; pix += p.k) {
    const uint4 q = p.cache ? slice[(size_t)(pix - pix0) * p.NV + cv]
                            : __ldg(reinterpret_cast<const uint4*>(gn_src(p, b, pix, c0)));
    const __half2* hp = reinterpret_cast<const __half2*>(&q);
    float o[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __half22float2(hp[u]);
      o[2 * u] = fmaf(f.x, sc[2 * u], sh[2 * u]);
      o[2 * u + 1] = fmaf(f.y, sc[2 * u + 1], sh[2 * u + 1]);
    }
    if (p.silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = silu_fast(o[j]);
    }
    uint4 w;
    w.x = pack_half2(o[0], o[1]);
    w.y = pack_half2(o[2], o[3]);
    w.z = pack_half2(o[4], o[5]);
    w.w = pack_half2(o[6], o[7]);
    *reinterpret_cast<uint4*>(p.y + ((int64_t)b * p.HW + pix) * p.C + c0) = w;
  }
  gn_stamp(p, 5);
}

// ---------------------------------------------------------------------------------------------------------------------
// GroupNorm from statistics accumulated by the producing GEMM epilogues (gemm.cu: epilogue_tail): ONE elementwise pass.
// stats: uint64 [B][C / bucket][2], 2^20-scaled fixed-point (sum, sumsq) of every (image, bucket of channels).
struct GNApplyParams {
  const __half* x0;
  const __half* x1;
  const unsigned long long* st0;
  const unsigned long long* st1;
  int C0, C1, C, B, HW, G, cg, NV, bucket;
  int ctas_per_b;
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  __half* y;
};

constexpr int GNA_THREADS = 256;

__global__ void __launch_bounds__(GNA_THREADS, 4) gn_apply_kernel(const GNApplyParams p) {
  extern __shared__ __align__(16) uint8_t gna_smem[];
  float* sc = reinterpret_cast<float*>(gna_smem);  // [C] gamma -> scale, then [C] beta -> shift
  float* sh = sc + p.C;
  __shared__ float s_mean[64], s_rstd[64];
  const int t = threadIdx.x;
  const int b = blockIdx.x / p.ctas_per_b;
  const int ci = blockIdx.x % p.ctas_per_b;
  // gamma / beta are weights: fetch them while the producer of x is still running (programmatic dependent launch)
  for (int c = t; c < p.C; c += GNA_THREADS) {
    sc[c] = __ldg(p.gamma + c);
    sh[c] = __ldg(p.beta + c);
  }
  pdl_trigger();
  pdl_wait();  // x and the statistics come from the previous kernels in the stream
  // this CTA's share of the image's 16-byte vectors, processed in batches of PF vectors per thread (all loads of a
  // batch in flight together); the first batch is requested before the statistics are reduced
  const int64_t nvec = (int64_t)p.HW * p.NV;
  const int64_t per = (nvec + p.ctas_per_b - 1) / p.ctas_per_b;
  const int64_t v_begin = (int64_t)ci * per;
  const int64_t v_end = min(nvec, v_begin + per);
  constexpr int PF = 4;
  uint4 pre[PF];
  auto load_batch = [&](int64_t base) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int64_t v = base + t + (int64_t)i * GNA_THREADS;
      if (v < v_end) {
        const int pix = (int)(v / p.NV);
        const int c0 = (int)(v - (int64_t)pix * p.NV) * 8;
        const __half* src = (c0 < p.C0) ? p.x0 + ((int64_t)b * p.HW + pix) * p.C0 + c0
                                        : p.x1 + ((int64_t)b * p.HW + pix) * p.C1 + (c0 - p.C0);
        pre[i] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
  };
  load_batch(v_begin);
  if (t < p.G) {
    // exact integer totals of the group's buckets (each bucket lies entirely in x0 or in x1)
    long long s1 = 0, s2 = 0;
    const int nb0 = p.C0 / p.bucket, nb1 = p.C1 / p.bucket;
    const int per_g = p.cg / p.bucket;
    const int kb = t * per_g;
    for (int k0 = 0; k0 < per_g; k0 += 4) {  // four buckets (eight 8-byte loads) in flight
      ulonglong2 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kb + k0 + u;
        if (k0 + u < per_g) {
          const unsigned long long* src =
              (k < nb0) ? p.st0 + ((size_t)b * nb0 + k) * 2 : p.st1 + ((size_t)b * nb1 + (k - nb0)) * 2;
          q[u] = __ldcg(reinterpret_cast<const ulonglong2*>(src));
        } else {
          q[u] = make_ulonglong2(0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1 += static_cast<long long>(q[u].x);
        s2 += static_cast<long long>(q[u].y);
      }
    }
    const double inv = 1.0 / (1048576.0 * (double)p.cg * (double)p.HW);
    const double mean = (double)s1 * inv;
    double var = (double)s2 * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[t] = (float)mean;
    s_rstd[t] = rsqrtf((float)var + p.eps);
  }
  __syncthreads();
  for (int c = t; c < p.C; c += GNA_THREADS) {
    const int g = c / p.cg;
    const float a = s_rstd[g] * sc[c];
    sh[c] = sh[c] - s_mean[g] * a;
    sc[c] = a;
  }
  __syncthreads();
  for (int64_t base = v_begin; base < v_end; base += (int64_t)PF * GNA_THREADS) {
    uint4 cur[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) cur[i] = pre[i];
    // next batch in flight while this one is normalised
    if (base + (int64_t)PF * GNA_THREADS < v_end) load_batch(base + (int64_t)PF * GNA_THREADS);
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int64_t v = base + t + (int64_t)i * GNA_THREADS;
      if (v >= v_end) continue;
      const int pix = (int)(v / p.NV);
      const int c0 = (int)(v - (int64_t)pix * p.NV) * 8;
      const uint4 q = cur[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&q);
      const float4 a0 = *reinterpret_cast<const float4*>(sc + c0);
      const float4 a1 = *reinterpret_cast<const float4*>(sc + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sh + c0);
      const float4 b1 = *reinterpret_cast<const float4*>(sh + c0 + 4);
      float o[8];
      float2 f = __half22float2(hp[0]);
      o[0] = fmaf(f.x, a0.x, b0.x);
      o[1] = fmaf(f.y, a0.y, b0.y);
      f = __half22float2(hp[1]);
      o[2] = fmaf(f.x, a0.z, b0.z);
      o[3] = fmaf(f.y, a0.w, b0.w);
      f = __half22float2(hp[2]);
      o[4] = fmaf(f.x, a1.x, b1.x);
      o[5] = fmaf(f.y, a1.y, b1.y);
      f = __half22float2(hp[3]);
      o[6] = fmaf(f.x, a1.z, b1.z);
      o[7] = fmaf(f.y, a1.w, b1.w);
      if (p.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = silu_fast(o[j]);
      }
      uint4 w;
      w.x = pack_half2(o[0], o[1]);
      w.y = pack_half2(o[2], o[3]);
      w.z = pack_half2(o[4], o[5]);
      w.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(p.y + ((int64_t)b * p.HW + pix) * p.C + c0) = w;
    }
  }
}

// One warp per row; the row lives in registers (C <= 2048).
constexpr int LN_MAX_VPL = 8;
__global__ void __launch_bounds__(256) layer_norm_kernel(const __half* __restrict__ x, int64_t ldx, int rows, int C,
                                                         float eps, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __half* __restrict__ y,
                                                         int64_t ldy) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int NV = C / 8;
  float v[LN_MAX_VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VPL; ++i) {
    const int cv = lane + i * 32;
    if (cv < NV) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (int64_t)row * ldx + cv * 8));
      const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(hp[t]);
        v[i][2 * t] = f.x;
        v[i][2 * t + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VPL; ++i) {
    const int cv = lane + i * 32;
    if (cv < NV) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_VPL; ++i) {
    const int cv = lane + i * 32;
    if (cv < NV) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        o[j] = (v[i][j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      }
      uint4 w;
      w.x = pack_half2(o[0], o[1]);
      w.y = pack_half2(o[2], o[3]);
      w.z = pack_half2(o[4], o[5]);
      w.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(y + (int64_t)row * ldy + cv * 8) = w;
    }
  }
}

// Row softmax of scale * x: one block per row, row in registers (cols <= 8192, cols % 8 == 0).  Input fp32 (scores
// from a GEMM with fp32 output) or fp16; output fp16 (may alias an fp16 input).
constexpr int SM_VPT = 4;
template <typename TIn>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const TIn* __restrict__ x, int64_t ldx,
                                                           __half* __restrict__ y, int64_t ldy, int cols,
                                                           float scale) {
  __shared__ float red[8];
  const TIn* rp = x + (int64_t)blockIdx.x * ldx;
  __half* wp = y + (int64_t)blockIdx.x * ldy;
  const int NV = cols / 8;
  float v[SM_VPT][8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SM_VPT; ++i) {
    const int cv = threadIdx.x + i * 256;
    if (cv < NV) {
      if constexpr (sizeof(TIn) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(rp + cv * 8);
        const float4 b = *reinterpret_cast<const float4*>(rp + cv * 8 + 4);
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
      } else {
        const uint4 q = *reinterpret_cast<const uint4*>(rp + cv * 8);
        const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __half22float2(hp[t]);
          v[i][2 * t] = f.x;
          v[i][2 * t + 1] = f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] *= scale;
        mx = fmaxf(mx, v[i][j]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SM_VPT; ++i) {
    const int cv = threadIdx.x + i * 256;
    if (cv < NV) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = __expf(v[i][j] - mx);
        sum += v[i][j];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < SM_VPT; ++i) {
    const int cv = threadIdx.x + i * 256;
    if (cv < NV) {
      uint4 w;
      w.x = pack_half2(v[i][0] * inv, v[i][1] * inv);
      w.y = pack_half2(v[i][2] * inv, v[i][3] * inv);
      w.z = pack_half2(v[i][4] * inv, v[i][5] * inv);
      w.w = pack_half2(v[i][6] * inv, v[i][7] * inv);
      *reinterpret_cast<uint4*>(wp + cv * 8) = w;
    }
  }
}

}  // namespace gn

using namespace gn;

extern "C" int gn_group_norm(gn_handle* h, const void* x0, int C0, const void* x1, int C1, int B, int HW, int groups,
                             float eps, const float* gamma, const float* beta, int silu, const float* stats_in,
                             void* y, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, x0 && y && gamma && beta, "gn_group_norm: null pointer");
  if (!x1) C1 = 0;
  const int C = C0 + C1;
  GN_CHECK_ARG(h, B > 0 && HW > 0 && groups > 0 && groups <= 64, "gn_group_norm: bad shape");
  GN_CHECK_ARG(h, (C0 % 8) == 0 && (C1 % 8) == 0 && (C % groups) == 0, "gn_group_norm: C0=%d C1=%d groups=%d", C0, C1,
               groups);
  GN_CHECK_ARG(h, stats_in == nullptr, "gn_group_norm: stats_in is not supported by this build");
  GN_CHECK_ARG(h, C / 8 <= 1024, "gn_group_norm: C=%d too large", C);
  GN_CHECK_ARG(h, B <= h->num_sms, "gn_group_norm: B=%d exceeds the SM count (grid barrier needs co-resident CTAs)", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(h, stream, GN_PROF_NORM, 0.0, 4.0 * B * HW * C);
  GNParams p;
  p.x0 = static_cast<const __half*>(x0);
  p.x1 = static_cast<const __half*>(x1);
  p.C0 = C0;
  p.C1 = C1;
  p.C = C;
  p.B = B;
  p.HW = HW;
  p.G = groups;
  p.cg = C / groups;
  p.NV = C / 8;
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.silu = silu;
  p.y = static_cast<__half*>(y);
  p.bar = static_cast<unsigned int*>(h->stats_scratch);
  p.trace = static_cast<unsigned long long*>(h->gemm_trace);
  p.partial = reinterpret_cast<float*>(static_cast<uint8_t*>(h->stats_scratch) + 256);
  // CTAs per image: enough bytes per CTA to amortise the barrier, never more than one CTA per SM in total
  const int64_t vectors = (int64_t)HW * (C / 8);
  int ctas = (int)((vectors + 1535) / 1536);  // ~1.5 sixteen-byte vectors per thread before the grid is capped
  const int sm_cap = (h->gn_max_ctas > 0 && h->gn_max_ctas < h->num_sms) ? h->gn_max_ctas : h->num_sms;
  GN_CHECK_ARG(h, B <= sm_cap, "gn_group_norm: B=%d exceeds the CTA cap %d", B, sm_cap);
  const int max_ctas = sm_cap / B;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas > HW) ctas = HW;
  if (ctas < 1) ctas = 1;
  p.pix_per_cta = (HW + ctas - 1) / ctas;
  p.ctas_per_b = (HW + p.pix_per_cta - 1) / p.pix_per_cta;
  p.k = 1024 / p.NV;
  if (p.k > p.pix_per_cta) p.k = p.pix_per_cta;
  if (p.k < 1) p.k = 1;
  const int T = p.NV * p.k;
  const int threads = (T + 31) / 32 * 32;
  GN_CHECK_ARG(h, (int64_t)B * (p.ctas_per_b + 1) * groups * 8 + 256 <= h->stats_scratch_bytes,
               "gn_group_norm: statistics scratch too small");
  const size_t fixed = (size_t)T * 16 * 4 + (size_t)C * 2 * 4 + (size_t)((groups * 2 + 3) & ~3) * 4 +
                       (size_t)((groups + 3) & ~3) * 4;
  GN_CHECK_ARG(h, p.ctas_per_b <= 160, "gn_group_norm: more than 160 CTAs per image");
  const size_t slice = (size_t)p.pix_per_cta * p.NV * 16;
  p.cache = (fixed + slice <= 200 * 1024) ? 1 : 0;
  const size_t smem = fixed + (p.cache ? slice : 0);
  if (!h->gn_attr_set) {
    GN_CHECK_CUDA(h, cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    h->gn_attr_set = true;
  }
  GN_CHECK_CUDA(h, launch_ex(h, gn_fused_kernel, dim3(B * p.ctas_per_b, 1, 1), dim3(threads, 1, 1), smem, st, 1, p));
  h->launches++;
  return GN_OK;
}

extern "C" int gn_group_norm_apply(gn_handle* h, const void* x0, int C0, const void* stats0, const void* x1, int C1,
                                   const void* stats1, int bucket, int B, int HW, int groups, float eps,
                                   const float* gamma, const float* beta, int silu, void* y, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, x0 && stats0 && y && gamma && beta, "gn_group_norm_apply: null pointer");
  {
    // timing ablation only (GENIMA_B200_SKIP bit 1): the launch is dropped, the output stays uninitialised
    static const char* skip_env = getenv("GENIMA_B200_SKIP");
    if (skip_env && (atoi(skip_env) & 1)) return GN_OK;
  }
  if (!x1) C1 = 0;
  GN_CHECK_ARG(h, x1 == nullptr || stats1 != nullptr, "gn_group_norm_apply: x1 given without its statistics");
  const int C = C0 + C1;
  GN_CHECK_ARG(h, B > 0 && HW > 0 && groups > 0 && groups <= 64, "gn_group_norm_apply: bad shape");
  GN_CHECK_ARG(h, (C0 % 8) == 0 && (C1 % 8) == 0 && (C % groups) == 0, "gn_group_norm_apply: C0=%d C1=%d groups=%d", C0,
               C1, groups);
  const int cg = C / groups;
  GN_CHECK_ARG(h, bucket > 0 && (cg % bucket) == 0 && (C0 % bucket) == 0 && (C1 % bucket) == 0,
               "gn_group_norm_apply: bucket %d must divide C0=%d, C1=%d and the %d channels per group", bucket, C0, C1, cg);
  GN_CHECK_ARG(h, (size_t)C * 8 <= 96 * 1024, "gn_group_norm_apply: C=%d too large", C);
  ProfScope prof(h, stream, GN_PROF_NORM, 0.0, 4.0 * B * HW * C);
  GNApplyParams p;
  p.x0 = static_cast<const __half*>(x0);
  p.x1 = static_cast<const __half*>(x1);
  p.st0 = static_cast<const unsigned long long*>(stats0);
  p.st1 = static_cast<const unsigned long long*>(stats1);
  p.C0 = C0;
  p.C1 = C1;
  p.C = C;
  p.B = B;
  p.HW = HW;
  p.G = groups;
  p.cg = cg;
  p.NV = C / 8;
  p.bucket = bucket;
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.silu = silu;
  p.y = static_cast<__half*>(y);
  // one batch of 4 sixteen-byte vectors per thread on small tensors, at most 4 CTAs per SM in total
  const int64_t nvec = (int64_t)HW * p.NV;
  int64_t ctas = (nvec + GNA_THREADS * 4 - 1) / (GNA_THREADS * 4);
  const int64_t cap = (int64_t)h->num_sms * 4 / B > 0 ? (int64_t)h->num_sms * 4 / B : 1;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  p.ctas_per_b = (int)ctas;
  const size_t smem = (size_t)C * 8;
  if (smem > 48 * 1024 && !h->gna_attr_set) {
    GN_CHECK_CUDA(h, cudaFuncSetAttribute(gn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    h->gna_attr_set = true;
  }
  GN_CHECK_CUDA(h, launch_ex(h, gn_apply_kernel, dim3(B * p.ctas_per_b, 1, 1), dim3(GNA_THREADS, 1, 1), smem,
                              static_cast<cudaStream_t>(stream), 1, p));
  h->launches++;
  return GN_OK;
}

extern "C" int gn_layer_norm(gn_handle* h, const void* x, int64_t ldx, int rows, int C, float eps, const float* gamma,
                             const float* beta, void* y, int64_t ldy, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, x && y && gamma && beta, "gn_layer_norm: null pointer");
  GN_CHECK_ARG(h, rows > 0 && C > 0 && (C % 8) == 0 && C <= 8 * 32 * LN_MAX_VPL, "gn_layer_norm: C=%d unsupported", C);
  GN_CHECK_ARG(h, (ldx % 8) == 0 && (ldy % 8) == 0, "gn_layer_norm: row strides must be multiples of 8");
  ProfScope prof(h, stream, GN_PROF_NORM, 0.0, 4.0 * rows * C);
  const int rows_per_block = 8;
  GN_CHECK_CUDA(h, launch_ex(h, layer_norm_kernel, dim3((rows + rows_per_block - 1) / rows_per_block, 1, 1),
                              dim3(256, 1, 1), 0, static_cast<cudaStream_t>(stream), 1, static_cast<const __half*>(x), ldx,
                              rows, C, eps, gamma, beta, static_cast<__half*>(y), ldy));
  h->launches++;
  return GN_OK;
}

extern "C" int gn_softmax_rows(gn_handle* h, const void* x, int x_fp32, int64_t ldx, void* y, int64_t ldy, int rows,
                               int cols, float scale, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, x && y && rows > 0 && cols > 0, "gn_softmax_rows: bad arguments");
  GN_CHECK_ARG(h, (cols % 8) == 0 && cols <= 8 * 256 * SM_VPT && (ldx % 8) == 0 && (ldy % 8) == 0,
               "gn_softmax_rows: cols=%d unsupported", cols);
  GN_CHECK_ARG(h, !(x_fp32 && x == y), "gn_softmax_rows: fp32 input cannot alias the fp16 output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(h, stream, GN_PROF_NORM, 0.0, (double)rows * cols * (x_fp32 ? 6.0 : 4.0));
  if (x_fp32)
    softmax_rows_kernel<float><<<rows, 256, 0, st>>>(static_cast<const float*>(x), ldx, static_cast<__half*>(y), ldy,
                                                     cols, scale);
  else
    softmax_rows_kernel<__half><<<rows, 256, 0, st>>>(static_cast<const __half*>(x), ldx, static_cast<__half*>(y),
                                                      ldy, cols, scale);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}
