// GroupNorm (NHWC, optional channel-concat of two sources, optional fused SiLU), LayerNorm and row softmax.
// All HBM/L2-bound: 16-byte vector loads, fp32 statistics, one read for the statistics and one read+write for the
// normalisation.  Grids are sized to ~2 waves of the SM count.
#include "common.h"
#include "ptx.cuh"

namespace gn {

struct GNParams {
  const __half* x0;
  const __half* x1;
  int C0, C1, C, B, HW, G, cg;
  int pix_per_block;
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  float* stats;  // [B, G, 2] shifted sums
  __half* y;
};

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

// grid: (blocks_per_image, B); block: (bdx, bdy) — x walks 8-channel vectors, y walks pixels
__global__ void __launch_bounds__(256) gn_stats_kernel(GNParams p) {
  __shared__ float s_sum[64];
  __shared__ float s_sq[64];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < p.G; i += blockDim.x * blockDim.y) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int pix0 = blockIdx.x * p.pix_per_block;
  const int pix1 = min(p.HW, pix0 + p.pix_per_block);
  const int NV = p.C / 8;
  for (int cv = threadIdx.x; cv < NV; cv += blockDim.x) {
    const int c0 = cv * 8;
    const int gA = c0 / p.cg;
    const int gB = (c0 + 7) / p.cg;  // a vector spans at most two groups when cg >= 8 ... general case handled below
    float kj[8];
    int gj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gj[j] = (c0 + j) / p.cg;
      kj[j] = gn_shift(p, b, gj[j]);
    }
    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
    for (int pix = pix0 + threadIdx.y; pix < pix1; pix += blockDim.y) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(gn_src(p, b, pix, c0)));
      const __half2* hp = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(hp[t]);
        const float d0 = f.x - kj[2 * t], d1 = f.y - kj[2 * t + 1];
        a1[2 * t] += d0;
        a2[2 * t] += d0 * d0;
        a1[2 * t + 1] += d1;
        a2[2 * t + 1] += d1 * d1;
      }
    }
    if (gA == gB) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1 += a1[j];
        s2 += a2[j];
      }
      atomicAdd(&s_sum[gA], s1);
      atomicAdd(&s_sq[gA], s2);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_sum[gj[j]], a1[j]);
        atomicAdd(&s_sq[gj[j]], a2[j]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < p.G; i += blockDim.x * blockDim.y) {
    atomicAdd(&p.stats[((int64_t)b * p.G + i) * 2 + 0], s_sum[i]);
    atomicAdd(&p.stats[((int64_t)b * p.G + i) * 2 + 1], s_sq[i]);
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(GNParams p) {
  const int NV = p.C / 8;
  const int64_t total = (int64_t)p.B * p.HW * NV;
  const float inv_n = 1.0f / ((float)p.cg * (float)p.HW);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % NV);
    const int64_t bp = idx / NV;
    const int pix = (int)(bp % p.HW);
    const int b = (int)(bp / p.HW);
    const int c0 = cv * 8;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(gn_src(p, b, pix, c0)));
    const __half2* hp = reinterpret_cast<const __half2*>(&q);
    float xv[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(hp[t]);
      xv[2 * t] = f.x;
      xv[2 * t + 1] = f.y;
    }
    float o[8];
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / p.cg;
      if (g != g_prev) {
        const float k = gn_shift(p, b, g);
        const float s1 = __ldg(&p.stats[((int64_t)b * p.G + g) * 2 + 0]) * inv_n;
        const float s2 = __ldg(&p.stats[((int64_t)b * p.G + g) * 2 + 1]) * inv_n;
        mean = k + s1;
        rstd = rsqrtf(fmaxf(s2 - s1 * s1, 0.f) + p.eps);
        g_prev = g;
      }
      float v = (xv[j] - mean) * rstd * __ldg(p.gamma + c0 + j) + __ldg(p.beta + c0 + j);
      if (p.silu) v = silu_f(v);
      o[j] = v;
    }
    uint4 w;
    w.x = pack_half2(o[0], o[1]);
    w.y = pack_half2(o[2], o[3]);
    w.z = pack_half2(o[4], o[5]);
    w.w = pack_half2(o[6], o[7]);
    *reinterpret_cast<uint4*>(p.y + ((int64_t)b * p.HW + pix) * p.C + c0) = w;
  }
}

// One warp per row; the row lives in registers (C <= 2048).
constexpr int LN_MAX_VPL = 8;
__global__ void __launch_bounds__(256) layer_norm_kernel(const __half* __restrict__ x, int64_t ldx, int rows, int C,
                                                         float eps, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __half* __restrict__ y,
                                                         int64_t ldy) {
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
  GN_CHECK_ARG(h, h->stats_scratch && (int64_t)B * groups * 2 * 4 <= h->stats_scratch_bytes,
               "gn_group_norm: B*groups too large for the statistics scratch");
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
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.silu = silu;
  p.stats = static_cast<float*>(h->stats_scratch);
  p.y = static_cast<__half*>(y);
  GN_CHECK_CUDA(h, cudaMemsetAsync(p.stats, 0, (size_t)B * groups * 2 * sizeof(float), st));
  const int NV = C / 8;
  int bdx = NV < 256 ? NV : 256;
  if (bdx > 32) bdx = (bdx / 32) * 32;  // whole warps along x when possible
  int bdy = 256 / bdx;
  if (bdy < 1) bdy = 1;
  int target_blocks = (2 * h->num_sms + B - 1) / B;
  int pix = (HW + target_blocks - 1) / target_blocks;
  if (pix < bdy) pix = bdy;
  p.pix_per_block = pix;
  dim3 grid((HW + pix - 1) / pix, B);
  dim3 block(bdx, bdy);
  gn_stats_kernel<<<grid, block, 0, st>>>(p);
  GN_CHECK_LAUNCH(h);
  const int64_t total = (int64_t)B * HW * NV;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  gn_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
  GN_CHECK_LAUNCH(h);
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
  layer_norm_kernel<<<(rows + rows_per_block - 1) / rows_per_block, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), ldx, rows, C, eps, gamma, beta, static_cast<__half*>(y), ldy);
  GN_CHECK_LAUNCH(h);
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
