// Data-movement and element-wise kernels around the contraction kernels: layout conversion, image pre/post-processing,
// view tiling, nearest upsampling, max pooling, timestep embedding and the Euler scheduler update.  All HBM-bound,
// vectorised to 16 bytes where the layout allows, grid-stride with grids capped at a few waves of the SM count.
#include "common.h"
#include "ptx.cuh"

namespace gn {

static inline unsigned grid_for(const gn_handle* h, int64_t work_items, int threads = 256) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)h->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

#define GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int CV) {
  const int64_t total = (int64_t)B * 2 * H * 2 * W * CV;
  GRID_STRIDE(i, total) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    y[i] = __ldg(x + (((int64_t)b * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv);
  }
}

__global__ void maxpool3x3s2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int CV = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * CV;
  GRID_STRIDE(i, total) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    __half2 m[4];
    bool any = false;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + iy) * W + ix) * C + cv * 8));
        const __half2* hp = reinterpret_cast<const __half2*>(&q);
        if (!any) {
          for (int t = 0; t < 4; ++t) m[t] = hp[t];
          any = true;
        } else {
          for (int t = 0; t < 4; ++t) m[t] = __hmax2(m[t], hp[t]);
        }
      }
    }
    uint4 w;
    w.x = *reinterpret_cast<uint32_t*>(&m[0]);
    w.y = *reinterpret_cast<uint32_t*>(&m[1]);
    w.z = *reinterpret_cast<uint32_t*>(&m[2]);
    w.w = *reinterpret_cast<uint32_t*>(&m[3]);
    *reinterpret_cast<uint4*>(y + (((int64_t)b * Ho + oy) * Wo + ox) * C + cv * 8) = w;
  }
}

__global__ void add_kernel(const __half2* __restrict__ a, const __half2* __restrict__ b, __half2* __restrict__ o,
                           int64_t n2) {
  GRID_STRIDE(i, n2) {
    const float2 fa = __half22float2(a[i]);
    const float2 fb = __half22float2(b[i]);
    o[i] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
  }
}

__global__ void scale_kernel(const __half* __restrict__ x, float s, __half* __restrict__ y, int64_t n) {
  GRID_STRIDE(i, n) y[i] = __float2half_rn(__half2float(x[i]) * s);
}

// diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], computed in fp32
__global__ void timestep_embedding_kernel(float t, int dim, __half* __restrict__ out) {
  const int half_dim = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half_dim) return;
  const float freq = expf(-logf(10000.0f) * (float)i / (float)half_dim);
  const float a = t * freq;
  out[i] = __float2half_rn(cosf(a));
  out[half_dim + i] = __float2half_rn(sinf(a));
}

__global__ void euler_step_kernel(const __half* __restrict__ x, const __half* __restrict__ eps, float dsigma,
                                  float inv_scale_next, __half* __restrict__ x_next, __half* __restrict__ x_scaled,
                                  int64_t n) {
  GRID_STRIDE(i, n) {
    // diffusers upcasts sample to fp32 for the update and casts the result back to the model dtype
    const float xn = __half2float(x[i]) + dsigma * __half2float(eps[i]);
    const __half h = __float2half_rn(xn);
    x_next[i] = h;
    if (x_scaled) x_scaled[i] = __float2half_rn(__half2float(h) * inv_scale_next);
  }
}

__global__ void euler_ancestral_step_kernel(const __half* __restrict__ x, const __half* __restrict__ eps,
                                            const __half* __restrict__ noise, float dsigma, float sigma_up,
                                            float inv_scale_next, __half* __restrict__ x_next,
                                            __half* __restrict__ x_scaled, int64_t n) {
  GRID_STRIDE(i, n) {
    // EulerAncestralDiscreteScheduler.step: fp32 update, noise drawn in the model dtype, result cast back to fp16
    float xn = __half2float(x[i]) + dsigma * __half2float(eps[i]);
    xn += __half2float(noise[i]) * sigma_up;
    const __half h = __float2half_rn(xn);
    x_next[i] = h;
    if (x_scaled) x_scaled[i] = __float2half_rn(__half2float(h) * inv_scale_next);
  }
}

struct Norm3 {
  float mean[3];
  float inv_std[3];
  int enabled;
};

template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, int B, int C, int H, int W, int Cpad, Norm3 nm,
                                    __half* __restrict__ dst) {
  const int64_t total = (int64_t)B * H * W * Cpad;
  GRID_STRIDE(i, total) {
    const int c = (int)(i % Cpad);
    int64_t r = i / Cpad;
    const int xw = (int)(r % W);
    r /= W;
    const int yh = (int)(r % H);
    const int b = (int)(r / H);
    float v = 0.f;
    if (c < C) {
      v = (float)src[(((int64_t)b * C + c) * H + yh) * W + xw];
      if (nm.enabled && c < 3) v = (v / 255.0f - nm.mean[c]) * nm.inv_std[c];
    }
    dst[i] = __float2half_rn(v);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ src, int B, int C, int H, int W, int Cpad,
                                    T* __restrict__ dst) {
  const int64_t total = (int64_t)B * C * H * W;
  GRID_STRIDE(i, total) {
    const int xw = (int)(i % W);
    int64_t r = i / W;
    const int yh = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    dst[i] = (T)__half2float(src[(((int64_t)b * H + yh) * W + xw) * Cpad + c]);
  }
}

// one thread per 8-channel vector: 3 bytes in, Cpad halves out (pad channels zero)
__global__ void u8_to_nhwc_kernel(const uint8_t* __restrict__ src, int64_t npix, int Cpad, Norm3 nm,
                                  __half* __restrict__ dst) {
  const int CV = Cpad / 8;
  GRID_STRIDE(i, npix * CV) {
    const int64_t pix = i / CV;
    const int cv = (int)(i % CV);
    uint4 w = make_uint4(0, 0, 0, 0);
    if (cv == 0) {
      const uint8_t* s = src + pix * 3;
      const float r = ((float)s[0] / 255.0f - nm.mean[0]) * nm.inv_std[0];
      const float g = ((float)s[1] / 255.0f - nm.mean[1]) * nm.inv_std[1];
      const float b = ((float)s[2] / 255.0f - nm.mean[2]) * nm.inv_std[2];
      w.x = pack_half2(r, g);
      w.y = pack_half2(b, 0.f);
    }
    *reinterpret_cast<uint4*>(dst + pix * Cpad + cv * 8) = w;
  }
}

__global__ void nhwc_to_u8_kernel(const __half* __restrict__ src, int64_t npix, int Cpad, uint8_t* __restrict__ dst) {
  GRID_STRIDE(i, npix) {
    const __half* s = src + i * Cpad;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __half2float(s[c]) * 0.5f + 0.5f;
      v = fminf(fmaxf(v, 0.f), 1.f);
      dst[i * 3 + c] = (uint8_t)rintf(v * 255.0f);
    }
  }
}

// views [B, 4, S, S, 3] <-> tile [B, 2S, 2S, 3]; view k sits at (x, y) = ((k % 2) * S, (k / 2) * S)   (S = 256 upstream)
__global__ void tile_views_kernel(const uint8_t* __restrict__ views, uint8_t* __restrict__ tile, int B, int S,
                                  int to_tile) {
  const int S2 = 2 * S;
  const int64_t total = (int64_t)B * S2 * S2;
  GRID_STRIDE(i, total) {
    const int x = (int)(i % S2);
    const int y = (int)((i / S2) % S2);
    const int b = (int)(i / ((int64_t)S2 * S2));
    const int k = (y / S) * 2 + (x / S);
    const int64_t vi = ((((int64_t)b * 4 + k) * S + (y % S)) * S + (x % S)) * 3;
    const int64_t ti = i * 3;
    if (to_tile) {
      tile[ti] = views[vi];
      tile[ti + 1] = views[vi + 1];
      tile[ti + 2] = views[vi + 2];
    } else {
      const_cast<uint8_t*>(views)[vi] = tile[ti];
      const_cast<uint8_t*>(views)[vi + 1] = tile[ti + 1];
      const_cast<uint8_t*>(views)[vi + 2] = tile[ti + 2];
    }
  }
}

// out[b, t, :] = tok_emb[ids[b, t], :] + pos_emb[t, :]   (CLIP text embeddings)
__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, const __half* __restrict__ tok,
                                    const __half* __restrict__ pos, __half* __restrict__ out, int64_t rows, int T,
                                    int D, int vocab) {
  const int DV = D / 8;
  GRID_STRIDE(i, rows * DV) {
    const int64_t r = i / DV;
    const int dv = (int)(i % DV);
    int64_t id = ids[r];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(tok + id * D + dv * 8));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + (r % T) * D + dv * 8));
    const __half2* ha = reinterpret_cast<const __half2*>(&a);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
    uint4 w;
    uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 fa = __half22float2(ha[t]);
      const float2 fb = __half22float2(hb[t]);
      wp[t] = pack_half2(fa.x + fb.x, fa.y + fb.y);
    }
    *reinterpret_cast<uint4*>(out + r * D + dv * 8) = w;
  }
}

// FiLM folded into a FrozenBatchNorm affine: film = [gamma | beta] (2C), (1 + gamma) * (s * x + t) + beta
__global__ void film_fold_kernel(const float* __restrict__ film, const float* __restrict__ bn_scale,
                                 const float* __restrict__ bn_shift, float* __restrict__ scale_out,
                                 float* __restrict__ shift_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float g = 1.0f + film[c];
  scale_out[c] = g * bn_scale[c];
  shift_out[c] = g * bn_shift[c] + film[C + c];
}

}  // namespace gn

using namespace gn;

extern "C" int gn_embed_tokens(gn_handle* h, const void* ids_i64, const void* tok_emb, const void* pos_emb, int B,
                               int T, int D, int vocab, void* out, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, ids_i64 && tok_emb && pos_emb && out && B > 0 && T > 0 && D > 0 && (D % 8) == 0 && vocab > 0,
               "gn_embed_tokens: bad arguments");
  const int64_t rows = (int64_t)B * T;
  embed_tokens_kernel<<<grid_for(h, rows * (D / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const int64_t*>(ids_i64), static_cast<const __half*>(tok_emb), static_cast<const __half*>(pos_emb),
      static_cast<__half*>(out), rows, T, D, vocab);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_film_fold(gn_handle* h, const float* film, const float* bn_scale, const float* bn_shift,
                            float* scale_out, float* shift_out, int C, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, film && bn_scale && bn_shift && scale_out && shift_out && C > 0, "gn_film_fold: bad arguments");
  film_fold_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(film, bn_scale, bn_shift, scale_out,
                                                                                  shift_out, C);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_upsample_nearest2x(gn_handle* h, const void* x, int B, int H, int W, int C, void* y, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && y && B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0, "gn_upsample_nearest2x: bad arguments");
  const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(h, total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), B, H, W, C / 8);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_maxpool3x3s2(gn_handle* h, const void* x, int B, int H, int W, int C, void* y, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && y && B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0, "gn_maxpool3x3s2: bad arguments");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(h, total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), B, H, W, C);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_add(gn_handle* h, const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, a && b && out && n > 0 && (n % 2) == 0, "gn_add: bad arguments");
  add_kernel<<<grid_for(h, n / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half2*>(a), static_cast<const __half2*>(b), static_cast<__half2*>(out), n / 2);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_scale(gn_handle* h, const void* x, float s, void* y, int64_t n, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && y && n > 0, "gn_scale: bad arguments");
  scale_kernel<<<grid_for(h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x), s,
                                                                              static_cast<__half*>(y), n);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

__global__ void tanh_clamp_kernel(const __half* __restrict__ x, float mag, __half* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(tanhf(__half2float(x[i]) / mag) * mag);
}

extern "C" int gn_tanh_clamp(gn_handle* h, const void* x, float mag, void* y, int64_t n, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && y && n > 0 && mag > 0.f, "gn_tanh_clamp: bad arguments");
  tanh_clamp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), mag, static_cast<__half*>(y), n);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_timestep_embedding(gn_handle* h, float t, int dim, void* out, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, out && dim > 0 && (dim % 2) == 0, "gn_timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<(dim / 2 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      t, dim, static_cast<__half*>(out));
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_euler_step(gn_handle* h, const void* x, const void* eps, float sigma, float sigma_next, void* x_next,
                             void* x_scaled, int64_t n, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && eps && x_next && n > 0, "gn_euler_step: bad arguments");
  const float inv = 1.0f / sqrtf(sigma_next * sigma_next + 1.0f);
  euler_step_kernel<<<grid_for(h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<const __half*>(eps), sigma_next - sigma, inv,
      static_cast<__half*>(x_next), static_cast<__half*>(x_scaled), n);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_euler_ancestral_step(gn_handle* h, const void* x, const void* eps, const void* noise, float sigma,
                                       float sigma_down, float sigma_up, float sigma_next, void* x_next,
                                       void* x_scaled, int64_t n, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, x && eps && noise && x_next && n > 0, "gn_euler_ancestral_step: bad arguments");
  const float inv = 1.0f / sqrtf(sigma_next * sigma_next + 1.0f);
  euler_ancestral_step_kernel<<<grid_for(h, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<const __half*>(eps), static_cast<const __half*>(noise),
      sigma_down - sigma, sigma_up, inv, static_cast<__half*>(x_next), static_cast<__half*>(x_scaled), n);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_nchw_to_nhwc(gn_handle* h, const void* src, int src_fp32, int B, int C, int H, int W, int Cpad,
                               const float* mean3, const float* std3, void* dst, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, src && dst && B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "gn_nchw_to_nhwc: bad arguments");
  const int64_t total = (int64_t)B * H * W * Cpad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Norm3 nm;
  nm.enabled = (mean3 && std3) ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = nm.enabled ? mean3[c] : 0.f;
    nm.inv_std[c] = nm.enabled ? 1.0f / std3[c] : 1.f;
  }
  if (src_fp32 == 1)
    nchw_to_nhwc_kernel<float><<<grid_for(h, total), 256, 0, st>>>(static_cast<const float*>(src), B, C, H, W, Cpad,
                                                                   nm, static_cast<__half*>(dst));
  else if (src_fp32 == 2)
    nchw_to_nhwc_kernel<uint8_t><<<grid_for(h, total), 256, 0, st>>>(static_cast<const uint8_t*>(src), B, C, H, W,
                                                                     Cpad, nm, static_cast<__half*>(dst));
  else
    nchw_to_nhwc_kernel<__half><<<grid_for(h, total), 256, 0, st>>>(static_cast<const __half*>(src), B, C, H, W, Cpad,
                                                                    nm, static_cast<__half*>(dst));
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_nhwc_to_nchw(gn_handle* h, const void* src, int B, int C, int H, int W, int Cpad, void* dst,
                               int dst_fp32, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, src && dst && B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "gn_nhwc_to_nchw: bad arguments");
  const int64_t total = (int64_t)B * C * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_fp32)
    nhwc_to_nchw_kernel<float><<<grid_for(h, total), 256, 0, st>>>(static_cast<const __half*>(src), B, C, H, W, Cpad,
                                                                   static_cast<float*>(dst));
  else
    nhwc_to_nchw_kernel<__half><<<grid_for(h, total), 256, 0, st>>>(static_cast<const __half*>(src), B, C, H, W, Cpad,
                                                                    static_cast<__half*>(dst));
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_u8_to_nhwc(gn_handle* h, const void* src_u8, int B, int H, int W, int Cpad, const float* mean3,
                             const float* std3, void* dst, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, src_u8 && dst && B > 0 && H > 0 && W > 0 && Cpad >= 8 && (Cpad % 8) == 0,
               "gn_u8_to_nhwc: bad arguments");
  Norm3 nm;
  nm.enabled = 1;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3 ? mean3[c] : 0.f;   // host pointers: three floats each
    nm.inv_std[c] = std3 ? 1.0f / std3[c] : 1.f;
  }
  const int64_t npix = (int64_t)B * H * W;
  u8_to_nhwc_kernel<<<grid_for(h, npix * (Cpad / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src_u8), npix, Cpad, nm, static_cast<__half*>(dst));
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_nhwc_to_u8(gn_handle* h, const void* src, int B, int H, int W, int Cpad, void* dst_u8,
                             void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, src && dst_u8 && B > 0 && H > 0 && W > 0 && Cpad >= 3, "gn_nhwc_to_u8: bad arguments");
  const int64_t npix = (int64_t)B * H * W;
  nhwc_to_u8_kernel<<<grid_for(h, npix), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(src), npix, Cpad, static_cast<uint8_t*>(dst_u8));
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_tile_views(gn_handle* h, const void* views_u8, int B, int S, void* tile_u8, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, views_u8 && tile_u8 && B > 0 && S > 0, "gn_tile_views: bad arguments");
  tile_views_kernel<<<grid_for(h, (int64_t)B * 4 * S * S), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(views_u8), static_cast<uint8_t*>(tile_u8), B, S, 1);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}

extern "C" int gn_untile_views(gn_handle* h, const void* tile_u8, int B, int S, void* views_u8, void* stream) {
  if (!h) return GN_ERR_INVALID;
  ProfScope prof(h, stream, GN_PROF_ELEMENTWISE, 0.0, 0.0);
  GN_CHECK_ARG(h, views_u8 && tile_u8 && B > 0 && S > 0, "gn_untile_views: bad arguments");
  tile_views_kernel<<<grid_for(h, (int64_t)B * 4 * S * S), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(views_u8), static_cast<uint8_t*>(const_cast<void*>(tile_u8)), B, S, 0);
  GN_CHECK_LAUNCH(h);
  return GN_OK;
}
