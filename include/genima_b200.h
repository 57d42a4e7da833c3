/*
 * genima_b200 — C ABI of the B200-native (sm_100a) kernels behind Genima's per-step inference hot path.
 *
 * The reference (MohitShridhar/genima) has no FFI of its own: its hot path is two Python calls,
 *   controller/agent/sd_controlnet_agent.py:67-76   SDControlNetAgent.infer -> diffusers pipe(...)
 *   controller/method/genima_act.py:165-214          GenimaACTPolicy.forward
 * whose arithmetic runs inside diffusers 0.29.0 / RoboBase / torchvision library kernels (cuDNN, cuBLAS,
 * xformers).  Each entry point below replaces one class of those library kernels; the Python
 * shims in genima_b200/ (pipeline.py, act_policy.py) keep the reference call signatures and call ONLY these.
 *
 * Conventions
 *  - every function returns 0 on success, a negative gn_status otherwise; nothing throws across the ABI;
 *    gn_last_error(h) returns a human-readable message for the last failure on that handle.
 *  - all tensor arguments are DEVICE pointers owned by the caller (torch allocates them); activations are
 *    fp16, channels-last: images are NHWC = a row-major [B*H*W, C] matrix, token sequences are [B*T, C].
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, no hidden sync, so a call
 *    chain can be captured into a CUDA graph.
 *  - one handle per (device, host thread); not re-entrant.
 */
#ifndef GENIMA_B200_H
#define GENIMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gn_handle gn_handle;

enum gn_status {
  GN_OK = 0,
  GN_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  GN_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  GN_ERR_NOMEM = -3,     /* workspace too small / allocation failed */
  GN_ERR_NODRIVER = -4   /* no CUDA driver / device (CPU-only box) */
};

enum gn_act { GN_ACT_NONE = 0, GN_ACT_SILU = 1, GN_ACT_GELU = 2, GN_ACT_RELU = 3, GN_ACT_QUICKGELU = 4 };

/* Fused GEMM epilogue, applied to the fp32 accumulator acc[m, n]:
 *   v   = act_pre(acc * scale[n] + bias[n] + rowvec[m / rows_per_batch, n])
 *   out = act_post(alpha * v + beta * residual[m, n])          (alpha = 1, beta = 1 when residual is given)
 * geglu = 1: weight rows are packed in 64-wide (value, gate) blocks; out[m, j] = v_val * gelu(v_gate), N/2 columns.
 * Replaces the bias-add / time-embedding add / residual add / activation torch kernels that follow every
 * conv or linear in diffusers' ResnetBlock2D, BasicTransformerBlock, FeedForward(GEGLU) and torchvision BasicBlock. */
typedef struct gn_epilogue {
  const float* scale;
  const float* bias;
  const float* rowvec;
  const void* residual; /* fp16 [M, ldr] */
  int64_t ldr;
  int32_t rows_per_batch;
  int32_t act_pre;
  int32_t act_post;
  float alpha;
  float beta;
  int32_t geglu;
  int32_t out_fp32; /* 0: fp16 output, 1: fp32 output */
  /* LayerNorm folded into this GEMM (gn_linear only; BasicTransformerBlock's norm1/2/3 -> to_q/k/v, to_q, ff.net.0):
   * A holds the UN-normalised rows x, W is pre-multiplied by the LayerNorm gamma, and the epilogue applies
   *   acc' = rstd[m] * (acc - mean[m] * ln_colsum[n]) + bias[n],  ln_colsum[n] = sum_k W[n, k],
   *   bias[n] = original bias + sum_k beta[k] * W_unscaled[n, k]
   * with (mean, rstd) of row m computed from `ln_parts` (sum, sumsq) partials per row written by the GEMM that
   * produced x (its rowstats_out).  `scale` must be NULL.  The LayerNorm kernel launch disappears. */
  const void* ln_stats;   /* float2 [M][ln_parts] */
  const float* ln_colsum; /* [N] */
  int32_t ln_parts;
  float ln_eps;
  /* Row statistics of this GEMM's fp16 output: (sum, sumsq) partials, one per (n-tile, K-split, column share), written
   * as float2 [M][parts] with parts = gn_get_last_rowstats_parts() <= rowstats_capacity (tile configurations that
   * would need more partials are not considered). */
  void* rowstats_out;
  int32_t rowstats_capacity;
  /* GroupNorm statistics of this GEMM's fp16 output for the gn_group_norm_apply that normalises it (ResnetBlock2D
   * norm1/norm2, Transformer2DModel.norm, conv_norm_out): channels are grouped in buckets of `gn_bucket` (even, divides
   * the output columns and the consumer's channels-per-group) and the (sum, sumsq) of every (image, bucket) is ADDED,
   * as 2^20-scaled 64-bit fixed point, to gnstats_out = uint64 [images][N_out / gn_bucket][2], which the caller zeroes
   * beforehand.  Integer accumulation makes the result independent of CTA arrival order (bit-reproducible).  Image of
   * row m = m / rows_per_batch; rows_per_batch must be a multiple of 128, or a power of two in [16, 64].  The separate
   * statistics pass over HBM (and its grid barrier) disappears. */
  int32_t gn_bucket;
  void* gnstats_out;
  /* gn_linear only: W is not a constant weight matrix but the output of an earlier kernel in the stream (attention
   * scores Q K^T, P V).  Constant weights are requested by the TMA producer BEFORE the programmatic-dependent-launch
   * wait, so that (cold, HBM-resident) weight tiles stream in while the previous kernel drains. */
  int32_t w_dynamic;
  int32_t reserved;
} gn_epilogue;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
int gn_create(int device, gn_handle** out);
int gn_destroy(gn_handle* h);
const char* gn_last_error(const gn_handle* h);
const char* gn_version(void);
/* Caller-owned device scratch (>= 140 MiB to enable L2-cold timing in the autotuner: it is overwritten to evict the
 * L2 before weight-heavy candidates are timed).  Split-K needs no workspace: partial sums are reduced through
 * distributed shared memory inside a thread-block cluster. */
int gn_set_workspace(gn_handle* h, void* dptr, int64_t bytes);
/* Force tile width / split count of the next GEMM-class calls (0 = heuristic); used by tests and tuning. */
int gn_set_gemm_tuning(gn_handle* h, int block_n, int splits);
/* Programmatic dependent launch (default on): gn_linear / gn_conv2d / gn_attention* / gn_group_norm / gn_layer_norm are
 * launched with cudaLaunchAttributeProgrammaticStreamSerialization and execute griddepcontrol.wait before their first
 * global-memory access, so their set-up (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of the
 * previous kernel in the stream, also inside a captured CUDA graph.  enable = 0 restores ordinary launches; enable = 2
 * keeps the dependent launch but switches the early weight prefetch of the GEMM producer off (A/B). */
int gn_set_pdl(gn_handle* h, int enable);
/* GEMM epilogues stage the finished fp16 tile in shared memory and write it with one TMA store per sub-tile (the
 * residual tile is prefetched the same way while the MMAs run); enable = 0 restores per-thread global stores (A/B). */
int gn_set_staged_epilogue(gn_handle* h, int enable);
/* gn_group_norm synchronises its CTAs with a grid barrier, so all of them must be co-resident: the grid never exceeds
 * the SM count.  When two streams may each run a gn_group_norm at the same time (each through its OWN handle), cap both
 * at half the SMs with this call; 0 restores the default. */
int gn_set_gn_max_ctas(gn_handle* h, int max_ctas);
/* enable = 1: the first call of gn_linear / gn_conv2d for a new problem shape times the tile-configuration candidates of
 * the cost model (CUDA events on the caller's stream, a few extra launches writing the same output) and caches the
 * fastest; later calls, and calls made while the stream is being captured into a CUDA graph, use the cache.
 * enable = 0 keeps the pure model; enable = -1 also clears the cache. */
int gn_set_autotune(gn_handle* h, int enable);
/* The measured tile configurations as text ("key=block_n,splits,stages,packed" lines, sorted).  Export returns the number
 * of bytes needed (nothing is written unless cap is large enough; call with buf = NULL to size the buffer).  Import adds
 * the lines to the handle's cache (replace != 0: clears it first).  Rank 0 of a multi-GPU evaluation tunes once and
 * broadcasts the text with the weights, so every rank launches the same tile / split-K configurations and therefore sums
 * in the same order: sharding the episodes of controller/eval_genima.py:115-142 over GPUs does not change any result. */
int64_t gn_tune_cache_export(const gn_handle* h, char* buf, int64_t cap);
int gn_tune_cache_import(gn_handle* h, const char* buf, int64_t n, int replace);
/* CTA pairs (tcgen05 cta_group::2): two CTAs on neighbouring SMs compute two consecutive 128-row m-tiles with ONE M = 256
 * MMA per k-step; each fetches its own A tile and half of the W tile, so the weight bytes an SM ingests per output tile
 * halve (batch-1 GEMMs are bound by L2 -> SM operand traffic, not by the tensor pipe).  mode 0 = never, 1 = a candidate
 * of the tile search (default), 2 = wherever the shape allows it (tests, A/B).  gn_last_gemm_pair reports whether the last
 * gn_linear / gn_conv2d launch used pairs. */
int gn_set_gemm_pair(gn_handle* h, int mode);
int gn_last_gemm_pair(const gn_handle* h);
/* Force the operand-ring sizing of the next GEMM-class calls for 1 or 2 resident CTAs per SM (0 = heuristic). */
int gn_set_gemm_occupancy(gn_handle* h, int ctas_per_sm);
/* gn_attention KV split across a 2-CTA cluster (partials merged through distributed shared memory): 0 = never,
 * 1 = when it shortens the critical path (default; e.g. 160 CTAs x 32 KV blocks on 148 SMs), 2 = whenever Tk spans at
 * least two 128-key blocks (tests). */
int gn_set_attention_kv_split(gn_handle* h, int mode);
/* Debug aid: when dptr (device uint64[8]) is non-NULL, CTA (0,0,0) of every following GEMM-class launch writes
 * %globaltimer stamps of its phases: 0 start, 1 set-up done, 2 first TMA issued, 3 first operands landed, 4 all MMAs
 * issued, 5 accumulator complete, 6 epilogue done, 7 exit.  NULL switches it off. */
int gn_set_gemm_trace(gn_handle* h, void* dptr_u64x8);
/* Last launch configuration chosen by gn_linear / gn_conv2d: out[0]=block_n, out[1]=splits, out[2]=stages, out[3]=ctas. */
int gn_get_last_gemm_config(const gn_handle* h, int32_t* out4);
/* Partials per row the last gn_linear call with rowstats_out wrote (the consumer's ln_parts). */
int gn_get_last_rowstats_parts(const gn_handle* h);
/* Number of kernels launched through this handle since creation (bench.py's gpu_launches). */
int64_t gn_launch_count(const gn_handle* h);

/* Per-call timing with CUDA events on the caller's stream (bench.py's roofline figures).  Between gn_profile_begin and
 * gn_profile_end every compute entry point records an event pair around its launches; gn_profile_end synchronises and
 * returns, per class, the summed milliseconds, call count, algorithmic FLOPs (2 x MACs on real channels) and
 * algorithmic bytes (external inputs + outputs + weights, each once).  Arrays hold GN_PROF_NUM_CLASSES entries.
 * Not usable while the stream is being captured into a CUDA graph. */
enum gn_prof_class {
  GN_PROF_LINEAR = 0,     /* gn_linear (gemm_tc_kernel [+ splitk_reduce_kernel]) */
  GN_PROF_CONV = 1,       /* gn_conv2d (gemm_tc_kernel, implicit GEMM [+ splitk_reduce_kernel]) */
  GN_PROF_ATTENTION = 2,  /* gn_attention (attn_tc_kernel) */
  GN_PROF_ATTN_SMALL = 3, /* gn_attention_small */
  GN_PROF_NORM = 4,       /* gn_group_norm, gn_layer_norm, gn_softmax_rows */
  GN_PROF_ELEMENTWISE = 5,
  GN_PROF_NUM_CLASSES = 6
};
int gn_profile_begin(gn_handle* h);
int gn_profile_end(gn_handle* h, double* ms, int64_t* calls, double* flops, double* bytes);

/* ---- dense contractions (tcgen05.mma, TMA-staged operands, TMEM accumulators) --------------------------------
 * out[M, N] = epilogue(A[M, K] @ W[N, K]^T).  A: fp16 row-major with row stride lda (elements, % 8 == 0), K % 8 == 0.
 * W: fp16 [N, K] row-major (torch nn.Linear layout).  Replaces cuBLAS GEMMs behind nn.Linear / 1x1 conv. */
int gn_linear(gn_handle* h, const void* A, int64_t lda, int M, int K, const void* W, int N, void* out, int64_t ldo,
              const gn_epilogue* epi, void* stream);

/* Implicit-GEMM convolution on NHWC fp16.  x: [B, H, W, C] (C % 8 == 0); w: packed [Cout][KH*KW][Cp] followed, per
 * output channel, by [Cp_e0][Cp_e1] for up to two fused 1x1 "extra" sources (Cp = C rounded up to 64) — see
 * genima_b200/packing.py.  The extra sources (ex0/ex1, NHWC at the OUTPUT resolution) implement ResnetBlock2D's
 * conv_shortcut over the (possibly concatenated) block input inside the same accumulation.
 * stride in {1, 2}.  out: [B, Ho, Wo, Cout] with row stride ldo.  Replaces cuDNN implicit-GEMM convolutions. */
int gn_conv2d(gn_handle* h, const void* x, int B, int H, int W, int C, const void* w, int Cout, int KH, int KW,
              int stride, int pad, const void* ex0, int C_ex0, const void* ex1, int C_ex1, void* out, int64_t ldo,
              const gn_epilogue* epi, void* stream);

/* gn_conv2d with separate top / left / bottom / right zero padding and no extra sources.  diffusers' VAE-encoder
 * Downsample2D is F.pad(x, (0, 1, 0, 1)) followed by a 3x3 stride-2 convolution with padding 0 (reached through
 * StableDiffusionInstructPix2PixPipeline.prepare_image_latents, controller/agent/sd_pix2pix_agent.py:52-60): here the
 * padding is TMA out-of-bounds zero fill, no padded copy of x exists.
 * out: [B, (H + pt + pb - KH) / stride + 1, (W + pl + pr - KW) / stride + 1, Cout]. */
int gn_conv2d_asym(gn_handle* h, const void* x, int B, int H, int W, int C, const void* w, int Cout, int KH, int KW,
                   int stride, int pad_top, int pad_left, int pad_bottom, int pad_right, void* out, int64_t ldo,
                   const gn_epilogue* epi, void* stream);

/* out[B, 2H, 2W, Cout] = conv3x3(pad 1)(nearest-neighbour x2 upsample of x[B, H, W, C]) WITHOUT materialising the
 * upsampled tensor (diffusers Upsample2D = F.interpolate(scale 2, "nearest") + conv): every output parity (py, px) is a
 * 2x2 convolution over x whose taps are sums of the 3x3 taps that fall on the same input pixel, so the op is four
 * implicit GEMMs with 4/9 of the multiply-adds, each TMA-storing its pixels with stride 2.  w4: four packed
 * [Cout][2*2][Cp] matrices, phase (py * 2 + px) major (packing.pack_upsample_conv_weight).  GroupNorm statistics
 * (gnstats_out) accumulate across the four launches.  Replaces upsample2x + cuDNN convolution. */
int gn_conv2d_up2x(gn_handle* h, const void* x, int B, int H, int W, int C, const void* w4, int Cout, void* out,
                   int64_t ldo, const gn_epilogue* epi, void* stream);

/* ---- attention (tcgen05 flash attention, head_dim 64) -------------------------------------------------------
 * q: rows [B*Tq] with row stride ldq, head h at columns [h*64, h*64+64); k, v likewise with Tk rows per batch.
 * out[B*Tq, heads*64] = softmax(q k^T * scale) v.   Replaces xformers memory_efficient_attention / torch SDPA. */
int gn_attention(gn_handle* h, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                 void* out, int64_t ldo, int B, int heads, int Tq, int Tk, float scale, void* stream);
/* Generic SIMT attention for small problems (ACT transformer: head_dim 32; any Tk); fp16 in/out, fp32 math.
 * Optional causal mask (CLIP text towers).  Replaces nn.MultiheadAttention's SDPA core. */
/* Cross-attention with the query projection inside the kernel (diffusers Attention.to_q + scaled_dot_product_attention of
 * BasicTransformerBlock.attn2, with norm2 folded): q = LayerNorm-folded(x W_q^T + bias) is computed per (128 queries, head)
 * CTA from x [B * Tq][C] (row stride ldx) and W_q [heads * 64][C]; ln_stats / ln_parts / ln_colsum as in gn_epilogue
 * (NULL ln_stats: plain projection).  k, v, out as in gn_attention.  One launch and one trip of Q through memory less. */
int gn_attention_qproj(gn_handle* h, const void* x, int64_t ldx, int C, const void* wq, const float* bias,
                       const void* ln_stats, int ln_parts, float ln_eps, const float* ln_colsum, const void* k, int64_t ldk,
                       const void* v, int64_t ldv, void* out, int64_t ldo, int B, int heads, int Tq, int Tk, float scale,
                       void* stream);
int gn_attention_small(gn_handle* h, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                       int64_t ldv, void* out, int64_t ldo, int B, int heads, int head_dim, int Tq, int Tk,
                       float scale, int causal, void* stream);

/* ---- normalisation --------------------------------------------------------------------------------------- */
/* GroupNorm over NHWC fp16, optional fused SiLU, optional channel-concat of two sources (x1 may be NULL).
 * y[B, HW, C0+C1] = act(gn(concat(x0, x1)) * gamma + beta).  Statistics in fp32.  If stats_in != NULL the (sum, sumsq)
 * pairs produced by a GEMM epilogue are used instead of a reduction pass.  Replaces torch group_norm + silu + cat. */
int gn_group_norm(gn_handle* h, const void* x0, int C0, const void* x1, int C1, int B, int HW, int groups, float eps,
                  const float* gamma, const float* beta, int silu, const float* stats_in, void* y, void* stream);
/* GroupNorm from statistics accumulated by the producing GEMM epilogues (gn_epilogue.gnstats_out): one elementwise pass,
 * no reduction, no grid barrier.  stats0 / stats1: uint64 [B][C0 / bucket][2] / [B][C1 / bucket][2] of x0 / x1 (x1, stats1
 * NULL without a concat); bucket must divide C0 and (C0 + C1) / groups.  Same result contract as gn_group_norm. */
int gn_group_norm_apply(gn_handle* h, const void* x0, int C0, const void* stats0, const void* x1, int C1,
                        const void* stats1, int bucket, int B, int HW, int groups, float eps, const float* gamma,
                        const float* beta, int silu, void* y, void* stream);
/* LayerNorm over the last dim of a [rows, C] fp16 matrix (fp32 statistics). */
int gn_layer_norm(gn_handle* h, const void* x, int64_t ldx, int rows, int C, float eps, const float* gamma,
                  const float* beta, void* y, int64_t ldy, void* stream);
/* y[rows, cols] (fp16) = softmax(scale * x[rows, cols]) row-wise; x is fp32 (x_fp32 = 1, e.g. attention scores from
 * gn_linear with out_fp32) or fp16 (may then alias y).  Replaces torch.softmax inside the VAE mid-block attention
 * (diffusers Attention with 1 head, d = 512). */
int gn_softmax_rows(gn_handle* h, const void* x, int x_fp32, int64_t ldx, void* y, int64_t ldy, int rows, int cols,
                    float scale, void* stream);

/* ---- data movement / elementwise --------------------------------------------------------------------------- */
/* y[B, 2H, 2W, C] = nearest-neighbour x2 upsample of x[B, H, W, C] (diffusers Upsample2D before its conv). */
int gn_upsample_nearest2x(gn_handle* h, const void* x, int B, int H, int W, int C, void* y, void* stream);
/* 3x3 stride-2 pad-1 max pool on NHWC fp16 (torchvision ResNet stem). */
int gn_maxpool3x3s2(gn_handle* h, const void* x, int B, int H, int W, int C, void* y, void* stream);
/* out = a + b (fp16, n elements). */
int gn_add(gn_handle* h, const void* a, const void* b, void* out, int64_t n, void* stream);
/* Sinusoidal timestep embedding (flip_sin_to_cos=True, freq_shift=0): out[dim] fp16 = [cos(t f), sin(t f)]. */
int gn_timestep_embedding(gn_handle* h, float t, int dim, void* out, void* stream);
/* Euler-discrete scheduler (diffusers EulerDiscreteScheduler.step, s_churn = 0), fp32 math on fp16 storage:
 *   x_next = x + (sigma_next - sigma) * eps;   x_scaled = x_next / sqrt(sigma_next^2 + 1)  (input of the next step).
 * x / eps / x_next / x_scaled: [n] fp16 element-wise; x_scaled may be NULL. */
int gn_euler_step(gn_handle* h, const void* x, const void* eps, float sigma, float sigma_next, void* x_next,
                  void* x_scaled, int64_t n, void* stream);
/* Euler-ancestral scheduler (diffusers EulerAncestralDiscreteScheduler.step, what stabilityai/sdxl-turbo ships; reached
 * through controller/agent/sdxl_controlnet_agent.py:66-75), epsilon prediction:
 *   x_next = x + (sigma_down - sigma) * eps + sigma_up * noise;   x_scaled = x_next / sqrt(sigma_next^2 + 1)
 * with sigma_up / sigma_down the host scalars of the step and `noise` [n] fp16 drawn by the caller's generator. */
int gn_euler_ancestral_step(gn_handle* h, const void* x, const void* eps, const void* noise, float sigma,
                            float sigma_down, float sigma_up, float sigma_next, void* x_next, void* x_scaled,
                            int64_t n, void* stream);
/* y = x * s (fp16), used for scale_model_input at step 0 and latents / scaling_factor before the VAE. */
int gn_scale(gn_handle* h, const void* x, float s, void* y, int64_t n, void* stream);
/* y = tanh(x / mag) * mag (fp16, n elements): AutoencoderTiny's soft clamp of the latents before its decoder
 * (diffusers DecoderTiny.forward, reached when eval_cfg.autoencoder selects TAESD: sd_controlnet_agent.py:45-49). */
int gn_tanh_clamp(gn_handle* h, const void* x, float mag, void* y, int64_t n, void* stream);
/* NCHW fp16/fp32/uint8 <-> NHWC fp16 with channel padding (pad channels zero-filled). src_fp32: source element type,
 * 0 = fp16, 1 = fp32, 2 = uint8 (the camera frames of the reference's obs dict, [T, 3, 256, 256] u8).
 * mean3/std3 (host pointers, may be NULL): channels 0..2 become (src / 255 - mean) / std — GenimaACTPolicy.forward's
 * `self.normalize(image / 255.0)` (controller/method/genima_act.py:188) fused into the layout change. */
int gn_nchw_to_nhwc(gn_handle* h, const void* src, int src_fp32, int B, int C, int H, int W, int Cpad,
                    const float* mean3, const float* std3, void* dst, void* stream);
int gn_nhwc_to_nchw(gn_handle* h, const void* src, int B, int C, int H, int W, int Cpad, void* dst, int dst_fp32,
                    void* stream);
/* uint8 NHWC RGB [B, H, W, 3] -> fp16 NHWC [B, H, W, Cpad] = (u8 / 255 - mean[c]) / std[c]; (mean, std) = (0, 1) gives
 * VaeImageProcessor.preprocess(do_normalize=False); ImageNet constants give GenimaACTPolicy's normalise. */
int gn_u8_to_nhwc(gn_handle* h, const void* src_u8, int B, int H, int W, int Cpad, const float* mean3,
                  const float* std3, void* dst, void* stream);
/* VaeImageProcessor.postprocess: fp16 NHWC [B, H, W, Cpad] in [-1, 1] -> uint8 [B, H, W, 3] = round(clamp(x/2+.5)*255). */
int gn_nhwc_to_u8(gn_handle* h, const void* src, int B, int H, int W, int Cpad, void* dst_u8, void* stream);
/* tile_images / untile_images of controller/utils/misc.py:6-47 on device: 4 views u8 [B, 4, S, S, 3] <-> one
 * [B, 2S, 2S, 3] tile (view k -> quadrant (k % 2, k / 2)); S = 256 in the reference. */
int gn_tile_views(gn_handle* h, const void* views_u8, int B, int S, void* tile_u8, void* stream);
int gn_untile_views(gn_handle* h, const void* tile_u8, int B, int S, void* views_u8, void* stream);

/* CLIP text embeddings: out[b, t, :] = tok_emb[ids[b, t], :] + pos_emb[t, :]; ids int64 [B, T] on device. */
int gn_embed_tokens(gn_handle* h, const void* ids_i64, const void* tok_emb, const void* pos_emb, int B, int T, int D,
                    int vocab, void* out, void* stream);
/* Fold FiLM (film = [gamma | beta], fp32 [2C]) into a FrozenBatchNorm affine (scale, shift fp32 [C]):
 * scale_out = (1 + gamma) * bn_scale, shift_out = (1 + gamma) * bn_shift + beta  (ACT ResNet-18 BasicBlock). */
int gn_film_fold(gn_handle* h, const float* film, const float* bn_scale, const float* bn_shift, float* scale_out,
                 float* shift_out, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENIMA_B200_H */
