// Attention kernels.
//
// attn_tc_kernel — tcgen05 flash attention, head_dim 64, fp16, one CTA per (128 queries, head, batch):
//   warp 0      TMA producer: Q tile once, K tiles (2-stage ring) and V tile (1 stage) per 128-key block
//   warp 1      single-thread MMA issuer:  S = Q K^T  (M128 N128 K64, K-major A/B)  -> TMEM cols [0,128)
//                                          O_blk = P V (M128 N64 K128, A = P from smem, B = V MN-major) -> TMEM [128,192)
//   warps 2-5   online softmax: thread r owns query row r (TMEM lane r): tcgen05.ld S, running max / sum in fp32,
//               exp2 with the softmax scale folded into one FFMA, P written to shared memory as fp16 in the
//               SWIZZLE_128B K-major layout the MMA expects, O accumulated in registers with the usual rescale.
// Two CTAs fit per SM (96 KiB smem, 256 TMEM columns each) so one CTA's MMAs overlap the other's softmax.
//
// attn_small_kernel — SIMT attention for tiny problems (ACT transformer, CLIP text towers): one warp per query.
#include "common.h"
#include "ptx.cuh"

namespace gn {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_D = 64;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;  // 16 KiB: Q, K and V tiles
constexpr int AT_P_BYTES = 128 * 128 * 2;    // 32 KiB
constexpr int AT_KSTAGES = 2;
constexpr int AT_THREADS = 192;
constexpr int AT_TMEM_COLS = 256;
constexpr int AT_SMEM_BYTES = AT_TILE_BYTES /*Q*/ + AT_P_BYTES + AT_KSTAGES * AT_TILE_BYTES + AT_TILE_BYTES /*V*/ + 256 + 1024;

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __half* out;
  int64_t ldo;
  int Tq, Tk;
  float scale_log2;  // softmax scale * log2(e)
};

__global__ void __launch_bounds__(AT_THREADS, 2) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + AT_TILE_BYTES;
  uint8_t* sK = sP + AT_P_BYTES;
  uint8_t* sV = sK + AT_KSTAGES * AT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int nblk = (p.Tk + AT_BKV - 1) / AT_BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_KSTAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, AT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // q / k / v come from the previous kernel in the stream
  const uint32_t tmem_s = tmem_base;        // 128 fp32 columns
  const uint32_t tmem_o = tmem_base + 128;  // 64 fp32 columns

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_arrive_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_2d(sQ, &p.tmQ, q_full, head * AT_D, batch * p.Tq + q0);
      for (int i = 0; i < nblk; ++i) {
        const int ks = i % AT_KSTAGES;
        mbar_wait(&k_empty[ks], ((i / AT_KSTAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], AT_TILE_BYTES);
        tma_load_2d(sK + ks * AT_TILE_BYTES, &p.tmK, &k_full[ks], head * AT_D, batch * p.Tk + i * AT_BKV);
        mbar_wait(v_empty, (i & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, AT_TILE_BYTES);
        tma_load_2d(sV, &p.tmV, v_full, head * AT_D, batch * p.Tk + i * AT_BKV);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc_qk = umma_idesc_f16(AT_BKV, 0, 0);  // N = 128 keys, both operands K-major
      const uint32_t idesc_pv = umma_idesc_f16(AT_D, 0, 1);    // N = 64 dims, B (= V) MN-major
      const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ), 1024, 0);
      auto issue_qk = [&](int i) {
        const int ks = i % AT_KSTAGES;
        mbar_wait(&k_full[ks], (i / AT_KSTAGES) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_desc_sw128(smem_u32(sK + ks * AT_TILE_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_f16_ss(tmem_s, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k > 0);
        umma_commit(s_full);
        umma_commit(&k_empty[ks]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(p_full, i & 1);
        mbar_wait(v_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < AT_BKV / 16; ++j) {
          // A: P[128 rows][16 k] slice j: 64-column sub-tile j/4, 32-byte step j%4 inside the swizzle row
          const uint64_t a_desc = umma_desc_sw128(smem_u32(sP + (j >> 2) * AT_TILE_BYTES), 1024, 0) + 2 * (j & 3);
          // B: V rows [16 j, 16 j + 16) x 64 dims, MN-major: two 8-row swizzle atoms 1024 B apart
          const uint64_t b_desc = umma_desc_sw128(smem_u32(sV + j * 2048), 1024, 1024);
          umma_f16_ss(tmem_o, a_desc, b_desc, idesc_pv, j > 0);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
        if (i + 1 < nblk) issue_qk(i + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps (2..5)
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    float m_run = -INFINITY;  // running max of raw scores
    float l_run = 0.f;
    float o_acc[AT_D];
#pragma unroll
    for (int j = 0; j < AT_D; ++j) o_acc[j] = 0.f;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const uint32_t swz = static_cast<uint32_t>(row & 7);

    for (int i = 0; i < nblk; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      const int kv0 = i * AT_BKV;
      const int nvalid = p.Tk - kv0;  // >= 1
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AT_BKV; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_s + lane_sel + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = (c + j < nvalid) ? __uint_as_float(r[j]) : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);  // m_run = -inf -> 0
      const float moff = m_new * p.scale_log2;
      float lsum = 0.f;
      // pass 2: probabilities -> fp16 -> swizzled smem
#pragma unroll 1
      for (int c = 0; c < AT_BKV; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_s + lane_sel + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = (c + j < nvalid) ? ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -moff)) : 0.f;
          float p1 = (c + j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2, -moff)) : 0.f;
          lsum += p0 + p1;
          pk[j >> 1] = pack_half2(p0, p1);
        }
        // 32 columns = 4 chunks of 16 bytes inside sub-tile (c / 64), chunk index ((c % 64) / 8 + t) ^ (row & 7)
        const uint32_t sub = p_row + (c >> 6) * AT_TILE_BYTES;
        const uint32_t chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t addr = sub + (((chunk0 + t) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * t]), "r"(pk[4 * t + 1]),
                       "r"(pk[4 * t + 2]), "r"(pk[4 * t + 3])
                       : "memory");
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      fence_proxy_async_smem();  // P (generic-proxy writes) must be visible to the tensor core (async proxy)
      tc_fence_before();         // our tcgen05.ld of S precede the MMA that overwrites S
      mbar_arrive(p_full);

      mbar_wait(o_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < AT_D; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_o + lane_sel + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o_acc[c + j] = fmaf(o_acc[c + j], alpha, __uint_as_float(r[j]));
      }
    }
    tc_fence_before();
    if (q0 + row < p.Tq) {
      const float inv = 1.0f / l_run;
      __half* op = p.out + ((int64_t)batch * p.Tq + q0 + row) * p.ldo + head * AT_D;
#pragma unroll
      for (int j = 0; j < AT_D; j += 8) {
        uint4 w;
        w.x = pack_half2(o_acc[j] * inv, o_acc[j + 1] * inv);
        w.y = pack_half2(o_acc[j + 2] * inv, o_acc[j + 3] * inv);
        w.z = pack_half2(o_acc[j + 4] * inv, o_acc[j + 5] * inv);
        w.w = pack_half2(o_acc[j + 6] * inv, o_acc[j + 7] * inv);
        *reinterpret_cast<uint4*>(op + j) = w;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ SIMT attention
// One warp per query row; keys processed 32 at a time (one key per lane), each lane owns D/32 output dims.
template <int D>
__global__ void __launch_bounds__(128) attn_small_kernel(const __half* __restrict__ q, int64_t ldq,
                                                         const __half* __restrict__ k, int64_t ldk,
                                                         const __half* __restrict__ v, int64_t ldv,
                                                         __half* __restrict__ out, int64_t ldo, int heads, int Tq,
                                                         int Tk, float scale, int causal) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= Tq) return;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  constexpr int DPL = D / 32;  // output dims per lane
  const __half* qp = q + ((int64_t)batch * Tq + qi) * ldq + head * D;
  float qf[D];
#pragma unroll
  for (int d = 0; d < D; d += 2) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(qp + d));
    qf[d] = f.x * scale;
    qf[d + 1] = f.y * scale;
  }
  float m_run = -INFINITY, l_run = 0.f;
  float o[DPL];
#pragma unroll
  for (int t = 0; t < DPL; ++t) o[t] = 0.f;
  const int kmax = causal ? min(Tk, qi + 1) : Tk;
  for (int kv0 = 0; kv0 < kmax; kv0 += 32) {
    const int kj = kv0 + lane;
    float s = -INFINITY;
    if (kj < kmax) {
      const __half* kp = k + ((int64_t)batch * Tk + kj) * ldk + head * D;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 8) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(kp + d));
        const __half2* hp = reinterpret_cast<const __half2*>(&w);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __half22float2(hp[t]);
          acc = fmaf(qf[d + 2 * t], f.x, acc);
          acc = fmaf(qf[d + 2 * t + 1], f.y, acc);
        }
      }
      s = acc;
    }
    float mx = s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float m_new = fmaxf(m_run, mx);
    const float alpha = __expf(m_run - m_new);
    const float pj = (kj < kmax) ? __expf(s - m_new) : 0.f;
    float ps = pj;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    l_run = l_run * alpha + ps;
    m_run = m_new;
#pragma unroll
    for (int t = 0; t < DPL; ++t) o[t] *= alpha;
    const int nk = min(32, kmax - kv0);
    for (int j = 0; j < nk; ++j) {
      const float pb = __shfl_sync(0xffffffffu, pj, j);
      const __half* vp = v + ((int64_t)batch * Tk + kv0 + j) * ldv + head * D;
#pragma unroll
      for (int t = 0; t < DPL; ++t) o[t] = fmaf(pb, __half2float(__ldg(vp + lane + 32 * t)), o[t]);
    }
  }
  const float inv = 1.0f / l_run;
  __half* op = out + ((int64_t)batch * Tq + qi) * ldo + head * D;
#pragma unroll
  for (int t = 0; t < DPL; ++t) op[lane + 32 * t] = __float2half_rn(o[t] * inv);
}

}  // namespace gn

using namespace gn;

extern "C" int gn_attention(gn_handle* h, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, void* out, int64_t ldo, int B, int heads, int Tq, int Tk, float scale,
                            void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, q && k && v && out, "gn_attention: null pointer");
  GN_CHECK_ARG(h, B > 0 && heads > 0 && Tq > 0 && Tk > 0, "gn_attention: bad shape");
  GN_CHECK_ARG(h, (ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 8) == 0 && (ldo % 8) == 0,
               "gn_attention: row strides must be multiples of 8 elements");
  ProfScope prof(h, stream, GN_PROF_ATTENTION, 4.0 * B * heads * (double)Tq * Tk * AT_D,
                 2.0 * B * heads * AT_D * (2.0 * Tq + 2.0 * Tk));
  static thread_local AttnParams p;
  memset(&p, 0, sizeof(p));
  const void* ptrs[3] = {q, k, v};
  const int64_t lds[3] = {ldq, ldk, ldv};
  const int rows[3] = {B * Tq, B * Tk, B * Tk};
  CUtensorMap* maps[3] = {&p.tmQ, &p.tmK, &p.tmV};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[2] = {(uint64_t)heads * AT_D, (uint64_t)rows[i]};
    uint64_t strides[1] = {(uint64_t)lds[i] * 2};
    uint32_t box[2] = {AT_D, 128};
    int rc = make_tmap_f16(h, maps[i], ptrs[i], 2, dims, strides, box);
    if (rc) return rc;
  }
  p.out = static_cast<__half*>(out);
  p.ldo = ldo;
  p.Tq = Tq;
  p.Tk = Tk;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (!h->attn_attr_set) {
    GN_CHECK_CUDA(h, cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
    h->attn_attr_set = true;
  }
  dim3 grid(ceil_div(Tq, AT_BQ), heads, B);
  GN_CHECK_CUDA(h, launch_ex(h, attn_tc_kernel, grid, dim3(AT_THREADS, 1, 1), AT_SMEM_BYTES,
                              static_cast<cudaStream_t>(stream), 1, p));
  h->launches++;
  return GN_OK;
}

extern "C" int gn_attention_small(gn_handle* h, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, void* out, int64_t ldo, int B, int heads, int head_dim, int Tq, int Tk,
                                  float scale, int causal, void* stream) {
  if (!h) return GN_ERR_INVALID;
  GN_CHECK_ARG(h, q && k && v && out, "gn_attention_small: null pointer");
  GN_CHECK_ARG(h, B > 0 && heads > 0 && Tq > 0 && Tk > 0, "gn_attention_small: bad shape");
  GN_CHECK_ARG(h, head_dim == 32 || head_dim == 64, "gn_attention_small: head_dim %d unsupported", head_dim);
  GN_CHECK_ARG(h, (ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 2) == 0, "gn_attention_small: bad row strides");
  ProfScope prof(h, stream, GN_PROF_ATTN_SMALL, 4.0 * B * heads * (double)Tq * Tk * head_dim,
                 2.0 * B * heads * head_dim * (2.0 * Tq + 2.0 * Tk));
  dim3 grid(ceil_div(Tq, 4), heads, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* qh = static_cast<const __half*>(q);
  const __half* kh = static_cast<const __half*>(k);
  const __half* vh = static_cast<const __half*>(v);
  __half* oh = static_cast<__half*>(out);
  if (head_dim == 32)
    GN_CHECK_CUDA(h, launch_ex(h, attn_small_kernel<32>, grid, dim3(128, 1, 1), 0, st, 1, qh, ldq, kh, ldk, vh, ldv, oh,
                                ldo, heads, Tq, Tk, scale, causal));
  else
    GN_CHECK_CUDA(h, launch_ex(h, attn_small_kernel<64>, grid, dim3(128, 1, 1), 0, st, 1, qh, ldq, kh, ldk, vh, ldv, oh,
                                ldo, heads, Tq, Tk, scale, causal));
  h->launches++;
  return GN_OK;
}
