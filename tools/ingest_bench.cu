// Micro-benchmark: how fast can one SM (and the whole chip) pull GEMM operand tiles into shared memory?
//   mode 0  W tile = TMA 2-D box {64 fp16, bn rows} of a row-major [N][K] matrix (rows 2*K bytes apart)   <- gemm.cu today
//   mode 1  W tile = TMA 2-D box {64, bn} of a dense panel [rows][64] (the bn rows are contiguous: bn * 128 bytes)
//   mode 2  W tile = ONE cp.async.bulk of bn * 128 contiguous bytes (pre-swizzled panel image)
// The A tile (TMA 2-D box {64, arows} of a row-major [M][K] activation) is fetched alongside when arows > 0.
// The consumer only waits for the data and releases the stage (an infinitely fast tensor core), so the figure is the
// ingest ceiling.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ingest_bench tools/ingest_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../genima_b200/csrc/ptx.cuh"

using namespace gn;

struct Params {
  CUtensorMap tmW, tmA;
  const uint8_t* wbase;
  int mode, bn, arows, iters, stages, kblocks_total, ntiles, passes, nmat;
  size_t mat_bytes;
  long long* clk;  // optional: per-operation cycle totals of CTA 0's producer / consumer threads
};

__global__ void __launch_bounds__(64, 1) ingest_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int wb = p.bn * 128, ab = p.arows * 128;
  uint8_t* sw = smem;
  uint8_t* sa = smem + (size_t)p.stages * wb;
  uint64_t* full = reinterpret_cast<uint64_t*>(sa + (size_t)p.stages * ab);
  uint64_t* empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int cta = blockIdx.x;
  const int tile = cta % p.ntiles, split = cta / p.ntiles;
  const int kb0 = split * p.iters;
  const int total = p.iters * p.passes;
  // (no runtime divisions inside the loops: a single thread issues everything, ~100 clk per integer division would
  // bound the loop long before the memory system does)
  if (threadIdx.x == 0) {
    int it = 0, pass = 0, mat = 0, s = 0;
    uint32_t ph = 1;
    long long c_wait = 0, c_arr = 0, c_tma = 0;
    for (int g = 0; g < total; ++g) {
      const long long t0 = clock64();
      mbar_wait(&empty[s], ph);
      const long long t1 = clock64();
      mbar_arrive_expect_tx(&full[s], wb + ab);
      const long long t2 = clock64();
      c_wait += t1 - t0;
      c_arr += t2 - t1;
      const int kb = kb0 + it;
      if (p.mode == 0) {
        tma_load_2d(sw + (size_t)s * wb, &p.tmW, &full[s], kb * 64, mat * 1280 + tile * p.bn);
      } else if (p.mode == 1) {
        tma_load_2d(sw + (size_t)s * wb, &p.tmW, &full[s], 0, ((mat * p.ntiles + tile) * p.kblocks_total + kb) * p.bn);
      } else {
        bulk_load(smem_u32(sw + (size_t)s * wb),
                  p.wbase + (size_t)mat * p.mat_bytes + ((size_t)tile * p.kblocks_total + kb) * wb, wb, &full[s]);
      }
      if (ab) tma_load_2d(sa + (size_t)s * ab, &p.tmA, &full[s], kb * 64, 0);
      c_tma += clock64() - t2;
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
      if (++it == p.iters) {
        it = 0;
        ++pass;
        if (++mat == p.nmat) mat = 0;
      }
    }
    if (p.clk && blockIdx.x == 0) {
      p.clk[0] = c_wait;
      p.clk[1] = c_arr;
      p.clk[2] = c_tma;
    }
  } else if (threadIdx.x == 32) {
    int s = 0;
    uint32_t ph = 0;
    long long c_wait = 0, c_arr = 0;
    for (int g = 0; g < total; ++g) {
      const long long t0 = clock64();
      mbar_wait(&full[s], ph);
      const long long t1 = clock64();
      mbar_arrive(&empty[s]);
      c_wait += t1 - t0;
      c_arr += clock64() - t1;
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }
    if (p.clk && blockIdx.x == 0) {
      p.clk[3] = c_wait;
      p.clk[4] = c_arr;
    }
  }
}

// Several producer warps: warp w (lane 0) issues the loads of stages w, w + P, ... (stages % P == 0); the consumer is the
// last warp.  Does the TMA issue cost (~145 clk per instruction of the issuing thread) overlap across warps?
__global__ void __launch_bounds__(160, 1) ingest_mp_kernel(const __grid_constant__ Params p, int nprod) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int wb = p.bn * 128;
  uint8_t* sw = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(sw + (size_t)p.stages * wb);
  uint64_t* empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % p.ntiles;
  const int total = p.iters * p.passes;
  if (warp < nprod && lane == 0) {
    int s = warp, it = warp % p.iters;
    uint32_t ph = 1;
    for (int g = warp; g < total; g += nprod) {
      mbar_wait(&empty[s], ph);
      mbar_arrive_expect_tx(&full[s], wb);
      tma_load_2d(sw + (size_t)s * wb, &p.tmW, &full[s], it * 64, tile * p.bn);
      s += nprod;
      if (s >= p.stages) {
        s -= p.stages;
        ph ^= 1;
      }
      it += nprod;
      if (it >= p.iters) it -= p.iters;
    }
  } else if (warp == 4 && lane == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int g = 0; g < total; ++g) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make2d(EncodeFn enc, CUtensorMap* m, void* base, uint64_t cols, uint64_t rows, uint64_t row_bytes, uint32_t bx,
                   uint32_t by) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t str[1] = {row_bytes};
  cuuint32_t box[2] = {bx, by};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
}

// Second experiment (argv[1] == "lat"): L2-warm, W only -- how does the time per stage depend on the ring depth, the box
// height and the number of CTAs?  (throughput = stages in flight / round-trip latency until something else saturates)
static int latency_sweep(EncodeFn enc, double clk) {
  const int K = 4096, KB = K / 64, N = 4096;
  void* w;
  cudaMalloc(&w, (size_t)N * K * 2);
  cudaMemset(w, 1, (size_t)N * K * 2);
  long long* dclk;
  cudaMalloc(&dclk, 64);
  printf("%5s %4s %4s %5s | %9s %10s %9s %9s | clk/stage: P.wait P.arrive P.tma C.wait C.arrive\n", "ctas", "bn", "stg",
         "mode", "us", "ns/stage", "B/clk/SM", "TB/s");
  for (int ctas : {1, 16, 74, 148}) {
    for (int bn : {32, 64, 128, 256}) {
      for (int stages : {1, 2, 4, 8, 12}) {
        for (int mode : {0, 2}) {
          if ((size_t)stages * bn * 128 > 200 * 1024) continue;
          Params p;
          memset(&p, 0, sizeof(p));
          p.mode = mode;
          p.bn = bn;
          p.arows = 0;
          p.iters = KB;
          p.kblocks_total = KB;
          p.ntiles = N / bn < 16 ? N / bn : 16;  // CTAs cycle over 16 n-tiles: 16 * bn * 8 KB <= 32 MB, L2 resident
          p.wbase = (const uint8_t*)w;
          p.passes = 8;
          p.nmat = 1;
          p.mat_bytes = 0;
          p.stages = stages;
          p.clk = dclk;
          if (mode == 0) make2d(enc, &p.tmW, w, K, (uint64_t)N, (uint64_t)K * 2, 64, bn);
          make2d(enc, &p.tmA, w, K, 128, (uint64_t)K * 2, 64, 64);
          const size_t smem = (size_t)stages * bn * 128 + stages * 16 + 2048;
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            ingest_kernel<<<ctas, 64, smem>>>(p);
            cudaEventRecord(e0);
            ingest_kernel<<<ctas, 64, smem>>>(p);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          cudaError_t err = cudaGetLastError();
          if (err != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(err));
            return 1;
          }
          const double us = best * 1e3;
          const double nst = (double)p.iters * p.passes;
          const double bytes_cta = nst * bn * 128;
          long long hc[5];
          cudaMemcpy(hc, dclk, sizeof(hc), cudaMemcpyDeviceToHost);
          printf("%5d %4d %4d %5d | %9.2f %10.1f %9.1f %9.2f | %6.0f %6.0f %6.0f %6.0f %6.0f\n", ctas, bn, stages, mode, us,
                 (us - 3.0) * 1e3 / nst, bytes_cta / ((us - 3.0) * 1e-6) / clk, bytes_cta * ctas / (us * 1e-6) / 1e12,
                 hc[0] / nst, hc[1] / nst, hc[2] / nst, hc[3] / nst, hc[4] / nst);
        }
      }
    }
  }
  printf("multi-producer: %5s %4s %4s %5s | %9s %10s %9s\n", "ctas", "bn", "stg", "nprod", "us", "ns/stage", "B/clk/SM");
  cudaFuncSetAttribute(ingest_mp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int ctas : {1, 148}) {
    for (int bn : {32, 96, 160, 256}) {
      for (int nprod : {1, 2, 4}) {
        Params p;
        memset(&p, 0, sizeof(p));
        p.bn = bn;
        p.iters = KB;
        p.kblocks_total = KB;
        p.ntiles = 16;
        p.passes = 8;
        p.stages = bn == 256 ? 4 : 8;
        make2d(enc, &p.tmW, w, K, (uint64_t)N, (uint64_t)K * 2, 64, bn);
        const size_t smem = (size_t)p.stages * bn * 128 + p.stages * 16 + 2048;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          ingest_mp_kernel<<<ctas, 160, smem>>>(p, nprod);
          cudaEventRecord(e0);
          ingest_mp_kernel<<<ctas, 160, smem>>>(p, nprod);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) {
          printf("error: %s\n", cudaGetErrorString(err));
          return 1;
        }
        const double us = best * 1e3, nst = (double)p.iters * p.passes;
        printf("multi-producer: %5d %4d %4d %5d | %9.2f %10.1f %9.1f\n", ctas, bn, p.stages, nprod, us,
               (us - 3.0) * 1e3 / nst, nst * bn * 128 / ((us - 3.0) * 1e-6) / clk);
      }
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'l') {
    void* fn0 = nullptr;
    cudaDriverEntryPointQueryResult q0;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn0, cudaEnableDefault, &q0);
    cudaDeviceProp prop0;
    cudaGetDeviceProperties(&prop0, 0);
    cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return latency_sweep((EncodeFn)fn0, prop0.clockRate * 1e3);
  }

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const double clk = prop.clockRate * 1e3;  // Hz (max SM clock)
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int N = 1280;
  uint8_t* flush;
  cudaMalloc(&flush, 512u << 20);
  printf("%-6s %4s %5s %6s %5s %4s %5s | %9s %9s %10s %9s\n", "mode", "bn", "arows", "K", "ctas", "stg", "cold", "us",
         "B/clk/SM", "TB/s total", "KB/CTA");
  for (int K : {5120, 11520}) {
    const int KB = K / 64;
    void *w, *a;
    const int NMAT = 12;                                  // 12 x 29.5 MB = 354 MB at K = 11520: far beyond the 126 MB L2
    cudaMalloc(&w, (size_t)NMAT * N * K * 2);
    cudaMalloc(&a, (size_t)128 * K * 2);
    cudaMemset(w, 1, (size_t)NMAT * N * K * 2);
    cudaMemset(a, 1, (size_t)128 * K * 2);
    for (int bn : {80, 128, 256}) {
      const int ntiles = N / bn;
      for (int splits : {4, 8, 9}) {
        if (KB % splits) continue;
        const int ctas = ntiles * splits;
        if (ctas > 148 || ctas < 60) continue;
        for (int arows : {0, 64}) {
          for (int mode = 0; mode < 3; ++mode) {
            for (int cold = 0; cold < 2; ++cold) {
              Params p;
              p.clk = nullptr;
              p.mode = mode;
              p.bn = bn;
              p.arows = arows;
              p.iters = KB / splits;
              p.kblocks_total = KB;
              p.ntiles = ntiles;
              p.wbase = (const uint8_t*)w;
              p.passes = 24;
              p.nmat = cold ? NMAT : 1;
              p.mat_bytes = (size_t)N * K * 2;
              const int stage_bytes = bn * 128 + arows * 128;
              int stages = (200 * 1024) / stage_bytes;
              if (stages > 12) stages = 12;
              if (stages > p.iters) stages = p.iters;
              p.stages = stages;
              if (mode == 0) make2d(enc, &p.tmW, w, K, (uint64_t)N * NMAT, (uint64_t)K * 2, 64, bn);
              else make2d(enc, &p.tmW, w, 64, (uint64_t)N * KB * NMAT, 128, 64, bn);
              make2d(enc, &p.tmA, a, K, 128, (uint64_t)K * 2, 64, arows ? arows : 64);
              const size_t smem = (size_t)stages * stage_bytes + stages * 16 + 2048;
              cudaEvent_t e0, e1;
              cudaEventCreate(&e0);
              cudaEventCreate(&e1);
              float best = 1e30f;
              for (int rep = 0; rep < 3; ++rep) {
                ingest_kernel<<<ctas, 64, smem>>>(p);
                cudaEventRecord(e0);
                ingest_kernel<<<ctas, 64, smem>>>(p);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
              }
              cudaError_t err = cudaGetLastError();
              if (err != cudaSuccess) {
                printf("error: %s\n", cudaGetErrorString(err));
                return 1;
              }
              const double us = best * 1e3;
              const double bytes_cta = (double)p.iters * p.passes * stage_bytes;
              const double t_net = (us - 4.0) * 1e-6;  // ~4 us of launch + first round trip are not streaming time
              printf("%-6d %4d %5d %6d %5d %4d %5d | %9.2f %9.1f %10.2f %9.1f\n", mode, bn, arows, K, ctas, stages, cold, us,
                     bytes_cta / (t_net > 0 ? t_net : 1e-9) / clk, bytes_cta * ctas / (us * 1e-6) / 1e12,
                     bytes_cta / 1024.0);
              cudaEventDestroy(e0);
              cudaEventDestroy(e1);
            }
          }
        }
      }
    }
    cudaFree(w);
    cudaFree(a);
  }
  return 0;
}
