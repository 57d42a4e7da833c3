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
  if (threadIdx.x == 0) {
    for (int g = 0; g < total; ++g) {
      const int it = g % p.iters, pass = g / p.iters;
      const int mat = pass % p.nmat;                      // a different weight matrix per pass when nmat > 1 (cold)
      const int s = g % p.stages;
      mbar_wait(&empty[s], ((g / p.stages) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], wb + ab);
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
    }
  } else if (threadIdx.x == 32) {
    for (int g = 0; g < total; ++g) {
      const int s = g % p.stages;
      mbar_wait(&full[s], (g / p.stages) & 1);
      mbar_arrive(&empty[s]);
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

int main() {
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
