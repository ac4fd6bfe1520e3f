// Microbenchmark: how fast can one SM pull GEMM operand tiles out of L2, and through which path?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2_ingest_bench tools/l2_ingest_bench.cu -lcuda
//   tools/l2_ingest_bench
//
// The C2 Linear GEMMs (2048 x 512 x 512, 128 CTAs of 128 x 64 tiles) spend their main loop waiting for operand bytes
// (~41 B/clk/SM).  This program separates the candidates: is that a per-SM limit of the TMA unit, a
// per-SM limit of the L2 -> SM port, or the chip-wide L2 output limit?  No MMA, no epilogue: CTAs only ingest.
//   path tma   : cp.async.bulk.tensor.2d, 128-byte swizzle, box {64 k, rows} (what the GEMM does)
//   path bulk  : cp.async.bulk (1-D) of contiguous pre-tiled 16 KB / 8 KB chunks (no tensor map)
//   path ldgsts: cp.async.cg 16 B per thread, 128 producer threads
//   path mc    : tma with the A tile multicast across a cluster of 8 (each CTA loads 1/8 of it)
// access gemm  : CTA (x = id % 8, y = id / 8) reads A slab y and W slab x (tiles shared by 8 / 16 CTAs, like the GEMM)
// access priv  : every CTA reads its own bytes
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CHECK(x)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); }            \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_load(uint64_t* bar, void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int KB_PER_GEMM = 8;          // K = 512 in blocks of 64
constexpr int A_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int B_BYTES = 64 * 64 * 2;    //  8 KB
constexpr int STAGE = A_BYTES + B_BYTES;

enum { PATH_TMA = 0, PATH_BULK = 1, PATH_LDGSTS = 2, PATH_MC = 3, PATH_TMA_A_ONLY = 4 };

struct Args {
  const uint8_t* a;     // A: [slabs][8 kb][16 KB] pre-tiled (bulk / ldgsts), or row-major [rows, 512] bf16 (tma)
  const uint8_t* w;     // W: [8 slabs][8 kb][8 KB] pre-tiled, or row-major [512, 512]
  int iters;            // k-blocks to ingest (a multiple of 8)
  int priv;             // 1: private bytes per CTA
  int stages;
  long long* cycles;    // per CTA
};

template <int PATH>
__global__ void __launch_bounds__(160, 1)
ingest_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tw, Args g) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full[8];
  const int S = g.stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) mbar_init(&full[i], PATH == PATH_LDGSTS ? 128 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (PATH == PATH_MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const int id = blockIdx.x;
  const int ay = g.priv ? id : (id / 8) % 16, wx = g.priv ? id % 8 : id % 8;
  long long t0 = clock64();
  if (PATH == PATH_LDGSTS) {
    if (threadIdx.x >= 32) {
      const int t = threadIdx.x - 32;
      for (int i = 0; i < g.iters; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait(&full[s], ((i / S) - 1) & 1);
        const int kb = i % KB_PER_GEMM;
        const uint8_t* sa = g.a + (static_cast<size_t>(ay) * KB_PER_GEMM + kb) * A_BYTES;
        const uint8_t* sw = g.w + (static_cast<size_t>(wx) * KB_PER_GEMM + kb) * B_BYTES;
        uint8_t* d = smem + s * STAGE;
#pragma unroll
        for (int j = 0; j < A_BYTES / (128 * 16); ++j) {
          const int off = (j * 128 + t) * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + off)), "l"(sa + off) : "memory");
        }
#pragma unroll
        for (int j = 0; j < B_BYTES / (128 * 16); ++j) {
          const int off = (j * 128 + t) * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + A_BYTES + off)), "l"(sw + off) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      }
      for (int i = max(0, g.iters - S); i < g.iters; ++i) mbar_wait(&full[i % S], (i / S) & 1);
    }
  } else if (threadIdx.x == 0) {
    uint32_t rank = 0;
    if (PATH == PATH_MC) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = 0; i < g.iters; ++i) {
      const int s = i % S;
      if (i >= S) mbar_wait(&full[s], ((i / S) - 1) & 1);
      const int kb = i % KB_PER_GEMM;
      uint8_t* d = smem + s * STAGE;
      if (PATH == PATH_TMA) {
        mbar_expect_tx(&full[s], STAGE);
        tma_load_2d(&ta, &full[s], d, kb * 64, ay * 128);
        tma_load_2d(&tw, &full[s], d + A_BYTES, kb * 64, wx * 64);
      } else if (PATH == PATH_TMA_A_ONLY) {
        mbar_expect_tx(&full[s], A_BYTES);
        tma_load_2d(&ta, &full[s], d, kb * 64, ay * 128);
      } else if (PATH == PATH_BULK) {
        mbar_expect_tx(&full[s], STAGE);
        bulk_load(&full[s], d, g.a + (static_cast<size_t>(ay) * KB_PER_GEMM + kb) * A_BYTES, A_BYTES);
        bulk_load(&full[s], d + A_BYTES, g.w + (static_cast<size_t>(wx) * KB_PER_GEMM + kb) * B_BYTES, B_BYTES);
      } else if (PATH == PATH_MC) {
        // every CTA receives the whole A tile (8 multicast slices of 16 rows) + its own W tile
        mbar_expect_tx(&full[s], STAGE);
        tma_load_2d_mc(&ta, &full[s], d + rank * (16 * 128), kb * 64, ay * 128 + rank * 16, 0xFF);
        tma_load_2d(&tw, &full[s], d + A_BYTES, kb * 64, wx * 64);
      }
    }
    for (int i = max(0, g.iters - S); i < g.iters; ++i) mbar_wait(&full[i % S], (i / S) & 1);
  }
  __syncthreads();
  if (PATH == PATH_MC) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) g.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return reinterpret_cast<EncodeTiledFn>(p);
}
static CUtensorMap make_map(const void* base, int rows, int cols, int box_cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r)); exit(1); }
  return m;
}

template <int PATH>
static void run(const char* name, int ctas, int per_sm, int priv, int stages, int iters, const uint8_t* a, const uint8_t* w,
                long long* cyc_d, double mhz) {
  const int a_rows = 148 * 2 * 128;
  CUtensorMap ta = make_map(a, a_rows, 512, 64, PATH == PATH_MC ? 16 : 128);
  CUtensorMap tw = make_map(w, 512, 512, 64, 64);
  // per_sm == 1: ask for more than half of the shared memory so that only one CTA fits
  const int smem = per_sm == 1 ? 120 * 1024 : stages * STAGE + 2048;
  if (stages * STAGE + 2048 > smem) { printf("%s: stages do not fit\n", name); return; }
  auto kern = ingest_kernel<PATH>;
  CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Args g{a, w, iters, priv, stages, cyc_d};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = PATH == PATH_MC ? 1 : 0;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CHECK(cudaEventRecord(e0));
    CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tw, g));
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
  }
  float ms = 0.f;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(ctas);
  CHECK(cudaMemcpy(cyc.data(), cyc_d, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  std::sort(cyc.begin(), cyc.end());
  const double per_iter = PATH == PATH_TMA_A_ONLY ? A_BYTES : STAGE;
  const double bytes = per_iter * iters;
  const double med = static_cast<double>(cyc[ctas / 2]), mx = static_cast<double>(cyc.back());
  printf("%-10s ctas=%3d /SM=%d %s stages=%d : median %6.1f B/clk/CTA, slowest %6.1f, aggregate %7.0f B/clk = %6.2f TB/s (event %.1f us)\n",
         name, ctas, per_sm, priv ? "priv" : "gemm", stages, bytes / med, bytes / mx, bytes * ctas / mx,
         bytes * ctas / mx * mhz * 1e6 / 1e12, ms * 1e3);
}

int main(int argc, char** argv) {
  int dev = 0;
  CHECK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, dev));
  int khz = 0;
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const double mhz = khz / 1e3;
  printf("%s, %d SMs, %.0f MHz (max)\n", p.name, p.multiProcessorCount, mhz);
  const size_t a_bytes = static_cast<size_t>(148) * 2 * 128 * 512 * 2;   // 38 MB: private slabs for up to 296 CTAs
  const size_t w_bytes = 512 * 512 * 2;
  uint8_t *a, *w;
  long long* cyc;
  CHECK(cudaMalloc(&a, a_bytes)); CHECK(cudaMalloc(&w, w_bytes)); CHECK(cudaMalloc(&cyc, sizeof(long long) * 512));
  CHECK(cudaMemset(a, 1, a_bytes)); CHECK(cudaMemset(w, 1, w_bytes));
  const int iters = 8 * 32;   // 32 GEMMs' worth of k-blocks per CTA
  const int grids[] = {8, 16, 32, 64, 128, 144};
  for (int gsz : grids) {
    run<PATH_TMA>("tma", gsz, 1, 0, 4, iters, a, w, cyc, mhz);
    run<PATH_TMA>("tma", gsz, 1, 1, 4, iters, a, w, cyc, mhz);
  }
  run<PATH_TMA>("tma", 128, 1, 0, 2, iters, a, w, cyc, mhz);
  run<PATH_TMA>("tma", 128, 1, 0, 3, iters, a, w, cyc, mhz);
  run<PATH_TMA>("tma", 256, 2, 0, 4, iters, a, w, cyc, mhz);
  run<PATH_TMA>("tma", 256, 2, 1, 4, iters, a, w, cyc, mhz);
  run<PATH_TMA_A_ONLY>("tma-A", 128, 1, 0, 4, iters, a, w, cyc, mhz);
  run<PATH_TMA_A_ONLY>("tma-A", 16, 1, 0, 4, iters, a, w, cyc, mhz);
  for (int gsz : grids) {
    run<PATH_BULK>("bulk", gsz, 1, 0, 4, iters, a, w, cyc, mhz);
    run<PATH_BULK>("bulk", gsz, 1, 1, 4, iters, a, w, cyc, mhz);
  }
  run<PATH_BULK>("bulk", 256, 2, 0, 4, iters, a, w, cyc, mhz);
  for (int gsz : grids) {
    run<PATH_LDGSTS>("ldgsts", gsz, 1, 0, 4, iters, a, w, cyc, mhz);
    run<PATH_LDGSTS>("ldgsts", gsz, 1, 1, 4, iters, a, w, cyc, mhz);
  }
  run<PATH_LDGSTS>("ldgsts", 256, 2, 0, 4, iters, a, w, cyc, mhz);
  for (int gsz : grids) run<PATH_MC>("mc8", gsz, 1, 0, 4, iters, a, w, cyc, mhz);
  run<PATH_MC>("mc8", 256, 2, 0, 4, iters, a, w, cyc, mhz);
  return 0;
}
