// Microbenchmark: how long until a GEMM epilogue's output tile is VISIBLE in L2, and through which store path?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2_store_bench tools/l2_store_bench.cu -lcuda
//
// tools/gemm_stamps.py shows that a 128 x 64 fp32 tile (+ its bf16 twin) takes ~1 700 cycles to issue with st.global.v4
// and another ~2 000 until __threadfence() returns (128 CTAs storing at once): 13 B/clk/SM.  Candidates:
//   st    : the epilogue's pattern -- 256 threads, float4 from a padded staging buffer, 4 rows x 128 B per warp store
//   strow : one warp writes whole 256 B rows (float2 per lane ... two rows per instruction)
//   tma   : cp.async.bulk.tensor.2d smem -> global (one elected thread), completion = cp.async.bulk.wait_group 0
// Each CTA owns tile (x = id % 8, y = id / 8) of a [2048, 512] fp32 matrix and of its bf16 twin.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CHECK(x)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); }            \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

enum { MODE_ST = 0, MODE_TMA = 1, MODE_ST_NOBF = 2, MODE_TMA_NOBF = 3 };

struct Out { long long issue, visible; };

template <int MODE>
__global__ void __launch_bounds__(256, 1)
store_kernel(const __grid_constant__ CUtensorMap tc, const __grid_constant__ CUtensorMap th, float* C, __nv_bfloat16* Hh, int ldc,
             int ldh, int iters, Out* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* tile = reinterpret_cast<float*>(smem);                          // [128][64] dense (tma) or [128][68] padded (st)
  __nv_bfloat16* tile_h = reinterpret_cast<__nv_bfloat16*>(smem + 40 * 1024);   // [128][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 68; i += 256) tile[i] = static_cast<float>(i);
  for (int i = tid; i < 128 * 64; i += 256) tile_h[i] = __float2bfloat16(static_cast<float>(i));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int n0 = (blockIdx.x % 8) * 64, m0 = (blockIdx.x / 8) * 128;
  long long t_issue = 0, t_vis = 0;
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == MODE_ST || MODE == MODE_ST_NOBF) {
      // warp w: rows 32*(w%4)..+32, columns 32*(w/4)..+32; 8 lanes per row, 4 rows per instruction
      const int lane_base = (warp & 3) * 32, cbase = (warp >> 2) * 32, cl = (lane % 8) * 4;
#pragma unroll 4
      for (int r = lane / 8; r < 32; r += 4) {
        const int row = m0 + lane_base + r, col = n0 + cbase + cl;
        const float4 x = *reinterpret_cast<const float4*>(tile + (lane_base + r) * 68 + cbase + cl);
        *reinterpret_cast<float4*>(C + static_cast<size_t>(row) * ldc + col) = x;
        if (MODE == MODE_ST) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
          uint2 u;
          u.x = *reinterpret_cast<const uint32_t*>(&lo);
          u.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(Hh + static_cast<size_t>(row) * ldh + col) = u;
        }
      }
      const long long t1 = clock64();
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        t_issue += t1 - t0;
        t_vis += clock64() - t0;
      }
    } else {
      if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tc)), "r"(smem_u32(tile)), "r"(n0), "r"(m0) : "memory");
        if (MODE == MODE_TMA)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&th)), "r"(smem_u32(tile_h)), "r"(n0), "r"(m0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        const long long t1 = clock64();
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        t_issue += t1 - t0;
        t_vis += clock64() - t0;
      }
    }
  }
  if (tid == 0) { out[blockIdx.x].issue = t_issue / iters; out[blockIdx.x].visible = t_vis / iters; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(void* base, int rows, int cols, int elem, int box_cols, int box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * elem};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(p)(&m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                                                  base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r)); exit(1); }
  return m;
}

template <int MODE>
static void run(const char* name, int ctas, float* C, __nv_bfloat16* H, Out* out_d) {
  CUtensorMap tc = make_map(C, 2048 * 2, 512, 4, 64, 128);
  CUtensorMap th = make_map(H, 2048 * 2, 512, 2, 64, 128);
  auto kern = store_kernel<MODE>;
  const int smem = 120 * 1024;
  CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<ctas, 256, smem>>>(tc, th, C, H, 512, 512, 16, out_d);
    CHECK(cudaDeviceSynchronize());
  }
  std::vector<Out> o(ctas);
  CHECK(cudaMemcpy(o.data(), out_d, sizeof(Out) * ctas, cudaMemcpyDeviceToHost));
  std::vector<long long> a, b;
  for (auto& x : o) { a.push_back(x.issue); b.push_back(x.visible); }
  std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
  const double bytes = (MODE == MODE_ST || MODE == MODE_TMA) ? 48.0 * 1024 : 32.0 * 1024;
  printf("%-9s ctas=%3d : issue median %6lld clk, visible median %6lld max %6lld clk  -> %5.1f B/clk/SM, %6.0f B/clk chip\n", name, ctas,
         a[ctas / 2], b[ctas / 2], b.back(), bytes / b[ctas / 2], bytes * ctas / b.back());
}

int main() {
  CHECK(cudaSetDevice(0));
  float* C; __nv_bfloat16* H; Out* out;
  CHECK(cudaMalloc(&C, sizeof(float) * 4096 * 512)); CHECK(cudaMalloc(&H, 2 * 4096 * 512)); CHECK(cudaMalloc(&out, sizeof(Out) * 512));
  const int grids[] = {8, 32, 64, 128, 256};
  for (int g : grids) {
    run<MODE_ST>("st+bf16", g, C, H, out);
    run<MODE_ST_NOBF>("st", g, C, H, out);
    run<MODE_TMA>("tma+bf16", g, C, H, out);
    run<MODE_TMA_NOBF>("tma", g, C, H, out);
  }
  return 0;
}
