// fp32 -> bf16 operand staging for the tcgen05 GEMMs (REPPO_GEMM_BF16 mode).
//  * pack_weights_kernel: after every Adam step the fp32 master weights W[out,in] of one network are
//    re-rounded into two bf16 shadows in one launch: W (K-major B operand of the forward GEMM) and
//    W^T (K-major B operand of the dgrad GEMM).  32x32 smem tiles make both writes coalesced.
//  * cast_bf16_kernel: plain row-padded cast (stand-alone reppo_linear calls).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

__global__ void __launch_bounds__(256)
pack_weights_kernel(const PackTable t) {
  REPPO_PDL_PROLOGUE();
  __shared__ float tile[32][33];
  // blockIdx.y = tensor, blockIdx.x = tile index inside the tensor (grid-strided)
  const PackEntry e = t.e[blockIdx.y];
  const int tiles_c = (e.in + 31) / 32, tiles_r = (e.out + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  __nv_bfloat16* wb = static_cast<__nv_bfloat16*>(e.wb);
  __nv_bfloat16* wtb = static_cast<__nv_bfloat16*>(e.wtb);
  for (int tl = blockIdx.x; tl < tiles_c * tiles_r; tl += gridDim.x) {
    const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const int r = r0 + ty + i, c = c0 + tx;
      float v = 0.f;
      if (r < e.out && c < e.in) {
        v = e.w[static_cast<size_t>(r) * e.in + c];
        if (wb != nullptr) wb[static_cast<size_t>(r) * e.ld_b + c] = __float2bfloat16_rn(v);
      }
      tile[ty + i][tx] = v;
    }
    __syncthreads();
    if (wtb != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;  // transposed: row index of W^T is the column of W
        if (c < e.in && r < e.out) wtb[static_cast<size_t>(c) * e.ld_t + r] = __float2bfloat16_rn(tile[tx][ty + i]);
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_pack_weights(const PackTable& t, cudaStream_t st) {
  if (t.n <= 0) return cudaSuccess;
  dim3 grid(64, t.n);
  launch_k((pack_weights_kernel), dim3(grid), dim3(256), 0, st, t);
  return cudaGetLastError();
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld_src, __nv_bfloat16* __restrict__ dst,
                                 int ld_dst) {
  REPPO_PDL_PROLOGUE();
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    dst[static_cast<size_t>(r) * ld_dst + c] = __float2bfloat16_rn(src[static_cast<size_t>(r) * ld_src + c]);
  }
}

cudaError_t launch_cast_bf16(const float* src, int rows, int cols, int ld_src, void* dst, int ld_dst, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * cols;
  if (total <= 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  launch_k((cast_bf16_kernel), dim3(static_cast<int>(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, st, src, rows, cols, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst);
  return cudaGetLastError();
}

}  // namespace reppo
