// fp32 -> bf16 operand staging for the tcgen05 GEMMs (REPPO_GEMM_BF16 mode).
//  * pack_weights_kernel: the fp32 master weights W[out,in] of one network re-rounded into their bf16 shadows in
//    one launch (entry points that receive parameters from the caller; inside an update the Adam kernel refreshes
//    the shadows itself).  The same shadow serves the forward GEMM (K-major B) and the dgrad GEMM (MN-major B).
//  * cast_bf16_kernel: plain row-padded cast (stand-alone reppo_linear calls).
//  * pad_rows_f32_kernel: row-padded fp32 copy of a ragged tf32 operand (REPPO_GEMM_TF32 mode).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

__global__ void __launch_bounds__(256)
pack_weights_kernel(const PackTable t) {
  REPPO_PDL_PROLOGUE();
  // blockIdx.y = tensor, blockIdx.x grid-strides over its elements (rows of the shadow are padded to ld_b)
  const PackEntry e = t.e[blockIdx.y];
  __nv_bfloat16* wb = static_cast<__nv_bfloat16*>(e.wb);
  float* wf = static_cast<float*>(e.wb);
  const int total = e.out * e.in;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / e.in, c = i - r * e.in;
    if (t.fp32) wf[static_cast<size_t>(r) * e.ld_b + c] = e.w[i];   // aligned, row-padded fp32 copy (tf32 operands)
    else wb[static_cast<size_t>(r) * e.ld_b + c] = __float2bfloat16_rn(e.w[i]);
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

// tf32 mode: an fp32 operand whose leading dimension TMA cannot address (not a multiple of four floats) re-laid with a padded one
__global__ void pad_rows_f32_kernel(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ dst, int ld_dst) {
  REPPO_PDL_PROLOGUE();
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    dst[static_cast<size_t>(r) * ld_dst + c] = src[static_cast<size_t>(r) * ld_src + c];
  }
}

cudaError_t launch_pad_rows_f32(const float* src, int rows, int cols, int ld_src, float* dst, int ld_dst, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * cols;
  if (total <= 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  launch_k((pad_rows_f32_kernel), dim3(static_cast<int>(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, st, src, rows, cols, ld_src, dst, ld_dst);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* src, int rows, int cols, int ld_src, void* dst, int ld_dst, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * cols;
  if (total <= 0) return cudaSuccess;
  const long long want = (total + 255) / 256;
  launch_k((cast_bf16_kernel), dim3(static_cast<int>(want < 148 * 8 ? want : 148 * 8)), dim3(256), 0, st, src, rows, cols, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst);
  return cudaGetLastError();
}

}  // namespace reppo
