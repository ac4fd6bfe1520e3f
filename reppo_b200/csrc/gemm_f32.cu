// fp32 FFMA GEMM family (REPPO_GEMM_FP32 mode): the reference-precision path
// (jax_default_matmul_precision="highest", src/jaxrl/__init__.py:3).  Serves the Linear layers of
// FCNN (src/networks/jax_models.py:109-124, torch_models.py:71-142):
//   forward  Y[M,N]  = X[M,K] W[N,K]^T + b      (ta=0, tb=1)
//   dgrad    dX[M,K] = dY[M,N] W[N,K]           (ta=0, tb=0)
//   wgrad    dW[N,K] = dY[M,N]^T X[M,K]         (ta=1, tb=0, split over M with atomics)
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.  The tensor-core path (gemm_bf16_tc.cu)
// replaces it for the H x H layers in REPPO_GEMM_BF16 mode; this one also handles the ragged
// shapes (K = obs+act, N = num_bins / 2*act).
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

constexpr int BM = 64, BN = 64, BK = 16;

// A(m,k): TA=0 -> A[m*lda + k]; TA=1 -> A[k*lda + m].   B(k,n): TB=0 -> B[k*ldb + n]; TB=1 -> B[n*ldb + k].
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                const float* __restrict__ bias, int M, int N, int K, int lda, int ldb, int ldc, int k_chunk, int mode) {
  REPPO_PDL_PROLOGUE();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kt = k_begin; kt < k_end; kt += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lin = tid + i * 256;
      int m, k;
      if (TA) { m = lin & (BM - 1); k = lin >> 6; } else { k = lin & (BK - 1); m = lin >> 4; }
      const int gm = m0 + m, gk = kt + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[static_cast<size_t>(gk) * lda + gm] : A[static_cast<size_t>(gm) * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lin = tid + i * 256;
      int n, k;
      if (TB) { k = lin & (BK - 1); n = lin >> 4; } else { n = lin & (BN - 1); k = lin >> 6; }
      const int gn = n0 + n, gk = kt + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? B[static_cast<size_t>(gn) * ldb + gk] : B[static_cast<size_t>(gk) * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      float* c = C + static_cast<size_t>(gm) * ldc + gn;
      if (mode == GEMM_STORE) *c = v;
      else if (mode == GEMM_ACCUM) *c += v;
      else atomicAdd(c, v);
    }
  }
}

cudaError_t launch_gemm_f32(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                            float* C, int ldc, const float* bias, int mode, int split_k, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if (split_k < 1) split_k = 1;
  int k_chunk = (K + split_k - 1) / split_k;
  k_chunk = ((k_chunk + BK - 1) / BK) * BK;
  split_k = (K + k_chunk - 1) / k_chunk;
  if (split_k > 1) mode = GEMM_ATOMIC;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split_k), block(256);
  if (!ta && tb) launch_k((gemm_f32_kernel<false, true>), dim3(grid), dim3(block), 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc, k_chunk, mode);
  else if (!ta && !tb) launch_k((gemm_f32_kernel<false, false>), dim3(grid), dim3(block), 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc, k_chunk, mode);
  else if (ta && !tb) launch_k((gemm_f32_kernel<true, false>), dim3(grid), dim3(block), 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc, k_chunk, mode);
  else launch_k((gemm_f32_kernel<true, true>), dim3(grid), dim3(block), 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc, k_chunk, mode);
  return cudaGetLastError();
}

// out[n] (+)= scale * sum_m X[m, n]  (bias gradients; zero_dist gradient = logit_bias_scale * d bias)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int M, int N, int ld, int rows_per_block, float* __restrict__ out,
              float* __restrict__ out2, float scale2) {
  REPPO_PDL_PROLOGUE();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + ty; r < r1; r += 8) s += X[static_cast<size_t>(r) * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][tx];
    atomicAdd(out + n, s);
    if (out2 != nullptr) atomicAdd(out2 + n, s * scale2);
  }
}

cudaError_t launch_colsum(const float* X, int M, int N, int ld, float* out, float* out2, float scale2, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  const int rows_per_block = 256;
  dim3 grid((N + 31) / 32, (M + rows_per_block - 1) / rows_per_block);
  launch_k((colsum_kernel), dim3(grid), dim3(256), 0, st, X, M, N, ld, rows_per_block, out, out2, scale2);
  return cudaGetLastError();
}

}  // namespace reppo
