// Row-wise kernels of the LayerNorm/RMSNorm-MLP: minibatch gather + concat (K2/K5), and the
// norm + swish block of FCNN forward and backward (K3/K4 epilogue math).
//   reference: FCNN.__call__ src/networks/jax_models.py:109-124 (nnx.RMSNorm eps 1e-6, swish),
//              normed_activation_layer src/networks/torch_models.py:71-79 (nn.RMSNorm eps=finfo.eps),
//              LayerNorm variant src/reppo_mj_playground.py:63-74 (flax "fast variance", eps 1e-6).
// One warp owns one row; lanes stride the feature dimension (coalesced 128 B per access).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace reppo {

// out[r, :] = [ a[idx[r], 0:da] | b[idx2 ? r : idx[r], 0:db] ]   (idx may be NULL = identity)
__global__ void gather_concat_kernel(const float* __restrict__ a, int da, const float* __restrict__ b, int db,
                                     const int* __restrict__ idx, int b_is_local, int rows, float* __restrict__ out,
                                     int out_ld, __nv_bfloat16* __restrict__ out_h, int out_ldh) {
  const int w = da + db;
  const long long total = static_cast<long long>(rows) * w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / w), c = static_cast<int>(i % w);
    const long long src = idx ? idx[r] : r;
    float v;
    if (c < da) v = a[src * da + c];
    else v = b[(b_is_local ? static_cast<long long>(r) : src) * db + (c - da)];
    out[static_cast<long long>(r) * out_ld + c] = v;
    if (out_h != nullptr) out_h[static_cast<long long>(r) * out_ldh + c] = __float2bfloat16_rn(v);
  }
}

cudaError_t launch_gather_concat(const float* a, int da, const float* b, int db, const int* idx, int b_is_local,
                                 int rows, float* out, int out_ld, void* out_h, int out_ldh, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * (da + db);
  if (total <= 0) return cudaSuccess;
  const int block = 256;
  const long long want = (total + block - 1) / block;
  const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  gather_concat_kernel<<<grid, block, 0, st>>>(a, da, b, db, idx, b_is_local, rows, out, out_ld,
                                               static_cast<__nv_bfloat16*>(out_h), out_ldh);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// forward: x = swish( norm(z) )      norm: none | rms (z * rsqrt(mean z^2 + eps) * g) | layer
// saves rstd (and mean for layer norm) per row for the backward.
// ------------------------------------------------------------------------------------------------
template <int NORM>
__global__ void __launch_bounds__(256)
norm_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int rows, int H, float eps, float* __restrict__ x, float* __restrict__ rstd_out,
                    float* __restrict__ mean_out, __nv_bfloat16* __restrict__ x_h, int ldh) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* zr = z + static_cast<size_t>(row) * H;
  float* xr = x + static_cast<size_t>(row) * H;
  float rstd = 1.f, mean = 0.f;
  if (NORM != NORM_NONE) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < H; c += 32) { const float v = zr[c]; s1 += v; s2 += v * v; }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (NORM == NORM_RMS) {
      rstd = rsqrtf(s2 / H + eps);
    } else {
      mean = s1 / H;
      const float var = s2 / H - mean * mean;   // flax use_fast_variance
      rstd = rsqrtf(var + eps);
    }
    if (lane == 0) { rstd_out[row] = rstd; if (NORM == NORM_LAYER) mean_out[row] = mean; }
  }
  for (int c = lane; c < H; c += 32) {
    float y = zr[c];
    if (NORM == NORM_RMS) y = y * rstd * gamma[c];
    else if (NORM == NORM_LAYER) y = (y - mean) * rstd * gamma[c] + beta[c];
    const float o = swishf_(y);
    xr[c] = o;
    if (x_h != nullptr) x_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
  }
}

cudaError_t launch_norm_act_fwd(int norm, const float* z, const float* gamma, const float* beta, int rows, int H,
                                float eps, float* x, float* rstd, float* mean, void* x_h, int ldh, cudaStream_t st) {
  __nv_bfloat16* xh = static_cast<__nv_bfloat16*>(x_h);
  if (rows <= 0) return cudaSuccess;
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  if (norm == NORM_NONE) norm_act_fwd_kernel<NORM_NONE><<<grid, wpb * 32, 0, st>>>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh);
  else if (norm == NORM_RMS) norm_act_fwd_kernel<NORM_RMS><<<grid, wpb * 32, 0, st>>>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh);
  else norm_act_fwd_kernel<NORM_LAYER><<<grid, wpb * 32, 0, st>>>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward: given dx = dL/d swish(y), y = norm(z): dz, and dgamma / dbeta accumulated with atomics
// (the gradient arena is zeroed at the start of a step).
//   rms  : zh = z*rstd;           dzh = dy*g;  dz = rstd * (dzh - zh * mean(dzh*zh))
//   layer: zh = (z-mean)*rstd;    dz = rstd * (dzh - mean(dzh) - zh * mean(dzh*zh))
// Each block walks ROWS_PER_BLOCK rows, 8 warps; per-column partial sums of dgamma/dbeta stay in
// registers of the lane that owns the column, are combined across the block's warps in shared
// memory and leave as one atomicAdd per column per block.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxColsPerLane = 64;  // H <= 2048

template <int NORM>
__global__ void __launch_bounds__(256)
norm_act_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ z, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ rstd_in,
                    const float* __restrict__ mean_in, int rows, int H, int rows_per_block, float* __restrict__ dz,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dz_h, int ldh) {
  extern __shared__ float sm[];  // [2][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  const int cpl = (H + 31) / 32;

  if (NORM != NORM_NONE) {
    for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
  }
  float pg[kMaxColsPerLane / 4], pb[kMaxColsPerLane / 4];  // flushed to smem every 16 columns-per-lane chunk
  // To keep register use bounded for large H the column range is processed in chunks of 16 per lane.
  for (int cbase = 0; cbase < cpl; cbase += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { pg[j] = 0.f; pb[j] = 0.f; }
    for (int row = r0 + warp; row < r1; row += nwarp) {
      const float* zr = z + static_cast<size_t>(row) * H;
      const float* dxr = dx + static_cast<size_t>(row) * H;
      float* dzr = dz + static_cast<size_t>(row) * H;
      if (NORM == NORM_NONE) {
        if (cbase == 0)
          for (int c = lane; c < H; c += 32) {
            const float o = dxr[c] * swish_grad_(zr[c]);
            dzr[c] = o;
            if (dz_h != nullptr) dz_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
          }
        continue;
      }
      const float rstd = rstd_in[row];
      const float mean = (NORM == NORM_LAYER) ? mean_in[row] : 0.f;
      // pass 1 (whole row): m1 = mean(dzh), m2 = mean(dzh*zh)
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < H; c += 32) {
        const float zh = (zr[c] - mean) * rstd;
        const float y = (NORM == NORM_LAYER) ? zh * gamma[c] + beta[c] : zh * gamma[c];
        const float dy = dxr[c] * swish_grad_(y);
        const float dzh = dy * gamma[c];
        s1 += dzh; s2 += dzh * zh;
      }
      s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
      if (NORM == NORM_RMS) s1 = 0.f;
      // pass 2 (this column chunk): dz, dgamma, dbeta
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = lane + 32 * (cbase + j);
        if (c < H) {
          const float zh = (zr[c] - mean) * rstd;
          const float y = (NORM == NORM_LAYER) ? zh * gamma[c] + beta[c] : zh * gamma[c];
          const float dy = dxr[c] * swish_grad_(y);
          const float dzh = dy * gamma[c];
          const float o = rstd * (dzh - s1 - zh * s2);
          dzr[c] = o;
          if (dz_h != nullptr) dz_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
          pg[j] += dy * zh;
          pb[j] += dy;
        }
      }
    }
    if (NORM != NORM_NONE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = lane + 32 * (cbase + j);
        if (c < H) {
          atomicAdd(&sm[c], pg[j]);
          if (NORM == NORM_LAYER) atomicAdd(&sm[H + c], pb[j]);
        }
      }
    }
  }
  if (NORM != NORM_NONE && dgamma != nullptr) {
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      atomicAdd(dgamma + c, sm[c]);
      if (NORM == NORM_LAYER) atomicAdd(dbeta + c, sm[H + c]);
    }
  }
}

cudaError_t launch_norm_act_bwd(int norm, const float* dx, const float* z, const float* gamma, const float* beta,
                                const float* rstd, const float* mean, int rows, int H, float* dz, float* dgamma,
                                float* dbeta, void* dz_h, int ldh, cudaStream_t st) {
  __nv_bfloat16* dzh = static_cast<__nv_bfloat16*>(dz_h);
  if (rows <= 0) return cudaSuccess;
  if (H > 32 * kMaxColsPerLane) return cudaErrorInvalidValue;
  // enough blocks to fill the machine, few enough that the per-block atomics stay cheap
  int rows_per_block = 16;
  while ((rows + rows_per_block - 1) / rows_per_block > 148 * 4 && rows_per_block < 128) rows_per_block *= 2;
  const int grid = (rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = (norm == NORM_NONE) ? 0 : 2 * static_cast<size_t>(H) * sizeof(float);
  if (norm == NORM_NONE) norm_act_bwd_kernel<NORM_NONE><<<grid, 256, smem, st>>>(dx, z, gamma, beta, rstd, mean, rows, H, rows_per_block, dz, dgamma, dbeta, dzh, ldh);
  else if (norm == NORM_RMS) norm_act_bwd_kernel<NORM_RMS><<<grid, 256, smem, st>>>(dx, z, gamma, beta, rstd, mean, rows, H, rows_per_block, dz, dgamma, dbeta, dzh, ldh);
  else norm_act_bwd_kernel<NORM_LAYER><<<grid, 256, smem, st>>>(dx, z, gamma, beta, rstd, mean, rows, H, rows_per_block, dz, dgamma, dbeta, dzh, ldh);
  return cudaGetLastError();
}

}  // namespace reppo
