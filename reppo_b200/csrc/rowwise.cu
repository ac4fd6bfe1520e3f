// Row-wise kernels of the LayerNorm/RMSNorm-MLP: minibatch gather + concat (K2/K5), and the
// norm + swish block of FCNN forward and backward (K3/K4 epilogue math).
//   reference: FCNN.__call__ src/networks/jax_models.py:109-124 (nnx.RMSNorm eps 1e-6, swish),
//              normed_activation_layer src/networks/torch_models.py:71-79 (nn.RMSNorm eps=finfo.eps),
//              LayerNorm variant src/reppo_mj_playground.py:63-74 (flax "fast variance", eps 1e-6).
// One warp owns one row.  Fast path (H a multiple of 128, H <= 1024): the whole row lives in
// registers as VPL float4 per lane, every global access is a 512 B warp-contiguous segment and all
// loads of a row are issued before the first use (HBM/L2-latency bound otherwise: these kernels
// move 4-8 MB per launch).  Generic path: any H <= 2048, scalar strided loops.
// The backward also produces the per-column reductions that autodiff needs from the same rows:
// d gamma (/ d beta) of the norm and d bias of the Linear that feeds it (= column sum of dz), so no
// separate column-sum pass runs over dz.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

REPPO_DEVINL void store_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&lo);
  u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

// out[r, :] = [ a[idx[r], 0:da] | b[(b_is_local ? r : idx[r]), 0:db] ]   (idx may be NULL = identity)
// One warp per row, lanes over columns: no integer division, 32-bit index arithmetic inside the row (these kernels sit
// directly behind the Adam kernel on both chains; the flat-index version compiled to 30-55 KB of 64-bit div/mod code).
__global__ void __launch_bounds__(256)
gather_concat_kernel(const float* __restrict__ a, int da, const float* __restrict__ b, int db,
                     const int* __restrict__ idx, int b_is_local, int rows, float* __restrict__ out,
                     int out_ld, __nv_bfloat16* __restrict__ out_h, int out_ldh) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const size_t src = idx ? static_cast<size_t>(idx[r]) : static_cast<size_t>(r);
    const float* ar = a + src * da;
    float* orow = out + static_cast<size_t>(r) * out_ld;
    __nv_bfloat16* hrow = out_h ? out_h + static_cast<size_t>(r) * out_ldh : nullptr;
    for (int c = lane; c < da; c += 32) {
      const float v = ar[c];
      orow[c] = v;
      if (hrow != nullptr) hrow[c] = __float2bfloat16_rn(v);
    }
    if (db > 0) {
      const float* br = b + (b_is_local ? static_cast<size_t>(r) : src) * db;
      for (int c = lane; c < db; c += 32) {
        const float v = br[c];
        orow[da + c] = v;
        if (hrow != nullptr) hrow[da + c] = __float2bfloat16_rn(v);
      }
    }
  }
}

cudaError_t launch_gather_concat(const float* a, int da, const float* b, int db, const int* idx, int b_is_local,
                                 int rows, float* out, int out_ld, void* out_h, int out_ldh, cudaStream_t st) {
  if (rows <= 0 || da + db <= 0) return cudaSuccess;
  const int want = (rows + 7) / 8;
  const int grid = want < 148 * 8 ? want : 148 * 8;
  launch_k((gather_concat_kernel), dim3(grid), dim3(256), 0, st, a, da, b, db, idx, b_is_local, rows, out, out_ld, static_cast<__nv_bfloat16*>(out_h), out_ldh);
  return cudaGetLastError();
}

// two gathers through the same index in one launch: a[idx[r]] -> out_a (+ bf16 twin), b[idx[r]] -> out_b (+ bf16 twin)
__global__ void __launch_bounds__(256)
gather_pair_kernel(const float* __restrict__ a, int da, const float* __restrict__ b, int db,
                   const int* __restrict__ idx, int rows, float* __restrict__ out_a, int lda,
                   __nv_bfloat16* __restrict__ out_ah, int ldah, float* __restrict__ out_b, int ldb,
                   __nv_bfloat16* __restrict__ out_bh, int ldbh) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const size_t src = idx ? static_cast<size_t>(idx[r]) : static_cast<size_t>(r);
    const float* ar = a + src * da;
    const float* br = b + src * db;
    for (int c = lane; c < da; c += 32) {
      const float v = ar[c];
      if (out_a != nullptr) out_a[static_cast<size_t>(r) * lda + c] = v;
      if (out_ah != nullptr) out_ah[static_cast<size_t>(r) * ldah + c] = __float2bfloat16_rn(v);
    }
    for (int c = lane; c < db; c += 32) {
      const float v = br[c];
      if (out_b != nullptr) out_b[static_cast<size_t>(r) * ldb + c] = v;
      if (out_bh != nullptr) out_bh[static_cast<size_t>(r) * ldbh + c] = __float2bfloat16_rn(v);
    }
  }
}

cudaError_t launch_gather_pair(const float* a, int da, const float* b, int db, const int* idx, int rows, float* out_a, int lda,
                               void* out_ah, int ldah, float* out_b, int ldb, void* out_bh, int ldbh, cudaStream_t st) {
  if (rows <= 0 || da + db <= 0) return cudaSuccess;
  const int want = (rows + 7) / 8;
  const int grid = want < 148 * 8 ? want : 148 * 8;
  launch_k((gather_pair_kernel), dim3(grid), dim3(256), 0, st, a, da, b, db, idx, rows, out_a, lda,
           static_cast<__nv_bfloat16*>(out_ah), ldah, out_b, ldb, static_cast<__nv_bfloat16*>(out_bh), ldbh);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// forward: x = swish( norm(z) )      norm: none | rms (z * rsqrt(mean z^2 + eps) * g) | layer
// saves rstd (and mean for layer norm) per row for the backward.  x (fp32) and x_h (bf16) are both
// optional outputs.
// ------------------------------------------------------------------------------------------------
template <int NORM>
__global__ void __launch_bounds__(256)
norm_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int rows, int H, float eps, float* __restrict__ x, float* __restrict__ rstd_out,
                    float* __restrict__ mean_out, __nv_bfloat16* __restrict__ x_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* zr = z + static_cast<size_t>(row) * H;
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
    if (x != nullptr) x[static_cast<size_t>(row) * H + c] = o;
    if (x_h != nullptr) x_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
  }
}

template <int NORM, int VPL, bool FAST>
__global__ void __launch_bounds__(256)
norm_act_fwd_vec_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                        int rows, float eps, float* __restrict__ x, float* __restrict__ rstd_out,
                        float* __restrict__ mean_out, __nv_bfloat16* __restrict__ x_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  constexpr int H = 128 * VPL;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* zr = reinterpret_cast<const float4*>(z + static_cast<size_t>(row) * H);
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = zr[lane + 32 * i];
  float rstd = 1.f, mean = 0.f;
  if (NORM != NORM_NONE) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      s1 += v[i].x + v[i].y + v[i].z + v[i].w;
      s2 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (NORM == NORM_RMS) {
      rstd = rsqrtf(s2 / H + eps);
    } else {
      mean = s1 / H;
      rstd = rsqrtf(s2 / H - mean * mean + eps);
    }
    if (lane == 0) { rstd_out[row] = rstd; if (NORM == NORM_LAYER) mean_out[row] = mean; }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float y[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    if (NORM != NORM_NONE) {
      // gamma / beta live in the flat parameter arena: only 4-byte aligned
      const int cg = 4 * (lane + 32 * i);
      const float gg[4] = {gamma[cg], gamma[cg + 1], gamma[cg + 2], gamma[cg + 3]};
      float bb[4] = {0.f, 0.f, 0.f, 0.f};
      if (NORM == NORM_LAYER) { bb[0] = beta[cg]; bb[1] = beta[cg + 1]; bb[2] = beta[cg + 2]; bb[3] = beta[cg + 3]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = (NORM == NORM_RMS) ? y[j] * rstd * gg[j] : (y[j] - mean) * rstd * gg[j] + bb[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = FAST ? y[j] * __fdividef(1.f, 1.f + __expf(-fmaxf(y[j], -80.f))) : swishf_(y[j]);
    const int c = 4 * (lane + 32 * i);
    if (x != nullptr) reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H)[lane + 32 * i] = make_float4(y[0], y[1], y[2], y[3]);
    if (x_h != nullptr) store_bf16x4(x_h + static_cast<size_t>(row) * ldh + c, y[0], y[1], y[2], y[3]);
  }
}

template <int NORM>
static void launch_fwd_norm(const float* z, const float* gamma, const float* beta, int rows, int H, float eps, float* x,
                            float* rstd, float* mean, __nv_bfloat16* xh, int ldh, cudaStream_t st) {
  const int wpb = 8, grid = (rows + wpb - 1) / wpb;
  const bool vec = (H % 128 == 0) && H <= 2048 && (ldh % 4 == 0) && !knobs().no_vec;
  const bool fast = (x == nullptr) && H >= 1024;   // bf16-operand mode at the wide shapes: SFU sigmoid
#define REPPO_FWD_VEC(V) do { if (fast) launch_k((norm_act_fwd_vec_kernel<NORM, V, true>), dim3(grid), dim3(wpb * 32), 0, st, z, gamma, beta, rows, eps, x, rstd, mean, xh, ldh); \
    else launch_k((norm_act_fwd_vec_kernel<NORM, V, false>), dim3(grid), dim3(wpb * 32), 0, st, z, gamma, beta, rows, eps, x, rstd, mean, xh, ldh); } while (0)
  if (vec && H == 128) REPPO_FWD_VEC(1);
  else if (vec && H == 256) REPPO_FWD_VEC(2);
  else if (vec && H == 384) REPPO_FWD_VEC(3);
  else if (vec && H == 512) REPPO_FWD_VEC(4);
  else if (vec && H == 768) REPPO_FWD_VEC(6);
  else if (vec && H == 1024) REPPO_FWD_VEC(8);
  else if (vec && H == 2048) REPPO_FWD_VEC(16);
  else launch_k((norm_act_fwd_kernel<NORM>), dim3(grid), dim3(wpb * 32), 0, st, z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh);
#undef REPPO_FWD_VEC
}

cudaError_t launch_norm_act_fwd(int norm, const float* z, const float* gamma, const float* beta, int rows, int H,
                                float eps, float* x, float* rstd, float* mean, void* x_h, int ldh, cudaStream_t st) {
  __nv_bfloat16* xh = static_cast<__nv_bfloat16*>(x_h);
  if (rows <= 0) return cudaSuccess;
  if (norm == NORM_NONE) launch_fwd_norm<NORM_NONE>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh, st);
  else if (norm == NORM_RMS) launch_fwd_norm<NORM_RMS>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh, st);
  else launch_fwd_norm<NORM_LAYER>(z, gamma, beta, rows, H, eps, x, rstd, mean, xh, ldh, st);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward: given dx = dL/d swish(y), y = norm(z): dz, plus (accumulated with atomics into the zeroed
// gradient arena) d gamma / d beta of the norm and d bias = column sums of dz.
//   none : dz = dx * swish'(z)
//   rms  : zh = z*rstd;           dzh = dy*g;  dz = rstd * (dzh - zh * mean(dzh*zh))
//   layer: zh = (z-mean)*rstd;    dz = rstd * (dzh - mean(dzh) - zh * mean(dzh*zh))
// Each block walks rows_per_block rows with 8 warps; the per-column partial sums stay in registers
// of the lane that owns the column, are combined across the block's warps in shared memory and
// leave as one atomicAdd per column per block.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxColsPerLane = 64;  // H <= 2048

template <int NORM>
__global__ void __launch_bounds__(256)
norm_act_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ dx2, const float* __restrict__ z,
                    const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ rstd_in,
                    const float* __restrict__ mean_in, int rows, int H, int rows_per_block, float* __restrict__ dz,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                    __nv_bfloat16* __restrict__ dz_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  extern __shared__ float sm[];  // [3][H]: dgamma, dbeta, dbias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  const int cpl = (H + 31) / 32;

  for (int c = threadIdx.x; c < 3 * H; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float pg[16], pb[16], pd[16];
  // To keep register use bounded for large H the column range is processed in chunks of 16 per lane.
  for (int cbase = 0; cbase < cpl; cbase += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { pg[j] = 0.f; pb[j] = 0.f; pd[j] = 0.f; }
    for (int row = r0 + warp; row < r1; row += nwarp) {
      const float* zr = z + static_cast<size_t>(row) * H;
      const float* dxr = dx + static_cast<size_t>(row) * H;
      const float* dx2r = dx2 ? dx2 + static_cast<size_t>(row) * H : nullptr;   // optional second upstream gradient (summed)
      float rstd = 1.f, mean = 0.f, s1 = 0.f, s2 = 0.f;
      if (NORM != NORM_NONE) {
        rstd = rstd_in[row];
        mean = (NORM == NORM_LAYER) ? mean_in[row] : 0.f;
        // pass 1 (whole row): m1 = mean(dzh), m2 = mean(dzh*zh)
        for (int c = lane; c < H; c += 32) {
          const float zh = (zr[c] - mean) * rstd;
          const float y = (NORM == NORM_LAYER) ? zh * gamma[c] + beta[c] : zh * gamma[c];
          const float dy = (dxr[c] + (dx2r ? dx2r[c] : 0.f)) * swish_grad_(y);
          const float dzh = dy * gamma[c];
          s1 += dzh; s2 += dzh * zh;
        }
        s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
        if (NORM == NORM_RMS) s1 = 0.f;
      }
      // pass 2 (this column chunk): dz and the column partial sums
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = lane + 32 * (cbase + j);
        if (c < H) {
          float o;
          const float dxv = dxr[c] + (dx2r ? dx2r[c] : 0.f);
          if (NORM == NORM_NONE) {
            o = dxv * swish_grad_(zr[c]);
          } else {
            const float zh = (zr[c] - mean) * rstd;
            const float y = (NORM == NORM_LAYER) ? zh * gamma[c] + beta[c] : zh * gamma[c];
            const float dy = dxv * swish_grad_(y);
            const float dzh = dy * gamma[c];
            o = rstd * (dzh - s1 - zh * s2);
            pg[j] += dy * zh;
            pb[j] += dy;
          }
          pd[j] += o;
          if (dz != nullptr) dz[static_cast<size_t>(row) * H + c] = o;
          if (dz_h != nullptr) dz_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = lane + 32 * (cbase + j);
      if (c < H) {
        if (NORM != NORM_NONE && dgamma != nullptr) atomicAdd(&sm[c], pg[j]);
        if (NORM == NORM_LAYER && dbeta != nullptr) atomicAdd(&sm[H + c], pb[j]);
        if (dbias != nullptr) atomicAdd(&sm[2 * H + c], pd[j]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    if (NORM != NORM_NONE && dgamma != nullptr) atomicAdd(dgamma + c, sm[c]);
    if (NORM == NORM_LAYER && dbeta != nullptr) atomicAdd(dbeta + c, sm[H + c]);
    if (dbias != nullptr) atomicAdd(dbias + c, sm[2 * H + c]);
  }
}

template <int NORM, int VPL>
__global__ void __launch_bounds__(256)
norm_act_bwd_vec_kernel(const float* __restrict__ dx, const float* __restrict__ dx2, const float* __restrict__ z,
                        const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ rstd_in,
                        const float* __restrict__ mean_in, int rows, int rows_per_block, float* __restrict__ dz,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                        __nv_bfloat16* __restrict__ dz_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  constexpr int H = 128 * VPL;
  __shared__ float sm[8][H];   // per-warp column partials (shared-memory float atomics are CAS loops: avoided)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);

  float g[VPL][4], b[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { g[i][j] = 1.f; b[i][j] = 0.f; }
    const int cg = 4 * (lane + 32 * i);  // gamma / beta: flat parameter arena, only 4-byte aligned
    if (NORM != NORM_NONE) { g[i][0] = gamma[cg]; g[i][1] = gamma[cg + 1]; g[i][2] = gamma[cg + 2]; g[i][3] = gamma[cg + 3]; }
    if (NORM == NORM_LAYER) { b[i][0] = beta[cg]; b[i][1] = beta[cg + 1]; b[i][2] = beta[cg + 2]; b[i][3] = beta[cg + 3]; }
  }
  float pg[VPL][4], pb[VPL][4], pd[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; pd[i][j] = 0.f; }

  for (int row = r0 + warp; row < r1; row += nwarp) {
    const float4* zr = reinterpret_cast<const float4*>(z + static_cast<size_t>(row) * H);
    const float4* dxr = reinterpret_cast<const float4*>(dx + static_cast<size_t>(row) * H);
    float4 zv[VPL], dv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) { zv[i] = zr[lane + 32 * i]; dv[i] = dxr[lane + 32 * i]; }
    if (dx2 != nullptr) {   // optional second upstream gradient (summed)
      const float4* dx2r = reinterpret_cast<const float4*>(dx2 + static_cast<size_t>(row) * H);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 t = dx2r[lane + 32 * i];
        dv[i].x += t.x; dv[i].y += t.y; dv[i].z += t.z; dv[i].w += t.w;
      }
    }
    float rstd = 1.f, mean = 0.f;
    if (NORM != NORM_NONE) { rstd = rstd_in[row]; if (NORM == NORM_LAYER) mean = mean_in[row]; }
    float zh[VPL][4], dy[VPL][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float zz[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
      const float dd[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (NORM == NORM_NONE) {
          zh[i][j] = zz[j];
          dy[i][j] = dd[j] * swish_grad_(zz[j]);
        } else {
          zh[i][j] = (zz[j] - mean) * rstd;
          const float y = zh[i][j] * g[i][j] + b[i][j];
          dy[i][j] = dd[j] * swish_grad_(y);
          const float dzh = dy[i][j] * g[i][j];
          s1 += dzh; s2 += dzh * zh[i][j];
        }
      }
    }
    if (NORM != NORM_NONE) {
      s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
      if (NORM == NORM_RMS) s1 = 0.f;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (NORM == NORM_NONE) {
          o[j] = dy[i][j];
        } else {
          o[j] = rstd * (dy[i][j] * g[i][j] - s1 - zh[i][j] * s2);
          pg[i][j] += dy[i][j] * zh[i][j];
          pb[i][j] += dy[i][j];
        }
        pd[i][j] += o[j];
      }
      if (dz != nullptr) reinterpret_cast<float4*>(dz + static_cast<size_t>(row) * H)[lane + 32 * i] = make_float4(o[0], o[1], o[2], o[3]);
      if (dz_h != nullptr) store_bf16x4(dz_h + static_cast<size_t>(row) * ldh + 4 * (lane + 32 * i), o[0], o[1], o[2], o[3]);
    }
  }
  // column sums: every warp parks its partials, the block adds the 8 copies and issues one global RED per column
  auto flush = [&](float (&acc)[VPL][4], float* __restrict__ dst) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      reinterpret_cast<float4*>(&sm[warp][0])[lane + 32 * i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sm[w][c];
      atomicAdd(dst + c, t);
    }
  };
  if (NORM != NORM_NONE && dgamma != nullptr) flush(pg, dgamma);
  if (NORM == NORM_LAYER && dbeta != nullptr) flush(pb, dbeta);
  if (dbias != nullptr) flush(pd, dbias);
}


// Wide rows (H = 512 * WPR, WPR warps per row): at H >= 1024 the one-warp-per-row kernel above needs > 230 registers
// (one block of 8 warps per SM, nothing to hide the load latency behind).  Here every warp owns a 512-column segment
// (VPL = 4, two blocks per SM); the row statistics of the norm backward are combined across the WPR warps of a row
// through a double-buffered shared-memory slot and ONE block barrier per row iteration.
// FAST (bf16-operand mode only): SFU sigmoid (ex2.approx / rcp.approx), far below the bf16 rounding of dz.
template <bool FAST>
REPPO_DEVINL float swish_grad_sel(float x) {
  if (FAST) {
    const float sg = __fdividef(1.f, 1.f + __expf(-fmaxf(x, -80.f)));
    return sg * (1.f + x * (1.f - sg));
  }
  return swish_grad_(x);
}

template <int NORM, int WPR, bool FAST>
__global__ void __launch_bounds__(256, 2)
norm_act_bwd_wide_kernel(const float* __restrict__ dx, const float* __restrict__ dx2, const float* __restrict__ z,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rstd_in,
                         const float* __restrict__ mean_in, int rows, int rows_per_block, float* __restrict__ dz,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                         __nv_bfloat16* __restrict__ dz_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  constexpr int VPL = 4, SEG = 128 * VPL, H = SEG * WPR, RPI = 8 / WPR;   // rows per block iteration
  __shared__ float xs[2][8][2];
  __shared__ float colp[8][SEG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = warp % WPR, rslot = warp / WPR;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  const int niter = (r1 - r0 + RPI - 1) / RPI;
  const int c0 = sub * SEG;

  float g[VPL][4], b[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { g[i][j] = 1.f; b[i][j] = 0.f; }
    const int cg = c0 + 4 * (lane + 32 * i);
    if (NORM != NORM_NONE) { g[i][0] = gamma[cg]; g[i][1] = gamma[cg + 1]; g[i][2] = gamma[cg + 2]; g[i][3] = gamma[cg + 3]; }
    if (NORM == NORM_LAYER) { b[i][0] = beta[cg]; b[i][1] = beta[cg + 1]; b[i][2] = beta[cg + 2]; b[i][3] = beta[cg + 3]; }
  }
  float pg[VPL][4], pb[VPL][4], pd[VPL][4];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; pd[i][j] = 0.f; }

  for (int it = 0; it < niter; ++it) {
    const int row = r0 + it * RPI + rslot;
    const bool valid = row < r1;
    float zh[VPL][4], dy[VPL][4];
    float s1 = 0.f, s2 = 0.f, rstd = 1.f, mean = 0.f;
    if (valid) {
      const float4* zr = reinterpret_cast<const float4*>(z + static_cast<size_t>(row) * H + c0);
      const float4* dxr = reinterpret_cast<const float4*>(dx + static_cast<size_t>(row) * H + c0);
      float4 zv[VPL], dv[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) { zv[i] = zr[lane + 32 * i]; dv[i] = dxr[lane + 32 * i]; }
      if (dx2 != nullptr) {
        const float4* dx2r = reinterpret_cast<const float4*>(dx2 + static_cast<size_t>(row) * H + c0);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const float4 t = dx2r[lane + 32 * i];
          dv[i].x += t.x; dv[i].y += t.y; dv[i].z += t.z; dv[i].w += t.w;
        }
      }
      if (NORM != NORM_NONE) { rstd = rstd_in[row]; if (NORM == NORM_LAYER) mean = mean_in[row]; }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float zz[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
        const float dd[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (NORM == NORM_NONE) {
            zh[i][j] = zz[j];
            dy[i][j] = dd[j] * swish_grad_sel<FAST>(zz[j]);
          } else {
            zh[i][j] = (zz[j] - mean) * rstd;
            const float y = zh[i][j] * g[i][j] + b[i][j];
            dy[i][j] = dd[j] * swish_grad_sel<FAST>(y);
            const float dzh = dy[i][j] * g[i][j];
            s1 += dzh; s2 += dzh * zh[i][j];
          }
        }
      }
    }
    if (NORM != NORM_NONE) {
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { xs[it & 1][warp][0] = s1; xs[it & 1][warp][1] = s2; }
      __syncthreads();
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) { s1 += xs[it & 1][rslot * WPR + w][0]; s2 += xs[it & 1][rslot * WPR + w][1]; }
      s1 *= (1.f / H); s2 *= (1.f / H);
      if (NORM == NORM_RMS) s1 = 0.f;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (NORM == NORM_NONE) {
            o[j] = dy[i][j];
          } else {
            o[j] = rstd * (dy[i][j] * g[i][j] - s1 - zh[i][j] * s2);
            pg[i][j] += dy[i][j] * zh[i][j];
            pb[i][j] += dy[i][j];
          }
          pd[i][j] += o[j];
        }
        const int c = c0 + 4 * (lane + 32 * i);
        if (dz != nullptr) *reinterpret_cast<float4*>(dz + static_cast<size_t>(row) * H + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (dz_h != nullptr) store_bf16x4(dz_h + static_cast<size_t>(row) * ldh + c, o[0], o[1], o[2], o[3]);
      }
    }
  }
  // column sums: the RPI warps that own the same column segment are added in shared memory, one global RED per column per block
  auto flush = [&](float (&acc)[VPL][4], float* __restrict__ dst) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      reinterpret_cast<float4*>(&colp[warp][0])[lane + 32 * i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      const int sg = c / SEG, cc = c - sg * SEG;
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < RPI; ++r) t += colp[r * WPR + sg][cc];
      atomicAdd(dst + c, t);
    }
  };
  if (NORM != NORM_NONE && dgamma != nullptr) flush(pg, dgamma);
  if (NORM == NORM_LAYER && dbeta != nullptr) flush(pb, dbeta);
  if (dbias != nullptr) flush(pd, dbias);
}

template <int NORM>
static cudaError_t launch_bwd_norm(const float* dx, const float* dx2, const float* z, const float* gamma, const float* beta, const float* rstd,
                                   const float* mean, int rows, int H, float* dz, float* dgamma, float* dbeta, float* dbias,
                                   __nv_bfloat16* dzh, int ldh, cudaStream_t st) {
  // one block per ~SM: enough parallelism, few enough blocks that the per-block column atomics stay cheap
  const int rpb_env = knobs().bwd_rpb;
  int rows_per_block = rpb_env > 0 ? rpb_env : 8;   // one row per warp: lowest latency (measured 134.4 vs 135.3 ms at 16)
  while ((rows + rows_per_block - 1) / rows_per_block > 148 * 4 && rows_per_block < 128) rows_per_block *= 2;
  const bool vec = (H % 128 == 0) && H <= 1024 && (ldh % 4 == 0) && !knobs().no_vec;
  if ((H == 1024 || H == 2048) && (ldh % 4 == 0) && !knobs().no_wide) {
    // wide rows: 2 / 4 warps per row; two blocks per SM worth of work items (C5: 225.7 ms, vs 227.8 / 228.2 / 230.2 at 4 / 8 / 16:
    // fewer blocks = fewer column-sum atomics)
    const int rpi = H == 1024 ? 4 : 2;
    const int bps = knobs().wide_bps;   // blocks per SM worth of work items
    int rpb = (rows + 148 * bps - 1) / (148 * bps);
    rpb = ((rpb + rpi - 1) / rpi) * rpi;
    if (rpb < 2 * rpi) rpb = 2 * rpi;
    const int gridw = (rows + rpb - 1) / rpb;
    const bool fast = (dz == nullptr);   // bf16-operand mode: only the bf16 twin of dz is produced
#define REPPO_BWD_WIDE(W, F) \
  launch_k((norm_act_bwd_wide_kernel<NORM, W, F>), dim3(gridw), dim3(256), 0, st, dx, dx2, z, gamma, beta, rstd, mean, rows, rpb, dz, dgamma, dbeta, dbias, dzh, ldh)
    if (H == 1024) { if (fast) REPPO_BWD_WIDE(2, true); else REPPO_BWD_WIDE(2, false); }
    else { if (fast) REPPO_BWD_WIDE(4, true); else REPPO_BWD_WIDE(4, false); }
#undef REPPO_BWD_WIDE
    return cudaGetLastError();
  }
  const int grid = (rows + rows_per_block - 1) / rows_per_block;
#define REPPO_BWD_VEC(V) \
  launch_k((norm_act_bwd_vec_kernel<NORM, V>), dim3(grid), dim3(256), 0, st, dx, dx2, z, gamma, beta, rstd, mean, rows, rows_per_block, dz, dgamma, dbeta, dbias, dzh, ldh)
  if (vec && H == 128) REPPO_BWD_VEC(1);
  else if (vec && H == 256) REPPO_BWD_VEC(2);
  else if (vec && H == 384) REPPO_BWD_VEC(3);
  else if (vec && H == 512) REPPO_BWD_VEC(4);
  else if (vec && H == 768) REPPO_BWD_VEC(6);
  else if (vec && H == 1024) REPPO_BWD_VEC(8);
  else {
    if (H > 32 * kMaxColsPerLane) return cudaErrorInvalidValue;
    const size_t smem = 3 * static_cast<size_t>(H) * sizeof(float);
    launch_k((norm_act_bwd_kernel<NORM>), dim3(grid), dim3(256), smem, st, dx, dx2, z, gamma, beta, rstd, mean, rows, H, rows_per_block, dz, dgamma, dbeta, dbias, dzh, ldh);
  }
#undef REPPO_BWD_VEC
  return cudaGetLastError();
}

cudaError_t launch_norm_act_bwd(int norm, const float* dx, const float* dx2, const float* z, const float* gamma, const float* beta,
                                const float* rstd, const float* mean, int rows, int H, float* dz, float* dgamma,
                                float* dbeta, float* dbias, void* dz_h, int ldh, cudaStream_t st) {
  __nv_bfloat16* dzh = static_cast<__nv_bfloat16*>(dz_h);
  if (rows <= 0) return cudaSuccess;
  if (norm == NORM_NONE) return launch_bwd_norm<NORM_NONE>(dx, dx2, z, gamma, beta, rstd, mean, rows, H, dz, dgamma, dbeta, dbias, dzh, ldh, st);
  if (norm == NORM_RMS) return launch_bwd_norm<NORM_RMS>(dx, dx2, z, gamma, beta, rstd, mean, rows, H, dz, dgamma, dbeta, dbias, dzh, ldh, st);
  return launch_bwd_norm<NORM_LAYER>(dx, dx2, z, gamma, beta, rstd, mean, rows, H, dz, dgamma, dbeta, dbias, dzh, ldh, st);
}

}  // namespace reppo
