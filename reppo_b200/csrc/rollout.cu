// Rollout-side kernels: what the collection loop runs per environment step around the
// network forwards.
//
//  policy_sample_kernel : tanh-Gaussian action sample + log-probabilities
//        torch: collect_fn, src/torchrl/reppo.py:107-109 (sample), :146 (log_prob of the action clipped to 0.999),
//               :131-134 (next action, log_prob at the 1 - 1e-6 clip)
//        jax  : step_env, src/jaxrl/reppo.py:347-364 (exploration-scaled policy, importance weights from the log-prob of
//               the clipped action under the unscaled and the scaled policy), :367-369 (sample_and_log_prob)
//  soft_reward_kernel   : reward - gamma * log_prob * exp(log_temp)           (:136-138 torch, :371-374 jax)
//  obs_norm_update / obs_norm_apply : EmpiricalNormalization, src/torchrl/reppo_util.py:394-471 (Chan's parallel
//        mean / variance update, then (x - mean) / (std + eps))
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "loss_math.cuh"

namespace reppo {

// one warp per row, lane k owns action dimension k (+32 ...)
__global__ void __launch_bounds__(256)
policy_sample_kernel(const float* __restrict__ out, const float* __restrict__ eps, const float* __restrict__ scale, int rows,
                     int A, float min_std, float clip, int store_clipped, float* __restrict__ action, int act_ld,
                     __nv_bfloat16* __restrict__ act_h, int act_ldh, float* __restrict__ logp, float* __restrict__ logp_scaled) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float* o = out + static_cast<size_t>(row) * 2 * A;
  const float sc = scale ? scale[row] : 1.f;
  float lp = 0.f, lps = 0.f;
  for (int k = lane; k < A; k += 32) {
    const float mu = o[k], sig = expf(o[A + k]) + min_std, sig_s = sig * sc;
    const float e = eps[static_cast<size_t>(row) * A + k];
    const float x = mu + sig_s * e;
    const float a = tanhf(x);
    float a_st = a;
    if (clip > 0.f) {
      const float ac = fminf(fmaxf(a, -clip), clip);
      if (store_clipped) a_st = ac;
      const float u = atanhf(ac);
      const float fu = fldj_(u);
      const float z = (u - mu) / sig, zs = (u - mu) / sig_s;
      lp += -0.5f * z * z - logf(sig) - kHalfLog2Pi - fu;
      lps += -0.5f * zs * zs - logf(sig_s) - kHalfLog2Pi - fu;
    } else {
      // sample_and_log_prob: base log-prob at the pre-tanh sample minus the tanh log-det-Jacobian
      const float l = -0.5f * e * e - logf(sig_s) - kHalfLog2Pi - fldj_(x);
      lp += l; lps += l;
    }
    action[static_cast<size_t>(row) * act_ld + k] = a_st;
    if (act_h != nullptr) act_h[static_cast<size_t>(row) * act_ldh + k] = __float2bfloat16_rn(a_st);
  }
  lp = warp_sum(lp); lps = warp_sum(lps);
  if (lane == 0) {
    if (logp != nullptr) logp[row] = lp;
    if (logp_scaled != nullptr) logp_scaled[row] = lps;
  }
}

cudaError_t launch_policy_sample(const float* out, const float* eps, const float* scale, int rows, int A, float min_std, float clip,
                                 int store_clipped, float* action, int act_ld, void* act_h, int act_ldh, float* logp,
                                 float* logp_scaled, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  launch_k((policy_sample_kernel), dim3((rows + 7) / 8), dim3(256), 0, st, out, eps, scale, rows, A, min_std, clip, store_clipped,
           action, act_ld, static_cast<__nv_bfloat16*>(act_h), act_ldh, logp, logp_scaled);
  return cudaGetLastError();
}

__global__ void soft_reward_kernel(const float* __restrict__ reward, const float* __restrict__ logp,
                                   const float* __restrict__ log_temp, float gamma, int rows, float* __restrict__ out) {
  REPPO_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  // (gamma * log_prob) * temperature, in the reference's evaluation order
  out[i] = reward[i] - gamma * logp[i] * expf(log_temp[0]);
}

cudaError_t launch_soft_reward(const float* reward, const float* logp, const float* log_temp, float gamma, int rows, float* out,
                               cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  launch_k((soft_reward_kernel), dim3((rows + 255) / 256), dim3(256), 0, st, reward, logp, log_temp, gamma, rows, out);
  return cudaGetLastError();
}

// ---- EmpiricalNormalization --------------------------------------------------------------------------
// One block per 32 columns: threads (32 columns x 8 row groups) sweep the rows with coalesced 128 B reads, two passes
// (batch mean, then centred second moment) like torch.mean / torch.mean((x - m)^2), then thread (c, 0) merges into the
// running statistics with Chan's update in the reference's fp32 evaluation order.
__global__ void __launch_bounds__(256)
obs_norm_update_kernel(const float* __restrict__ x, long long rows, int D, float* __restrict__ mean, float* __restrict__ var,
                       float* __restrict__ stdv, const long long* __restrict__ count, long long until) {
  REPPO_PDL_PROLOGUE();
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long cnt = *count;
  if (until >= 0 && cnt >= until) return;
  float s = 0.f;
  if (c < D)
    for (long long r = rg; r < rows; r += 8) s += x[r * D + c];
  red[rg][cl] = s;
  __syncthreads();
  float bm = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g) bm += red[g][cl];
  bm /= static_cast<float>(rows);
  __syncthreads();
  const float m_old = (c < D) ? mean[c] : 0.f;
  const float new_cnt = static_cast<float>(cnt + rows);
  const float m_new = m_old + (static_cast<float>(rows) / new_cnt) * (bm - m_old);
  // first batch: variance about the UPDATED running mean (reppo_util.py:461-462); later: about the batch mean (:452)
  const float centre = (cnt > 0) ? bm : m_new;
  float q = 0.f;
  if (c < D)
    for (long long r = rg; r < rows; r += 8) { const float d = x[r * D + c] - centre; q += d * d; }
  red[rg][cl] = q;
  __syncthreads();
  if (rg == 0 && c < D) {
    float bv = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) bv += red[g][cl];
    bv /= static_cast<float>(rows);
    float v;
    if (cnt > 0) {
      const float d2 = bm - m_new;
      const float m_a = var[c] * static_cast<float>(cnt), m_b = bv * static_cast<float>(rows);
      const float w = static_cast<float>(cnt * rows) / new_cnt;
      v = (m_a + m_b + d2 * d2 * w) / new_cnt;
    } else {
      v = bv;
    }
    mean[c] = m_new; var[c] = v; stdv[c] = sqrtf(v);
  }
}

__global__ void obs_norm_apply_kernel(const float* __restrict__ x, long long total, int D, const float* __restrict__ mean,
                                      const float* __restrict__ stdv, float eps, int center, float* __restrict__ out,
                                      long long* __restrict__ count, long long add_rows, long long until) {
  REPPO_PDL_PROLOGUE();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D);
    const float v = x[i];
    out[i] = (center ? v - mean[c] : v) / (stdv[c] + eps);
  }
  if (count != nullptr && add_rows > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long cnt = *count;
    if (!(until >= 0 && cnt >= until)) *count = cnt + add_rows;
  }
}

cudaError_t launch_obs_normalize(const float* x, long long rows, int D, float* mean, float* var, float* stdv, long long* count,
                                 float eps, int update, int center, long long until, float* out, cudaStream_t st) {
  if (rows <= 0 || D <= 0) return cudaSuccess;
  if (update) launch_k((obs_norm_update_kernel), dim3((D + 31) / 32), dim3(256), 0, st, x, rows, D, mean, var, stdv, count, until);
  const long long total = rows * D;
  const long long want = (total + 255) / 256;
  const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  launch_k((obs_norm_apply_kernel), dim3(grid), dim3(256), 0, st, x, total, D, mean, stdv, eps, center, out, update ? count : nullptr,
           update ? rows : 0LL, until);
  return cudaGetLastError();
}

}  // namespace reppo
