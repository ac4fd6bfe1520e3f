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
// update (src/torchrl/reppo_util.py:432-466): batch mean and batch variance of x [rows, D] merged into the running
// statistics with Chan's parallel update.  ONE pass over x on the whole machine: the rows are cut into one contiguous
// chunk per block; inside a block T = (256 / D) * D threads sweep the chunk as a flat array with stride T, so every thread
// keeps ONE column and the reads are fully coalesced.  Each thread accumulates shifted sums  S1 = sum(x - K),
// S2 = sum((x - K)^2)  with K = the column's first value in the chunk (shifted-data variance: no cancellation when the
// mean is large against the spread); a block publishes per column (n, mean, M2) and the LAST block to finish (ticket
// counter) merges the blocks' triples in a fixed order with Chan's formula in double precision -- deterministic, no float
// atomics -- and then applies the reference's running-statistics update in its fp32 evaluation order.
// D > 256 (never on the named configurations) runs the same kernel with one block.
constexpr int kNormMaxBlocks = 148 * 4;
constexpr int kNormThreads = 256;
__device__ float g_norm_part[kNormMaxBlocks * kNormThreads * 3];   // [block][column][n, mean, M2]
__device__ unsigned int g_norm_ticket = 0;

__global__ void __launch_bounds__(kNormThreads)
obs_norm_update_kernel(const float* __restrict__ x, long long rows, int D, float* __restrict__ mean, float* __restrict__ var,
                       float* __restrict__ stdv, const long long* __restrict__ count, long long until) {
  REPPO_PDL_PROLOGUE();
  __shared__ float sh[3][kNormThreads];
  __shared__ double shd[3][kNormThreads];
  __shared__ bool is_last;
  const long long cnt = *count;
  if (until >= 0 && cnt >= until) return;
  const int t = threadIdx.x;
  const int G = kNormThreads / D;          // threads per column
  const int T = G * D;                     // active threads (a multiple of D: a thread's column never changes)
  const int c = t % D, g = t / D;
  const long long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = rows_per_block * blockIdx.x;
  const long long r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
  float n = 0.f, s1 = 0.f, s2 = 0.f, K = 0.f;
  if (t < T && r0 < r1) {
    const float* base = x + r0 * D;
    const long long total = (r1 - r0) * D;
    K = base[c];
    // element i = t + j T has column (t + j T) % D = c
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
    long long i = t;
    for (; i + T < total; i += 2 * T) {
      const float d0 = base[i] - K, d1 = base[i + T] - K;
      a1 += d0; a2 += d0 * d0; b1 += d1; b2 += d1 * d1;
      n += 2.f;
    }
    if (i < total) { const float d0 = base[i] - K; a1 += d0; a2 += d0 * d0; n += 1.f; }
    s1 = a1 + b1; s2 = a2 + b2;
  }
  sh[0][t] = n; sh[1][t] = s1; sh[2][t] = s2;
  __syncthreads();
  if (t < D) {
    // all threads of a column used the same shift K (the chunk's first row)
    float nn = 0.f, t1 = 0.f, t2 = 0.f;
    for (int q = 0; q < G; ++q) { nn += sh[0][t + q * D]; t1 += sh[1][t + q * D]; t2 += sh[2][t + q * D]; }
    float* out = g_norm_part + (static_cast<size_t>(blockIdx.x) * kNormThreads + t) * 3;
    const float m = (nn > 0.f) ? K + t1 / nn : 0.f;
    const float M2 = (nn > 0.f) ? fmaxf(t2 - t1 * t1 / nn, 0.f) : 0.f;
    out[0] = nn; out[1] = m; out[2] = M2;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = (atomicAdd(&g_norm_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- last block: merge the per-block triples (fixed order), then the reference's running update
  double an = 0.0, am = 0.0, aM2 = 0.0;
  if (t < T) {
    for (int b = g; b < static_cast<int>(gridDim.x); b += G) {
      const float* in = g_norm_part + (static_cast<size_t>(b) * kNormThreads + c) * 3;
      const double bn = __ldcg(in), bmn = __ldcg(in + 1), bM2 = __ldcg(in + 2);
      if (bn > 0.0) {
        const double tot = an + bn, dl = bmn - am;
        am += dl * bn / tot;
        aM2 += bM2 + dl * dl * an * bn / tot;
        an = tot;
      }
    }
  }
  shd[0][t] = an; shd[1][t] = am; shd[2][t] = aM2;
  __syncthreads();
  if (t < D) {
    an = 0.0; am = 0.0; aM2 = 0.0;
    for (int q = 0; q < G; ++q) {
      const double bn = shd[0][t + q * D], bmn = shd[1][t + q * D], bM2 = shd[2][t + q * D];
      if (bn > 0.0) {
        const double tot = an + bn, dl = bmn - am;
        am += dl * bn / tot;
        aM2 += bM2 + dl * dl * an * bn / tot;
        an = tot;
      }
    }
    const float bm = static_cast<float>(am);                                     // torch.mean(x, dim=0)
    const float bv = static_cast<float>(aM2 / static_cast<double>(rows));         // torch.mean((x - batch_mean)^2, dim=0)
    const float m_old = mean[t];
    const float new_cnt = static_cast<float>(cnt + rows);
    const float m_new = m_old + (static_cast<float>(rows) / new_cnt) * (bm - m_old);
    float v;
    if (cnt > 0) {
      const float d2 = bm - m_new;
      const float m_a = var[t] * static_cast<float>(cnt), m_b = bv * static_cast<float>(rows);
      const float w = static_cast<float>(cnt * rows) / new_cnt;
      v = (m_a + m_b + d2 * d2 * w) / new_cnt;
    } else {
      // first batch: the variance about the UPDATED running mean (reppo_util.py:461-462) = batch variance + (bm - m_new)^2
      const float d2 = bm - m_new;
      v = bv + d2 * d2;
    }
    mean[t] = m_new; var[t] = v; stdv[t] = sqrtf(v);
  }
  if (t == 0) g_norm_ticket = 0u;   // replayable (CUDA graphs, the next call)
}

// D > 256: one block per 32 columns, two passes (never on the named configurations)
__global__ void __launch_bounds__(256)
obs_norm_update_wide_kernel(const float* __restrict__ x, long long rows, int D, float* __restrict__ mean, float* __restrict__ var,
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
  // grid-stride over the flat array; the column index advances by (stride % D) per step (no 64-bit modulo in the loop) and
  // four independent loads are in flight per thread
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  int c = static_cast<int>(i % D);
  const int dc = static_cast<int>(stride % D);
  auto step = [&](int cc) { cc += dc; return cc >= D ? cc - D : cc; };
  for (; i + 3 * stride < total; i += 4 * stride) {
    const int c1 = step(c), c2 = step(c1), c3 = step(c2);
    const float v0 = x[i], v1 = x[i + stride], v2 = x[i + 2 * stride], v3 = x[i + 3 * stride];
    out[i] = (center ? v0 - mean[c] : v0) / (stdv[c] + eps);
    out[i + stride] = (center ? v1 - mean[c1] : v1) / (stdv[c1] + eps);
    out[i + 2 * stride] = (center ? v2 - mean[c2] : v2) / (stdv[c2] + eps);
    out[i + 3 * stride] = (center ? v3 - mean[c3] : v3) / (stdv[c3] + eps);
    c = step(c3);
  }
  for (; i < total; i += stride) {
    const float v = x[i];
    out[i] = (center ? v - mean[c] : v) / (stdv[c] + eps);
    c = step(c);
  }
  if (count != nullptr && add_rows > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long cnt = *count;
    if (!(until >= 0 && cnt >= until)) *count = cnt + add_rows;
  }
}

cudaError_t launch_obs_normalize(const float* x, long long rows, int D, float* mean, float* var, float* stdv, long long* count,
                                 float eps, int update, int center, long long until, float* out, cudaStream_t st) {
  if (rows <= 0 || D <= 0) return cudaSuccess;
  if (update) {
    if (D <= kNormThreads) {
      // >= 64 rows per block; at most kNormMaxBlocks blocks (4 per SM).  The per-block scratch is a device global: calls must
      // be stream-ordered against each other (the collection loop normalises on one stream)
      const long long want_b = (rows + 63) / 64;
      const int blocks = static_cast<int>(want_b < kNormMaxBlocks ? want_b : kNormMaxBlocks);
      launch_k((obs_norm_update_kernel), dim3(blocks), dim3(kNormThreads), 0, st, x, rows, D, mean, var, stdv, count, until);
    } else {
      launch_k((obs_norm_update_wide_kernel), dim3((D + 31) / 32), dim3(256), 0, st, x, rows, D, mean, var, stdv, count, until);
    }
  }
  const long long total = rows * D;
  const long long want = (total + 1023) / 1024;
  const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  launch_k((obs_norm_apply_kernel), dim3(grid), dim3(256), 0, st, x, total, D, mean, stdv, eps, center, out, update ? count : nullptr,
           update ? rows : 0LL, until);
  return cudaGetLastError();
}

}  // namespace reppo
