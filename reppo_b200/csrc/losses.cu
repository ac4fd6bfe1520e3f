// Fused loss kernels of the REPPO update (K6-K11).
//
//  critic_ce_kernel  : categorical value head + HL-Gauss target + cross-entropy, forward AND
//                      d logits, one warp per row (src/jaxrl/reppo.py:474-485,505-513,
//                      src/jaxrl/utils.py:43-59; src/torchrl/reppo.py:243-254,
//                      src/torchrl/reppo_util.py:756-777; value head jax_models.py:297-310).
//  critic_aux_kernel : next-embedding (+ reward) prediction loss and d pred (reppo.py:492-502;
//                      torchrl/reppo.py:255-262).
//  actor_sample_kernel / actor_q_kernel / actor_grad_kernel : tanh-Gaussian reparameterised
//                      sample + log-prob, 16-sample forward KL to the behaviour policy, the
//                      KL-clipped pathwise loss with temperature / Lagrange terms, and their
//                      hand-derived gradients (reppo.py:526-622; torchrl/reppo.py:291-338;
//                      distrax Tanh fldj = 2 (log 2 - x - softplus(-2x)), torch_models.py:52-55).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "loss_math.cuh"

namespace reppo {

// block-level accumulation of per-warp metric partials -> one atomicAdd per metric per block
template <int NM>
REPPO_DEVINL void block_metric_add(float (&vals)[NM], float* sm /* [8][NM] */, float* const* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NM; ++k) sm[warp * NM + k] = vals[k];
  __syncthreads();
  if (threadIdx.x < NM) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += sm[w * NM + threadIdx.x];
    if (dst[threadIdx.x] != nullptr) atomicAdd(dst[threadIdx.x], s);
  }
}

constexpr int kLossRowsPerBlock = 8;   // one row per warp: these kernels are latency-bound per row, parallelism wins
constexpr int kLossMaxBlocks = 148 * 8;   // beyond that the blocks walk the rows grid-stride (large-batch sweep)

// dbias (+)= column sums of dlogits (bias gradient of the head's last Linear); dzero (+)= bias_scale * the same
// (gradient of zero_dist, which enters the logits as bias_scale * zero_dist).  Both nullable.
template <int NB>   // bins per lane the row loops are unrolled for: ceil(num_bins / 32) <= NB
__global__ void __launch_bounds__(256, NB <= 5 ? 4 : 3)
critic_ce_kernel(const float* __restrict__ head_out, int ld, const float* __restrict__ zero_dist, HLConsts hl,
                 const float* __restrict__ target, const float* __restrict__ truncated, const int* __restrict__ idx,
                 int rows, float inv_rows, int no_mask, float* __restrict__ dlogits, float* __restrict__ metrics,
                 __nv_bfloat16* __restrict__ dl_h, int ldh, float* __restrict__ dbias, float* __restrict__ dzero) {
  REPPO_PDL_PROLOGUE();
  __shared__ float sm[8 * 5];
  __shared__ float smax[8], smin[8];
  __shared__ float scol[8][32 * NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int B = hl.num_bins;
  float vals[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float tmax = -INFINITY, tmin = INFINITY;
  float cs[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) cs[j] = 0.f;
  const float inv_bw = static_cast<float>(B) / (hl.edges[B] - hl.edges[0]);
  const float inv_den = 1.f / hl.den;
  // row-invariant per-lane constants: the logit bias (bias_scale * zero_dist) and the support of this lane's bins
  float zb[NB], sup[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int c = lane + 32 * j;
    zb[j] = (c < B) ? hl.bias_scale * zero_dist[c] : 0.f;
    sup[j] = (c < B) ? hl.support[c] : 0.f;
  }
  const int jB = B >> 5, laneB = B & 31;   // where the erf of the top edge lives
  // grid-stride over rows (the launch caps the grid at a few blocks per SM): the per-block metric / column-sum atomics
  // below are issued once per block, not once per 8 rows
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const long long src = idx ? idx[row] : row;
    const float tgt = target[src];
    const float mask = no_mask ? 1.f : 1.f - truncated[src];
    RowSoftmaxT<NB> sx;
    row_softmax_pre(head_out + static_cast<size_t>(row) * ld, zb, sup, B, lane, sx);
    // HL-Gauss histogram of the clipped target
    const float x = fminf(fmaxf(tgt, hl.vmin), hl.vmax);
    float ce = 0.f, sum_tp = 0.f, tp[NB];
    // erf at the lower edge of this lane's bins; the upper edge of bin c is the lower edge of bin c + 1 = the next lane's
    // value (lane 31: this lane's value of the next group of 32 bins), so every edge is evaluated once.
    // Only the edges near the target need the evaluation at all: sigma = 0.75 bin widths (utils.py:47, reppo_util.py:766),
    // so an edge more than 5.5 bins away from x sits at |t| > 5.1 where erff(t) is exactly +-1.0f (it saturates at
    // |t| >= 3.92); those edges take the constant, and whole groups of 32 edges outside the window skip erff.
    const float pos = (x - hl.edges[0]) * inv_bw;   // target position in units of bins (edge index scale)
    float lo_e[NB + 1];
    float e_top = 0.f;
#pragma unroll
    for (int j = 0; j <= NB; ++j) {
      const int c = lane + 32 * j;
      const float dc = static_cast<float>(c) - pos;
      float e = (dc > 0.f) ? 1.f : -1.f;
      if (fabsf(dc) < 5.5f && c <= B) e = erff((hl.edges[c] - x) * inv_den);
      lo_e[j] = (c <= B) ? e : 0.f;
      if (j == jB) e_top = lo_e[j];
    }
    // z = erf(top edge) - erf(bottom edge) (+ eps): both already sit in registers (lane laneB group jB, lane 0 group 0)
    const float z = __shfl_sync(0xffffffffu, e_top, laneB) - __shfl_sync(0xffffffffu, lo_e[0], 0) + hl.z_eps;
    const float inv_z = 1.f / z;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int c = lane + 32 * j;
      tp[j] = 0.f;
      const float nxt = __shfl_down_sync(0xffffffffu, lo_e[j], 1);
      const float wrap = __shfl_sync(0xffffffffu, lo_e[j + 1], 0);
      if (c < B) {
        const float lo = lo_e[j], hi = (lane == 31) ? wrap : nxt;
        if (hi != lo) {   // bins outside the window have probability exactly 0
          tp[j] = (hi - lo) * inv_z;
          ce -= tp[j] * (sx.lg[j] - sx.lse);
          sum_tp += tp[j];
        }
      }
    }
    ce = warp_sum(ce);
    sum_tp = warp_sum(sum_tp);
    const float coef = mask * inv_rows;
    float* drow = dlogits + static_cast<size_t>(row) * ld;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int c = lane + 32 * j;
      if (c < B) {
        const float o = coef * (sum_tp * sx.p[j] - tp[j]);
        drow[c] = o;
        cs[j] += o;
        if (dl_h != nullptr) dl_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
      }
    }
    vals[0] += mask * ce * inv_rows;                         // REPPO_M_CRITIC_LOSS (CE part)
    vals[1] += ce * inv_rows;                                // REPPO_M_CE_MEAN
    vals[2] += (sx.value - tgt) * (sx.value - tgt) * inv_rows;  // REPPO_M_VALUE_LOSS
    vals[3] += sx.value * inv_rows;                          // REPPO_M_Q_MEAN
    vals[4] += tgt * inv_rows;                               // REPPO_M_TARGET_MEAN
    tmax = fmaxf(tmax, tgt); tmin = fminf(tmin, tgt);
  }
  float* dst[5] = {metrics + M_CRITIC_LOSS, metrics + M_CE_MEAN, metrics + M_VALUE_LOSS, metrics + M_Q_MEAN,
                   metrics + M_TARGET_MEAN};
  if (lane == 0) { smax[warp] = tmax; smin[warp] = tmin; }
#pragma unroll
  for (int j = 0; j < NB; ++j) scol[warp][lane + 32 * j] = cs[j];
  block_metric_add<5>(vals, sm, dst);   // contains the __syncthreads that publishes smax / smin / scol
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { tmax = fmaxf(tmax, smax[w]); tmin = fminf(tmin, smin[w]); }
    if (tmax > -INFINITY) { atomic_max_float(metrics + M_TARGET_MAX, tmax); atomic_min_float(metrics + M_TARGET_MIN, tmin); }
  }
  if (dbias != nullptr) {
    for (int c = threadIdx.x; c < B; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += scol[w][c];
      atomicAdd(dbias + c, t);
      if (dzero != nullptr) atomicAdd(dzero + c, t * hl.bias_scale);
    }
  }
}

// pred [rows, P]: JAX P = H + 1 (slot 0 = reward), Torch P = H
__global__ void __launch_bounds__(256)
critic_aux_kernel(const float* __restrict__ pred, int P, int H, int reward_head, int mask_done,
                  const float* __restrict__ next_emb, const float* __restrict__ reward, const float* __restrict__ done,
                  const float* __restrict__ truncated, const int* __restrict__ idx, int rows, float inv_rows, int no_mask,
                  float aux_mult, float* __restrict__ dpred, float* __restrict__ metrics, __nv_bfloat16* __restrict__ dp_h,
                  int ldh, float* __restrict__ dbias) {
  REPPO_PDL_PROLOGUE();
  __shared__ float sm[8 * 3];
  extern __shared__ float scol[];  // [8 warps][P]: each warp parks its row of dpred (only when dbias != nullptr)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float vals[3] = {0.f, 0.f, 0.f};
  if (dbias != nullptr)
    for (int c = lane; c < P; c += 32) scol[warp * P + c] = 0.f;   // rows past the end contribute nothing
  __syncwarp();   // with the reward head (off = 1) a lane accumulates into columns another lane zeroed
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const long long src = idx ? idx[row] : row;
    const float mask = no_mask ? 1.f : 1.f - truncated[src];
    const float nd = mask_done ? 1.f - done[src] : 1.f;
    const float* pr = pred + static_cast<size_t>(row) * P;
    const float* ne = next_emb + static_cast<size_t>(src) * H;
    float* dr = dpred + static_cast<size_t>(row) * P;
    const int off = reward_head ? 1 : 0;
    const float g = 2.f * mask * nd * aux_mult * inv_rows / static_cast<float>(P);
    float sq = 0.f;
#pragma unroll 8
    for (int c = lane; c < H; c += 32) {
      const float d = pr[off + c] - ne[c];
      sq += d * d;
      dr[off + c] = g * d;
      if (dbias != nullptr) scol[warp * P + off + c] += g * d;
      if (dp_h != nullptr) dp_h[static_cast<size_t>(row) * ldh + off + c] = __float2bfloat16_rn(g * d);
    }
    float r = 0.f;
    if (reward_head) {
      r = reward[src];
      if (lane == 0) {
        const float d = pr[0] - r;
        sq += d * d;
        dr[0] = g * d;
        if (dbias != nullptr) scol[warp * P] += g * d;
        if (dp_h != nullptr) dp_h[static_cast<size_t>(row) * ldh] = __float2bfloat16_rn(g * d);
      }
    }
    sq = warp_sum(sq);
    const float aux = nd * sq / static_cast<float>(P);
    vals[0] += mask * aux_mult * aux * inv_rows;                    // joins REPPO_M_CRITIC_LOSS
    vals[1] += (reward_head ? aux : mask * aux) * inv_rows;         // REPPO_M_AUX_MEAN (jax: mean aux; torch: embedding_loss)
    vals[2] += r * inv_rows;                                        // REPPO_M_REWARD_MEAN
  }
  float* dst[3] = {metrics + M_CRITIC_LOSS, metrics + M_AUX_MEAN, metrics + M_REWARD_MEAN};
  block_metric_add<3>(vals, sm, dst);   // contains a __syncthreads: scol is complete afterwards
  if (dbias != nullptr)
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += scol[w * P + c];
      atomicAdd(dbias + c, t);
    }
}

// -------------------------------------------------------------------------------------------------
// actor: sample + log-prob + KL estimate (thread per row)
// -------------------------------------------------------------------------------------------------

// A block of 256 threads owns 16 rows.  Phase 1: one thread per (row, action dimension) evaluates everything that needs
// a transcendental -- the policy sample a = tanh(mu + sig eps), its log-prob term, and the per-dimension constants of the
// current and the behaviour policy -- and parks the constants in shared memory.  Phase 2: thread (row, lane sl) owns KL
// sample sl (+16, ...) of its row, reads that sample's A normals as one contiguous run and accumulates the row's KL term
// and, per dimension, d lp_new / d mu and d lp_new / d sigma; the 16 lanes of a row meet through shared memory (one
// column sum per thread) instead of 2 A shuffle reductions.
//
// KL samples (src/jaxrl/reppo.py:555-571, src/torchrl/reppo.py:306-312).  The reference evaluates
//   u = atanh(clip(tanh(x_o), +-c)),  lp_old = N(.; mu_o, sig_o) - fldj,  lp_new = N(u; mu, sig) - fldj(u).
// Two exact identities remove every transcendental from the 16 x A inner loop:
//   (1) atanh(clip(tanh(x), +-c)) = clamp(x, +-atanh(c))  (tanh is monotone), and
//   (2) fldj(u) appears in lp_old and lp_new with opposite signs whenever both are evaluated at u: Torch semantics always,
//       JAX semantics whenever the sample is not clipped (there lp_old uses fldj(x_o) and u = x_o).  Only a clipped JAX
//       sample (|x_o| > atanh(1 - 1e-4) = 4.95) keeps the correction fldj(x_o) - fldj(u).
// In fp32 the reference's tanh -> atanh round trip carries a per-row error of 2e-5 .. 2e-3 against a float64 evaluation
// (1 - tanh(x) loses relative precision as |x| grows); the identity form is within 1e-6 of float64, and the two agree to
// <= 3e-6 on the batch-mean KL -- inside the fp32 tolerance the parity tests state (rel 2e-4 + 2e-6 on `kl`; the tests
// additionally hold the kernel to 5e-5 against a float64 evaluation of the oracle).
constexpr int kGradMaxA = 64;      // action_dim <= 64 on the actor-loss path
constexpr int kSampleRows = 16;    // rows per block
constexpr int kSampleChunk = 8;    // action dimensions per pass of phase 2
__global__ void __launch_bounds__(256, 4)
actor_sample_kernel(const float* __restrict__ out, const float* __restrict__ out_old, const int* __restrict__ old_idx,
                    const float* __restrict__ eps_pi, const float* __restrict__ eps_kl, int rows, int A, int kl_samples,
                    ActorConsts ac, float* __restrict__ act_out, int act_ld, float* __restrict__ logp_out,
                    float* __restrict__ kl_out, float* __restrict__ gmu_kl, float* __restrict__ gsig_kl,
                    __nv_bfloat16* __restrict__ act_h, int act_ldh) {
  REPPO_PDL_PROLOGUE();
  __shared__ float4 cst[kSampleRows][kGradMaxA];                     // (mu, 1 / sig, mu_o, sig_o)
  __shared__ float2 rowk[kSampleRows][kGradMaxA];                    // (log-prob term, log sig - log sig_o) per dimension
  __shared__ float red[kSampleRows][16][2 * kSampleChunk + 1];       // per (row, sample lane): gm[8] | gs[8] (+1: banks)
  const int t = threadIdx.x;
  const int r = t >> 4, sl = t & 15;
  const float u_max = atanhf(ac.action_clip);
  const float inv_s = 1.f / static_cast<float>(kl_samples);
  const bool vec2 = ((A & 1) == 0) && ((reinterpret_cast<uintptr_t>(eps_kl) & 7) == 0);
  for (int base = blockIdx.x * kSampleRows; base < rows; base += gridDim.x * kSampleRows) {
    // ---- phase 1: one thread per (row, dimension)
    for (int i = t; i < kSampleRows * A; i += 256) {
      const int rr = i / A, k = i - rr * A, row = base + rr;
      if (row < rows) {
        const float* o = out + static_cast<size_t>(row) * 2 * A;
        const float* oo = out_old + static_cast<size_t>(old_idx ? old_idx[row] : row) * 2 * A;
        const float mu = o[k], sig = expf(o[A + k]) + ac.min_std;
        const float mu_o = oo[k], sig_o = expf(oo[A + k]) + ac.min_std;
        const float log_sig = logf(sig);
        cst[rr][k] = make_float4(mu, 1.f / sig, mu_o, sig_o);
        const float e = eps_pi[static_cast<size_t>(row) * A + k];
        const float x = mu + sig * e;
        const float a = tanhf(x);
        act_out[static_cast<size_t>(row) * act_ld + k] = a;
        if (act_h != nullptr) act_h[static_cast<size_t>(row) * act_ldh + k] = __float2bfloat16_rn(a);
        float lp;
        if (ac.clip_policy_sample) {
          // torchrl/reppo.py:301: log_prob(actions.clip(+-c)) -> base.log_prob(atanh(.)) - fldj(atanh(.))
          const float xc = atanhf(fminf(fmaxf(a, -ac.action_clip), ac.action_clip));
          const float zz = (xc - mu) / sig;
          lp = -0.5f * zz * zz - log_sig - kHalfLog2Pi - fldj_(xc);
        } else {
          // distrax sample_and_log_prob: base log-prob at the pre-tanh sample minus fldj
          lp = -0.5f * e * e - log_sig - kHalfLog2Pi - fldj_(x);
        }
        rowk[rr][k] = make_float2(lp, log_sig - logf(sig_o));
      }
    }
    __syncthreads();
    // ---- phase 2: thread (r, sl) = KL sample sl (+16, ...) of row base + r
    const int row = base + r;
    const bool ok = row < rows;
    float dsum = 0.f;   // sum over this lane's samples and all dimensions of (lp_old - lp_new) minus the constant part
    for (int c0 = 0; c0 < A; c0 += kSampleChunk) {
      float gm[kSampleChunk], gs[kSampleChunk];
#pragma unroll
      for (int j = 0; j < kSampleChunk; ++j) { gm[j] = 0.f; gs[j] = 0.f; }
      if (ok) {
        for (int sidx = sl; sidx < kl_samples; sidx += 16) {
          const float* ep = eps_kl + (static_cast<size_t>(sidx) * rows + row) * A + c0;
          float eo[kSampleChunk];
          if (vec2 && c0 + kSampleChunk <= A) {
#pragma unroll
            for (int j = 0; j < kSampleChunk; j += 2) {
              const float2 v = *reinterpret_cast<const float2*>(ep + j);
              eo[j] = v.x; eo[j + 1] = v.y;
            }
          } else if (vec2) {
#pragma unroll
            for (int j = 0; j < kSampleChunk; j += 2) {
              float2 v = make_float2(0.f, 0.f);
              if (c0 + j < A) v = *reinterpret_cast<const float2*>(ep + j);   // A even: the pair is inside or outside
              eo[j] = v.x; eo[j + 1] = v.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < kSampleChunk; ++j) eo[j] = (c0 + j < A) ? ep[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < kSampleChunk; ++j) {
            if (c0 + j < A) {
              const float4 cc = cst[r][c0 + j];
              if (ac.reverse_kl) {
                // sample of the CURRENT policy, gradient through it: kl = lp_pi(x) - lp_old(clip(tanh x)); per dimension
                //   0.5 zo^2 - 0.5 eps^2 - (log sig - log sig_o) - (fldj(x) - fldj(u)),  zo = (u - mu_o) / sig_o
                // (the log-sigma part is row-constant and joins below); gm / gs hold -d kl / d(mu, sigma) because the
                // gradient kernel multiplies them by d L / d lp_new = -d L / d kl.
                const float sig = 1.f / cc.y;
                const float xs = cc.x + sig * eo[j];
                const float u = fminf(fmaxf(xs, -u_max), u_max);
                const float zo = (u - cc.z) / cc.w;
                dsum += 0.5f * zo * zo - 0.5f * eo[j] * eo[j];
                float dx;
                if (u != xs) {   // clipped: the old policy's term is constant, -fldj(x) keeps its slope 2 tanh(x)
                  dsum -= fldj_(xs) - fldj_(u);
                  dx = 2.f * tanhf(xs);
                } else {
                  dx = zo / cc.w;
                }
                gm[j] -= dx;
                gs[j] -= dx * eo[j] - cc.y;
                continue;
              }
              const float xo = cc.z + cc.w * eo[j];
              const float u = fminf(fmaxf(xo, -u_max), u_max);
              const float zn = (u - cc.x) * cc.y;
              if (ac.old_logp_at_clipped) {
                const float zo = (u - cc.z) / cc.w;
                dsum += 0.5f * zn * zn - 0.5f * zo * zo;
              } else {
                dsum += 0.5f * zn * zn - 0.5f * eo[j] * eo[j];
                if (u != xo) dsum -= fldj_(xo) - fldj_(u);   // clipped sample: the two fldj terms no longer cancel
              }
              gm[j] += zn * cc.y;                   // d lp_new / d mu
              gs[j] += (zn * zn - 1.f) * cc.y;      // d lp_new / d sigma
            }
          }
        }
      }
      __syncthreads();   // the previous chunk's column sums have been read
#pragma unroll
      for (int j = 0; j < kSampleChunk; ++j) { red[r][sl][j] = gm[j]; red[r][sl][kSampleChunk + j] = gs[j]; }
      __syncthreads();
      {
        // thread (r, q): q < 8 -> d/d mu of dimension c0 + q, q >= 8 -> d/d sigma of dimension c0 + q - 8
        const int q = sl, k = c0 + (q & (kSampleChunk - 1));
        if (ok && k < A) {
          float acc = 0.f;
#pragma unroll
          for (int l = 0; l < 16; ++l) acc += red[r][l][q];
          (q < kSampleChunk ? gmu_kl : gsig_kl)[static_cast<size_t>(row) * A + k] = acc * inv_s;
        }
      }
    }
    // ---- row scalars: log-prob of the policy sample and the KL estimate
    float logp = 0.f, klc = 0.f;
    if (ok)
      for (int k = sl; k < A; k += 16) { const float2 v = rowk[r][k]; logp += v.x; klc += v.y; }
#pragma unroll
    for (int of = 8; of > 0; of >>= 1) {
      logp += __shfl_xor_sync(0xffffffffu, logp, of);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, of);
      klc += __shfl_xor_sync(0xffffffffu, klc, of);
    }
    if (sl == 0 && ok) {
      logp_out[row] = logp;
      kl_out[row] = ac.reverse_kl ? dsum * inv_s - klc : dsum * inv_s + klc;
    }
    __syncthreads();   // cst / rowk / red are rewritten by the next row block
  }
}

// Q_i = softmax(logits_i) . support  and  d loss / d logits = -(w_path_i / rows) * p (s - Q)
__global__ void __launch_bounds__(256)
actor_q_kernel(const float* __restrict__ head_out, int ld, const float* __restrict__ zero_dist, HLConsts hl,
               const float* __restrict__ kl, int rows, float inv_rows, int kl_mode, float kl_bound,
               float* __restrict__ q_out, float* __restrict__ dlogits, __nv_bfloat16* __restrict__ dl_h, int ldh) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int B = hl.num_bins;
  RowSoftmax sx;
  row_softmax(head_out + static_cast<size_t>(row) * ld, zero_dist, hl.support, hl.bias_scale, B, lane, sx);
  const float w_path = (kl_mode == KL_CLIPPED) ? (kl[row] < kl_bound ? 1.f : 0.f) : 1.f;
  const float coef = -w_path * inv_rows;
  float* drow = dlogits + static_cast<size_t>(row) * ld;
#pragma unroll
  for (int j = 0; j < kMaxBinsPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < B) {
      const float o = coef * sx.p[j] * (hl.support[c] - sx.value);
      drow[c] = o;
      if (dl_h != nullptr) dl_h[static_cast<size_t>(row) * ldh + c] = __float2bfloat16_rn(o);
    }
  }
  if (lane == 0) q_out[row] = sx.value;
}

// d loss / d (mu, log_std) per row, scalar gradients for log_temp / log_lagrange, metrics.
// One warp per row (lane k owns action dimension k, +32 ...): 2048 rows give 2048 warps instead of 16 blocks of
// row-serial threads; metric partials and the bias-gradient column sums meet per block in shared memory.
__global__ void __launch_bounds__(256)
actor_grad_kernel(const float* __restrict__ out, const float* __restrict__ eps_pi, const float* __restrict__ act, int act_ld,
                  const float* __restrict__ dact, int dact_ld, const float* __restrict__ logp_in,
                  const float* __restrict__ kl_in, const float* __restrict__ q_in, const float* __restrict__ gmu_kl,
                  const float* __restrict__ gsig_kl, const float* __restrict__ actor_params, int rows, int A, float inv_rows,
                  ActorConsts ac, float* __restrict__ dout, float* __restrict__ actor_grads, float* __restrict__ metrics,
                  __nv_bfloat16* __restrict__ dout_h, int ldh, float* __restrict__ dbias) {
  REPPO_PDL_PROLOGUE();
  __shared__ float scol[8][2 * kGradMaxA];
  __shared__ float smet[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  const float alpha = expf(actor_params[0]), beta = expf(actor_params[1]);
  float m[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < 2 * A; j += 32) scol[warp][j] = 0.f;
  __syncwarp();
  if (row < rows) {
    const float logp = logp_in[row], kl = kl_in[row], q = q_in[row];
    float w_path, w_kl;
    if (ac.kl_mode == KL_CLIPPED) { w_path = kl < ac.kl_bound ? 1.f : 0.f; w_kl = 1.f - w_path; }
    else if (ac.kl_mode == KL_FULL) { w_path = 1.f; w_kl = 1.f; }
    else { w_path = 1.f; w_kl = 0.f; }
    const float g_logp = inv_rows * w_path * alpha;                // d L / d logp   (alpha is stop-gradient here)
    const float g_lpnew = -inv_rows * w_kl * beta * ac.reduce_kl;  // d L / d lp_new = - d L / d kl
    const float* o = out + static_cast<size_t>(row) * 2 * A;
    float* d = dout + static_cast<size_t>(row) * 2 * A;
    float abs_a = 0.f;
    for (int k = lane; k < A; k += 32) {
      const float mu = o[k], ex = expf(o[A + k]), sig = ex + ac.min_std;
      const float e = eps_pi[static_cast<size_t>(row) * A + k];
      const float a = act[static_cast<size_t>(row) * act_ld + k];
      const float da = dact[static_cast<size_t>(row) * dact_ld + k] * (1.f - a * a);  // through a = tanh(x)
      float dlp_mu, dlp_sig;
      const bool clamped = ac.clip_policy_sample && !(a >= -ac.action_clip && a <= ac.action_clip);
      if (!clamped) {
        dlp_mu = 2.f * a;
        dlp_sig = -1.f / sig + 2.f * a * e;
      } else {
        const float xc = atanhf(fminf(fmaxf(a, -ac.action_clip), ac.action_clip));
        const float zz = (xc - mu) / sig;
        dlp_mu = zz / sig;
        dlp_sig = (zz * zz - 1.f) / sig;
      }
      const float dmu = da + g_logp * dlp_mu + g_lpnew * gmu_kl[static_cast<size_t>(row) * A + k];
      const float dls = (da * e + g_logp * dlp_sig + g_lpnew * gsig_kl[static_cast<size_t>(row) * A + k]) * ex;
      d[k] = dmu;
      d[A + k] = dls;
      if (dout_h != nullptr) {
        dout_h[static_cast<size_t>(row) * ldh + k] = __float2bfloat16_rn(dmu);
        dout_h[static_cast<size_t>(row) * ldh + A + k] = __float2bfloat16_rn(dls);
      }
      scol[warp][k] = dmu;
      scol[warp][A + k] = dls;
      abs_a += fabsf(a);
    }
    abs_a = warp_sum(abs_a);
    m[0] = w_path * (alpha * logp - q) + w_kl * beta * ac.reduce_kl * kl;
    m[1] = kl;
    m[2] = -logp;
    m[3] = q;
    m[4] = abs_a / static_cast<float>(A);
    m[5] = ac.target_entropy - logp;
    m[6] = 1.f;
  }
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < 7; ++j) smet[warp][j] = m[j];
  __syncthreads();
  if (dbias != nullptr) {
    for (int j = threadIdx.x; j < 2 * A; j += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += scol[w][j];
      atomicAdd(dbias + j, t);
    }
  }
  if (threadIdx.x == 0) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < 8; ++w)
#pragma unroll
      for (int j = 0; j < 7; ++j) v[j] += smet[w][j];
    const float s_path = v[0] * inv_rows, s_kl = v[1] * inv_rows, s_ent = v[2] * inv_rows, s_q = v[3] * inv_rows;
    const float s_abs = v[4] * inv_rows, s_ent_t = v[5] * inv_rows, frac = v[6] * inv_rows;
    // entropy_loss = alpha * mean(sg(target_entropy - logp)); lagrangian_loss = -beta * mean(sg(kl - bound))
    const float ent_loss = alpha * s_ent_t;
    const float lag_loss = -beta * (s_kl - ac.kl_bound * frac);
    float total = s_path;
    if (ac.update_entropy) { total += ent_loss; atomicAdd(actor_grads + 0, ent_loss); }   // d/d log_temp = alpha * c
    if (ac.update_kl) { total += lag_loss; atomicAdd(actor_grads + 1, lag_loss); }        // d/d log_lagrange = -beta * c
    atomicAdd(metrics + M_ACTOR_LOSS, total);
    atomicAdd(metrics + M_POLICY_LOSS, s_path);
    atomicAdd(metrics + M_KL, s_kl);
    atomicAdd(metrics + M_ENTROPY, s_ent);
    atomicAdd(metrics + M_ENTROPY_LOSS, ent_loss);
    atomicAdd(metrics + M_LAGRANGIAN_LOSS, lag_loss);
    atomicAdd(metrics + M_ACTOR_Q_MEAN, s_q);
    atomicAdd(metrics + M_ABS_PRED_ACTION, s_abs);
    if (blockIdx.x == 0) { metrics[M_TEMPERATURE] = alpha; metrics[M_LAGRANGIAN] = beta; }
  }
}

// ---- launchers -----------------------------------------------------------------------------------
cudaError_t launch_critic_ce(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, const float* target,
                             const float* truncated, const int* idx, int rows, float inv_rows, int no_mask, float* dlogits,
                             float* metrics, void* dl_h, int ldh, float* dbias, float* dzero, cudaStream_t st) {
  if (hl.num_bins > 32 * kMaxBinsPerLane) return cudaErrorInvalidValue;
  const int want = (rows + kLossRowsPerBlock - 1) / kLossRowsPerBlock;
  const dim3 grid(want < kLossMaxBlocks ? want : kLossMaxBlocks);
#define REPPO_CE(NB) launch_k((critic_ce_kernel<NB>), grid, dim3(256), 0, st, head_out, ld, zero_dist, hl, target, truncated, idx, rows, inv_rows, no_mask, dlogits, metrics, static_cast<__nv_bfloat16*>(dl_h), ldh, dbias, dzero)
  if (hl.num_bins <= 64) REPPO_CE(2);
  else if (hl.num_bins <= 160) REPPO_CE(5);
  else REPPO_CE(8);
#undef REPPO_CE
  return cudaGetLastError();
}

cudaError_t launch_critic_aux(const float* pred, int P, int H, int reward_head, int mask_done, const float* next_emb,
                              const float* reward, const float* done, const float* truncated, const int* idx, int rows,
                              float inv_rows, int no_mask, float aux_mult, float* dpred, float* metrics, void* dp_h, int ldh,
                              float* dbias, cudaStream_t st) {
  const size_t smem = dbias ? 8 * static_cast<size_t>(P) * sizeof(float) : 0;
  if (smem > 48 * 1024) {   // per device, and cheap: set unconditionally
    cudaError_t e = cudaFuncSetAttribute(critic_aux_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2049 * 4);
    if (e != cudaSuccess) return e;
  }
  const int want = (rows + kLossRowsPerBlock - 1) / kLossRowsPerBlock;
  launch_k((critic_aux_kernel), dim3(want < kLossMaxBlocks ? want : kLossMaxBlocks), dim3(256), smem, st, pred, P, H, reward_head, mask_done, next_emb, reward, done, truncated, idx, rows, inv_rows, no_mask, aux_mult, dpred, metrics, static_cast<__nv_bfloat16*>(dp_h), ldh, dbias);
  return cudaGetLastError();
}

cudaError_t launch_actor_sample(const float* out, const float* out_old, const int* old_idx, const float* eps_pi,
                                const float* eps_kl, int rows, int A, int kl_samples, const ActorConsts& ac, float* act_out,
                                int act_ld, float* logp, float* kl, float* gmu, float* gsig, void* act_h, int act_ldh,
                                cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if (A > kGradMaxA) return cudaErrorInvalidValue;
  const int want = (rows + kSampleRows - 1) / kSampleRows;
  launch_k((actor_sample_kernel), dim3(want < kLossMaxBlocks ? want : kLossMaxBlocks), dim3(256), 0, st, out, out_old, old_idx, eps_pi, eps_kl, rows, A, kl_samples, ac, act_out, act_ld, logp, kl, gmu, gsig, static_cast<__nv_bfloat16*>(act_h), act_ldh);
  return cudaGetLastError();
}

cudaError_t launch_actor_q(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, const float* kl, int rows,
                           float inv_rows, int kl_mode, float kl_bound, float* q_out, float* dlogits, void* dl_h, int ldh,
                           cudaStream_t st) {
  if (hl.num_bins > 32 * kMaxBinsPerLane) return cudaErrorInvalidValue;
  launch_k((actor_q_kernel), dim3((rows + 7) / 8), dim3(256), 0, st, head_out, ld, zero_dist, hl, kl, rows, inv_rows, kl_mode, kl_bound, q_out, dlogits, static_cast<__nv_bfloat16*>(dl_h), ldh);
  return cudaGetLastError();
}

cudaError_t launch_actor_grad(const float* out, const float* eps_pi, const float* act, int act_ld, const float* dact,
                              int dact_ld, const float* logp, const float* kl, const float* q, const float* gmu,
                              const float* gsig, const float* actor_params, int rows, int A, float inv_rows,
                              const ActorConsts& ac, float* dout, float* actor_grads, float* metrics, void* dout_h, int ldh,
                              float* dbias, cudaStream_t st) {
  if (A > kGradMaxA) return cudaErrorInvalidValue;
  launch_k((actor_grad_kernel), dim3((rows + 7) / 8), dim3(256), 0, st, out, eps_pi, act, act_ld, dact, dact_ld, logp, kl, q, gmu, gsig, actor_params, rows, A, inv_rows, ac, dout, actor_grads, metrics, static_cast<__nv_bfloat16*>(dout_h), ldh, dbias);
  return cudaGetLastError();
}

// Critic.forward outputs (torch_models.py:278-285): logits = head + scale * zero_dist, value = softmax . support
__global__ void __launch_bounds__(256)
value_head_kernel(const float* __restrict__ head_out, int ld, const float* __restrict__ zero_dist, HLConsts hl, int rows,
                  float* __restrict__ value, float* __restrict__ logits) {
  REPPO_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  RowSoftmax sx;
  row_softmax(head_out + static_cast<size_t>(row) * ld, zero_dist, hl.support, hl.bias_scale, hl.num_bins, lane, sx);
  if (logits != nullptr) {
#pragma unroll
    for (int j = 0; j < kMaxBinsPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < hl.num_bins) logits[static_cast<size_t>(row) * hl.num_bins + c] = sx.lg[j];
    }
  }
  if (value != nullptr && lane == 0) value[row] = sx.value;
}

__global__ void mean_std_kernel(const float* __restrict__ out, int rows, int A, float min_std, float* __restrict__ mean,
                                float* __restrict__ std) {
  REPPO_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * A) return;
  const int r = i / A, k = i % A;
  if (mean != nullptr) mean[i] = out[static_cast<size_t>(r) * 2 * A + k];
  if (std != nullptr) std[i] = expf(out[static_cast<size_t>(r) * 2 * A + A + k]) + min_std;
}

cudaError_t launch_value_head(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, int rows,
                              float* value, float* logits, cudaStream_t st) {
  if (hl.num_bins > 32 * kMaxBinsPerLane) return cudaErrorInvalidValue;
  launch_k((value_head_kernel), dim3((rows + 7) / 8), dim3(256), 0, st, head_out, ld, zero_dist, hl, rows, value, logits);
  return cudaGetLastError();
}

cudaError_t launch_mean_std(const float* out, int rows, int A, float min_std, float* mean, float* std, cudaStream_t st) {
  launch_k((mean_std_kernel), dim3((rows * A + 255) / 256), dim3(256), 0, st, out, rows, A, min_std, mean, std);
  return cudaGetLastError();
}

}  // namespace reppo
