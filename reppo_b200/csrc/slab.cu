// Slab kernel: ONE persistent tcgen05 kernel runs a whole critic (or actor) chain of a REPPO minibatch step.
//
//   reference path: critic_loss_fn / actor_loss and their backward, src/jaxrl/reppo.py:474-622;
//                   update_critic / update_actor, src/torchrl/reppo.py:227-358; FCNN src/networks/jax_models.py:57-124.
//
// Why.  A step is a chain of ~25 dependent few-microsecond kernels per network; on B200 the chain, not the math
// (17 us of bf16 FLOPs per step), sets the step time.  Forward and input-gradient passes never mix minibatch rows, so
// a thread-block CLUSTER owns a slab of 128 rows and walks the whole chain inside one launch:
//   * grid = (C, ceil(rows / 128)), cluster = (C, 1, 1), C = hidden / 64: CTA `rank` owns the 64-wide output-column
//     tiles rank, rank + C, ... of every layer (UMMA 128 x 64 x 16, bf16 -> fp32 accumulators in TMEM);
//   * operands arrive by TMA (128 B swizzle) through a 4-stage mbarrier ring; warp 0 = producer, warp 1 = MMA issuer,
//     warps 2..9 = epilogue (tcgen05.ld -> shared-memory transpose -> row-contiguous global accesses);
//   * a layer's output leaves as a bf16 "twin" in global memory (the weight-gradient GEMMs need it there anyway) and
//     the next layer's TMA loads read it back from L2 after a cluster barrier (generic -> async proxy fence);
//   * row statistics of RMSNorm / LayerNorm (forward and backward) are exchanged through distributed shared memory;
//   * the losses (HL-Gauss cross-entropy, auxiliary MSE, tanh-Gaussian sampling + KL, Q, policy gradient) run as ROW
//     phases between GEMM steps: one warp per row, the slab's rows spread over every warp of the cluster.
// What stays outside: weight gradients (reduction over ALL rows), the gradient norm and Adam.
#include "slab.h"

#include "common.cuh"
#include "launch.h"
#include "loss_math.cuh"
#include "tc_ptx.cuh"

namespace reppo {
namespace {

using namespace tcx;
constexpr int kABytes = SB_M * SB_K * 2;            // 16 KB
constexpr int kBBytes = SB_N * SB_K * 2;            //  8 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kRingBytes = kSlabStages * kStageBytes;   // 96 KB; doubles as epilogue staging and row-phase scratch
constexpr int kBarOffset = kRingBytes;
constexpr int kStatOffset = kBarOffset + 128;
// part[2][2][128] | rowa[8][32] | rowb[8][32] | colpart[3][8][32] | metpart[8][4] | colc[3][64] | allstat[8][2][128]
constexpr int kOffRowA = 512, kOffRowB = 768, kOffColPart = 1024, kOffMetPart = 1792, kOffColC = 1824, kOffAllStat = 2016, kStatFloats = 2016 + 2048;
constexpr int kInfoOffset = kStatOffset + kStatFloats * 4;   // shared-memory copy of the program (minus tensor maps)
constexpr int kInfoBytes = ((kSlabMaxSteps * static_cast<int>(sizeof(SlabStepInfo)) + static_cast<int>(sizeof(SlabRowArgs)) + 15) / 16) * 16;
constexpr int kSlabSmem = kInfoOffset + kInfoBytes + 1024;
constexpr int kStgLd = 36;   // floats per staged row (32 + 4: conflict-free float4 rows)
static_assert(2 * 8 * 32 * kStgLd * 4 <= kRingBytes, "epilogue staging must fit in the ring");
static_assert((8 * 256 + 8 * 8 + 16) * 4 <= kRingBytes, "row-phase scratch must fit in the ring");

// ---- row phases ------------------------------------------------------------------------------------
// gw / nw: this warp's index among, and the number of, the cluster's epilogue warps; ew: index inside the CTA.
struct RowCtx {
  int m0, rows_here, M, gw, nw, ew, lane, rank;
  float* scratch;   // ring memory (idle during row phases)
};

// Row phases run two rows per warp pass with the rows' independent instruction streams interleaved (8 warps per
// SM cannot hide the load / SFU latency of one row's dependent chain).
__device__ void row_gather(const SlabRowArgs& a, const RowCtx& rc, bool actor) {
  for (int r = rc.gw; r < rc.rows_here; r += 2 * rc.nw) {
    int rows[2] = {rc.m0 + r, rc.m0 + r + rc.nw};
    const bool ok[2] = {true, r + rc.nw < rc.rows_here};
    long long src[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) src[q] = ok[q] ? (a.idx ? a.idx[rows[q]] : rows[q]) : 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
      const int row = rows[q];
      if (!actor) {
        for (int c = rc.lane; c < a.K0; c += 32) {
          const float v = c < a.Dc ? a.critic_obs[src[q] * a.Dc + c] : a.action[src[q] * a.A + (c - a.Dc)];
          a.x0_twin[static_cast<size_t>(row) * a.x0_ld + c] = __float2bfloat16_rn(v);
        }
      } else {
        for (int c = rc.lane; c < a.D; c += 32)
          a.xa0_twin[static_cast<size_t>(row) * a.xa0_ld + c] = __float2bfloat16_rn(a.obs[src[q] * a.D + c]);
        for (int c = rc.lane; c < a.Dc; c += 32)
          a.x0_twin[static_cast<size_t>(row) * a.x0_ld + c] = __float2bfloat16_rn(a.critic_obs[src[q] * a.Dc + c]);
      }
    }
  }
}

// HL-Gauss cross-entropy + d logits (critic_ce_kernel of losses.cu; lane owns columns lane + 32 j)
__device__ void row_ce(const SlabRowArgs& a, const RowCtx& rc) {
  const HLConsts& hl = a.hl;
  const int B = hl.num_bins, lane = rc.lane;
  float* scol = rc.scratch;              // [8][256]
  float* smet = rc.scratch + 8 * 256;    // [8][8]
  float vals[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float tmax = -INFINITY, tmin = INFINITY;
  float cs[kMaxBinsPerLane];
  // per-column constants of this lane (shared by every row): bin edges, support, logit offset
  float e_lo[kMaxBinsPerLane], e_hi[kMaxBinsPerLane], sup[kMaxBinsPerLane], zoff[kMaxBinsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxBinsPerLane; ++j) {
    const int c = lane + 32 * j;
    cs[j] = 0.f;
    e_lo[j] = (c < B) ? hl.edges[c] : 0.f;
    e_hi[j] = (c < B) ? hl.edges[c + 1] : 0.f;
    sup[j] = (c < B) ? hl.support[c] : 0.f;
    zoff[j] = (c < B) ? hl.bias_scale * a.zero_dist[c] : 0.f;
  }
  const float e_first = hl.edges[0], e_last = hl.edges[B];
  for (int r = rc.gw; r < rc.rows_here; r += 2 * rc.nw) {
    const int rows[2] = {rc.m0 + r, rc.m0 + r + rc.nw};
    const bool ok[2] = {true, r + rc.nw < rc.rows_here};
    long long src[2];
    float tgt[2], mask[2], lg[2][kMaxBinsPerLane], tp[2][kMaxBinsPerLane];
#pragma unroll
    for (int q = 0; q < 2; ++q) src[q] = ok[q] ? (a.idx ? a.idx[rows[q]] : rows[q]) : 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float* hr = a.head_out + static_cast<size_t>(ok[q] ? rows[q] : rows[0]) * a.ld_head;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        const int c = lane + 32 * j;
        lg[q][j] = (c < B) ? hr[c] + zoff[j] : -INFINITY;
      }
      tgt[q] = a.target[src[q]];
      mask[q] = a.no_mask ? 1.f : 1.f - a.truncated[src[q]];
    }
    float m[2], ssum[2], val[2], ce[2], sum_tp[2], zden[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      m[q] = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) m[q] = fmaxf(m[q], lg[q][j]);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) m[q] = warp_max(m[q]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      ssum[q] = 0.f; val[q] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        const float p = (lane + 32 * j < B) ? __expf(lg[q][j] - m[q]) : 0.f;
        ssum[q] += p;
        val[q] += p * sup[j];
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) { ssum[q] = warp_sum(ssum[q]); val[q] = warp_sum(val[q]); }
    const float inv_den = 1.f / hl.den;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float x = fminf(fmaxf(tgt[q], hl.vmin), hl.vmax);
      zden[q] = erff((e_last - x) * inv_den) - erff((e_first - x) * inv_den) + hl.z_eps;
      const float inv_z = 1.f / zden[q];
      const float lse = m[q] + logf(ssum[q]);
      ce[q] = 0.f; sum_tp[q] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        tp[q][j] = 0.f;
        if (lane + 32 * j < B) {
          tp[q][j] = (erff((e_hi[j] - x) * inv_den) - erff((e_lo[j] - x) * inv_den)) * inv_z;
          ce[q] -= tp[q][j] * (lg[q][j] - lse);
          sum_tp[q] += tp[q][j];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) { ce[q] = warp_sum(ce[q]); sum_tp[q] = warp_sum(sum_tp[q]); }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
      const float coef = mask[q] * a.inv_rows, inv_s = 1.f / ssum[q];
      const float value = val[q] * inv_s;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < B) {
          const float p = __expf(lg[q][j] - m[q]) * inv_s;
          const float o = coef * (sum_tp[q] * p - tp[q][j]);
          cs[j] += o;
          a.dl_twin[static_cast<size_t>(rows[q]) * a.dl_ld + c] = __float2bfloat16_rn(o);
        }
      }
      vals[0] += mask[q] * ce[q] * a.inv_rows;
      vals[1] += ce[q] * a.inv_rows;
      vals[2] += (value - tgt[q]) * (value - tgt[q]) * a.inv_rows;
      vals[3] += value * a.inv_rows;
      vals[4] += tgt[q] * a.inv_rows;
      tmax = fmaxf(tmax, tgt[q]); tmin = fminf(tmin, tgt[q]);
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxBinsPerLane; ++j) scol[rc.ew * 256 + lane + 32 * j] = cs[j];
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) smet[rc.ew * 8 + k] = vals[k];
    smet[rc.ew * 8 + 5] = tmax; smet[rc.ew * 8 + 6] = tmin;
  }
  epi_bar();
  if (rc.rank * 8 < rc.rows_here) {   // this CTA had rows
    const int t = rc.ew * 32 + lane;
    if (t < B) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += scol[w * 256 + t];
      if (a.g_bias_head != nullptr) atomicAdd(a.g_bias_head + t, s);
      if (a.g_zero != nullptr) atomicAdd(a.g_zero + t, s * hl.bias_scale);
    }
    if (t == 255) {
      float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, mx = -INFINITY, mn = INFINITY;
      for (int w = 0; w < 8; ++w) {
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] += smet[w * 8 + k];
        mx = fmaxf(mx, smet[w * 8 + 5]); mn = fminf(mn, smet[w * 8 + 6]);
      }
      atomicAdd(a.metrics + M_CRITIC_LOSS, v[0]);
      atomicAdd(a.metrics + M_CE_MEAN, v[1]);
      atomicAdd(a.metrics + M_VALUE_LOSS, v[2]);
      atomicAdd(a.metrics + M_Q_MEAN, v[3]);
      atomicAdd(a.metrics + M_TARGET_MEAN, v[4]);
      if (mx > -INFINITY) { atomic_max_float(a.metrics + M_TARGET_MAX, mx); atomic_min_float(a.metrics + M_TARGET_MIN, mn); }
    }
  }
  epi_bar();
}

// tanh-Gaussian sample, log-prob, forward KL to the behaviour policy and its gradient (actor_sample_kernel of
// losses.cu).  A warp takes two rows per pass, one per half-warp; inside a half, lane sl owns KL sample sl (+16 ...)
// and, for the policy sample itself, action dimension sl (+16).
__device__ void row_sample(const SlabRowArgs& a, const RowCtx& rc) {
  const int A = a.A, hw = rc.lane >> 4, sl = rc.lane & 15;
  const ActorConsts& ac = a.ac;
  const float inv_s = 1.f / static_cast<float>(a.kl_samples);
  for (int r = rc.gw; r < rc.rows_here; r += 2 * rc.nw) {
    const int rl = r + hw * rc.nw;
    const bool ok = rl < rc.rows_here;
    const int row = rc.m0 + (ok ? rl : r);
    const float* o = a.actor_out + static_cast<size_t>(row) * 2 * A;
    const float* oo = a.old_out + static_cast<size_t>(a.old_idx ? a.old_idx[row] : row) * 2 * A;
    float logp = 0.f, lp_old = 0.f, lp_new = 0.f;
    for (int k = sl; k < A; k += 16) {
      const float mu = o[k], sig = __expf(o[A + k]) + ac.min_std;
      const float e = a.eps_pi[static_cast<size_t>(row) * A + k];
      const float x = mu + sig * e;
      const float act = tanhf(x);
      if (ok) {
        a.x0a[static_cast<size_t>(row) * a.K0 + a.Dc + k] = act;
        a.x0_twin[static_cast<size_t>(row) * a.x0_ld + a.Dc + k] = __float2bfloat16_rn(act);
      }
      const float log_sig = logf(sig);
      if (ac.clip_policy_sample) {
        const float xc = atanhf(fminf(fmaxf(act, -ac.action_clip), ac.action_clip));
        const float zz = (xc - mu) / sig;
        logp += -0.5f * zz * zz - log_sig - kHalfLog2Pi - fldj_(xc);
      } else {
        logp += -0.5f * e * e - log_sig - kHalfLog2Pi - fldj_(x);
      }
    }
    for (int k = 0; k < A; ++k) {
      const float mu = o[k], sig = __expf(o[A + k]) + ac.min_std, log_sig = logf(sig), inv_sig = 1.f / sig;
      const float mu_o = oo[k], sig_o = __expf(oo[A + k]) + ac.min_std, log_sig_o = logf(sig_o);
      float gm = 0.f, gs = 0.f;
      for (int s = sl; s < a.kl_samples; s += 16) {
        const float eo = a.eps_kl[(static_cast<size_t>(s) * rc.M + row) * A + k];
        const float xo = mu_o + sig_o * eo;
        const float u = atanhf(fminf(fmaxf(tanhf(xo), -ac.action_clip), ac.action_clip));
        const float fu = fldj_(u);
        if (ac.old_logp_at_clipped) {
          const float zo = (u - mu_o) / sig_o;
          lp_old += -0.5f * zo * zo - log_sig_o - kHalfLog2Pi - fu;
        } else {
          lp_old += -0.5f * eo * eo - log_sig_o - kHalfLog2Pi - fldj_(xo);
        }
        const float zn = (u - mu) * inv_sig;
        lp_new += -0.5f * zn * zn - log_sig - kHalfLog2Pi - fu;
        gm += zn * inv_sig;
        gs += (zn * zn - 1.f) * inv_sig;
      }
#pragma unroll
      for (int of = 8; of > 0; of >>= 1) { gm += __shfl_xor_sync(0xffffffffu, gm, of); gs += __shfl_xor_sync(0xffffffffu, gs, of); }
      if (sl == 0 && ok) {
        a.gmu[static_cast<size_t>(row) * A + k] = gm * inv_s;
        a.gsig[static_cast<size_t>(row) * A + k] = gs * inv_s;
      }
    }
#pragma unroll
    for (int of = 8; of > 0; of >>= 1) {
      logp += __shfl_xor_sync(0xffffffffu, logp, of);
      lp_old += __shfl_xor_sync(0xffffffffu, lp_old, of);
      lp_new += __shfl_xor_sync(0xffffffffu, lp_new, of);
    }
    if (sl == 0 && ok) {
      a.logp[row] = logp;
      a.kl[row] = (lp_old - lp_new) * inv_s;
    }
  }
}

// Q = softmax(logits) . support and d loss / d logits of the pathwise term (actor_q_kernel of losses.cu)
__device__ void row_q(const SlabRowArgs& a, const RowCtx& rc) {
  const HLConsts& hl = a.hl;
  const int B = hl.num_bins, lane = rc.lane;
  float sup[kMaxBinsPerLane], zoff[kMaxBinsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxBinsPerLane; ++j) {
    const int c = lane + 32 * j;
    sup[j] = (c < B) ? hl.support[c] : 0.f;
    zoff[j] = (c < B) ? hl.bias_scale * a.zero_dist[c] : 0.f;
  }
  for (int r = rc.gw; r < rc.rows_here; r += 2 * rc.nw) {
    const int rows[2] = {rc.m0 + r, rc.m0 + r + rc.nw};
    const bool ok[2] = {true, r + rc.nw < rc.rows_here};
    float lg[2][kMaxBinsPerLane], klv[2], m[2], ssum[2], val[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int row = ok[q] ? rows[q] : rows[0];
      const float* hr = a.head_out + static_cast<size_t>(row) * a.ld_head;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) lg[q][j] = (lane + 32 * j < B) ? hr[lane + 32 * j] + zoff[j] : -INFINITY;
      klv[q] = a.kl[row];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      m[q] = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) m[q] = fmaxf(m[q], lg[q][j]);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) m[q] = warp_max(m[q]);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      ssum[q] = 0.f; val[q] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        lg[q][j] = (lane + 32 * j < B) ? __expf(lg[q][j] - m[q]) : 0.f;   // un-normalised probabilities from here on
        ssum[q] += lg[q][j];
        val[q] += lg[q][j] * sup[j];
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) { ssum[q] = warp_sum(ssum[q]); val[q] = warp_sum(val[q]); }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
      const float inv_s = 1.f / ssum[q], value = val[q] * inv_s;
      const float w_path = (a.ac.kl_mode == KL_CLIPPED) ? (klv[q] < a.ac.kl_bound ? 1.f : 0.f) : 1.f;
      const float coef = -w_path * a.inv_rows * inv_s;
#pragma unroll
      for (int j = 0; j < kMaxBinsPerLane; ++j) {
        const int c = lane + 32 * j;
        if (c < B) a.dl_twin[static_cast<size_t>(rows[q]) * a.dl_ld + c] = __float2bfloat16_rn(coef * lg[q][j] * (sup[j] - value));
      }
      if (lane == 0) a.q[rows[q]] = value;
    }
  }
}

// d loss / d (mu, log_std), temperature / Lagrange gradients, metrics (actor_grad_kernel of losses.cu)
__device__ void row_actor_grad(const SlabRowArgs& a, const RowCtx& rc) {
  const int A = a.A, k = rc.lane;
  const ActorConsts& ac = a.ac;
  float* scol = rc.scratch;             // [8][64]
  float* smet = rc.scratch + 8 * 64;    // [8][8]
  const float alpha = expf(a.actor_params[0]), beta = expf(a.actor_params[1]);
  float m_path = 0.f, m_kl = 0.f, m_ent = 0.f, m_q = 0.f, m_abs = 0.f, m_ent_t = 0.f, m_cnt = 0.f;
  float c_mu = 0.f, c_sig = 0.f;
  for (int r = rc.gw; r < rc.rows_here; r += rc.nw) {
    const int row = rc.m0 + r;
    const float logp = a.logp[row], kl = a.kl[row], q = a.q[row];
    float w_path, w_kl;
    if (ac.kl_mode == KL_CLIPPED) { w_path = kl < ac.kl_bound ? 1.f : 0.f; w_kl = 1.f - w_path; }
    else if (ac.kl_mode == KL_FULL) { w_path = 1.f; w_kl = 1.f; }
    else { w_path = 1.f; w_kl = 0.f; }
    const float g_logp = a.inv_rows * w_path * alpha;
    const float g_lpnew = -a.inv_rows * w_kl * beta * ac.reduce_kl;
    float abs_a = 0.f;
    if (k < A) {
      const float* o = a.actor_out + static_cast<size_t>(row) * 2 * A;
      const float mu = o[k], ex = expf(o[A + k]), sig = ex + ac.min_std;
      const float e = a.eps_pi[static_cast<size_t>(row) * A + k];
      const float act = a.x0a[static_cast<size_t>(row) * a.K0 + a.Dc + k];
      const float da = a.dx0[static_cast<size_t>(row) * a.K0 + a.Dc + k] * (1.f - act * act);
      float dlp_mu, dlp_sig;
      const bool clamped = ac.clip_policy_sample && !(act >= -ac.action_clip && act <= ac.action_clip);
      if (!clamped) {
        dlp_mu = 2.f * act;
        dlp_sig = -1.f / sig + 2.f * act * e;
      } else {
        const float xc = atanhf(fminf(fmaxf(act, -ac.action_clip), ac.action_clip));
        const float zz = (xc - mu) / sig;
        dlp_mu = zz / sig;
        dlp_sig = (zz * zz - 1.f) / sig;
      }
      const float dmu = da + g_logp * dlp_mu + g_lpnew * a.gmu[static_cast<size_t>(row) * A + k];
      const float dsig = (da * e + g_logp * dlp_sig + g_lpnew * a.gsig[static_cast<size_t>(row) * A + k]) * ex;
      a.dout_twin[static_cast<size_t>(row) * a.dout_ld + k] = __float2bfloat16_rn(dmu);
      a.dout_twin[static_cast<size_t>(row) * a.dout_ld + A + k] = __float2bfloat16_rn(dsig);
      c_mu += dmu; c_sig += dsig;
      abs_a = fabsf(act);
    }
    abs_a = warp_sum(abs_a);
    m_path += w_path * (alpha * logp - q) + w_kl * beta * ac.reduce_kl * kl;
    m_kl += kl;
    m_ent += -logp;
    m_q += q;
    m_abs += abs_a / static_cast<float>(A);
    m_ent_t += ac.target_entropy - logp;
    m_cnt += 1.f;
  }
  scol[rc.ew * 64 + k] = c_mu;
  scol[rc.ew * 64 + 32 + k] = c_sig;
  if (k == 0) {
    float* m = smet + rc.ew * 8;
    m[0] = m_path; m[1] = m_kl; m[2] = m_ent; m[3] = m_q; m[4] = m_abs; m[5] = m_ent_t; m[6] = m_cnt;
  }
  epi_bar();
  if (rc.rank * 8 < rc.rows_here) {
    const int t = rc.ew * 32 + k;
    if (t < 2 * A && a.g_bias_actor_last != nullptr) {
      const int slot = t < A ? t : 32 + (t - A);
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += scol[w * 64 + slot];
      atomicAdd(a.g_bias_actor_last + t, s);
    }
    if (t == 255) {
      float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int j = 0; j < 7; ++j) v[j] += smet[w * 8 + j];
      const float inv = a.inv_rows;
      const float s_path = v[0] * inv, s_kl = v[1] * inv, s_ent = v[2] * inv, s_q = v[3] * inv, s_abs = v[4] * inv;
      const float s_ent_t = v[5] * inv, frac = v[6] * inv;
      const float ent_loss = alpha * s_ent_t;
      const float lag_loss = -beta * (s_kl - ac.kl_bound * frac);
      float total = s_path;
      if (ac.update_entropy) { total += ent_loss; atomicAdd(a.actor_grads + 0, ent_loss); }
      if (ac.update_kl) { total += lag_loss; atomicAdd(a.actor_grads + 1, lag_loss); }
      atomicAdd(a.metrics + M_ACTOR_LOSS, total);
      atomicAdd(a.metrics + M_POLICY_LOSS, s_path);
      atomicAdd(a.metrics + M_KL, s_kl);
      atomicAdd(a.metrics + M_ENTROPY, s_ent);
      atomicAdd(a.metrics + M_ENTROPY_LOSS, ent_loss);
      atomicAdd(a.metrics + M_LAGRANGIAN_LOSS, lag_loss);
      atomicAdd(a.metrics + M_ACTOR_Q_MEAN, s_q);
      atomicAdd(a.metrics + M_ABS_PRED_ACTION, s_abs);
      if (blockIdx.x == 0 && blockIdx.y == 0) { a.metrics[M_TEMPERATURE] = alpha; a.metrics[M_LAGRANGIAN] = beta; }
    }
  }
  epi_bar();
}

// ---- the kernel --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSlabThreads, 1) slab_kernel(const __grid_constant__ SlabProgram prog) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
  uint64_t* empty_bar = full_bar + kSlabStages;
  uint64_t* tmem_full_bar = empty_bar + kSlabStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* stats = reinterpret_cast<float*>(smem + kStatOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x, ncta = gridDim.x;
  const int m0 = blockIdx.y * SB_M, M = prog.rows;
  const int rows_here = min(SB_M, M - m0);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(prog.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSlabStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The program lives in the kernel-parameter constant bank; cluster barriers invalidate the constant caches along
  // with L1, so every descriptor field read after a barrier would be an L2 round trip: keep a shared-memory copy.
  SlabStepInfo* sinfo = reinterpret_cast<SlabStepInfo*>(smem + kInfoOffset);
  SlabRowArgs* rargs_s = reinterpret_cast<SlabRowArgs*>(sinfo + kSlabMaxSteps);
  {
    constexpr int kW = sizeof(SlabStepInfo) / 4;
    uint32_t* dst = reinterpret_cast<uint32_t*>(sinfo);
    for (int w = threadIdx.x; w < prog.num_steps * kW; w += kSlabThreads)
      dst[w] = reinterpret_cast<const uint32_t*>(&prog.steps[w / kW].f)[w % kW];
    uint32_t* dr = reinterpret_cast<uint32_t*>(rargs_s);
    for (int w = threadIdx.x; w < static_cast<int>(sizeof(SlabRowArgs) / 4); w += kSlabThreads)
      dr[w] = reinterpret_cast<const uint32_t*>(&prog.r)[w];
  }
  __syncthreads();
  const SlabRowArgs& rargs = *rargs_s;
  const int num_steps = prog.num_steps;
  long long* const timing = prog.timing;
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above overlapped the previous kernel's tail.  The dependents are released at the LAST step, not here:
  // this kernel runs for tens of microseconds and early-launched dependents would only park on its SMs.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // epilogue-warp coordinates (warps 2..9): TMEM lane quarter = warp % 4; the two warps of a quarter split the 64 columns
  const int ew = warp - 2, half = (ew >> 2) & 1, lane_base = (warp & 3) * 32, cbase = half * 32;
  const int rrow = lane_base + lane;            // row of the slab this thread owns in the row-per-thread layout
  const int cl = (lane & 7) * 4, rg = lane >> 3; // coalesced layout: 8 lanes sweep 32 columns, 4 rows per pass
  float* stg = reinterpret_cast<float*>(smem) + (ew & 7) * 32 * kStgLd;
  float* stg2 = stg + 8 * 32 * kStgLd;   // second staging plane (norm backward keeps dy and zh)
  float* part = stats;
  float* rowa = stats + kOffRowA + (ew & 7) * 32;
  float* rowb = stats + kOffRowB + (ew & 7) * 32;
  float* colpart = stats + kOffColPart;
  float* metpart = stats + kOffMetPart;
  float* allstat = stats + kOffAllStat;   // [source rank][2][128]: every CTA's row statistics, pushed by the owners
  float* colc = stats + kOffColC;   // bias | gamma | beta of this CTA's first tile (fetched while the MMAs run)

  RowCtx rc;
  rc.m0 = m0; rc.rows_here = rows_here; rc.M = M; rc.gw = rank * 8 + ew; rc.nw = ncta * 8; rc.ew = ew; rc.lane = lane; rc.rank = rank;
  rc.scratch = reinterpret_cast<float*>(smem);

  uint32_t it = 0, tf = 0;   // ring iterations and accumulator hand-overs so far (uniform over the CTA)
  const bool timed = timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define SLAB_STAMP(k) do { if (timed) timing[64 + s * 12 + (k)] = clock64(); } while (0)
  for (int s = 0; s < num_steps; ++s) {
    const SlabStepInfo& st = sinfo[s];
    const SlabStep& maps = prog.steps[s];
    handoff_sync();
    if (timed && threadIdx.x == 64) timing[s] = clock64();
    if (s == num_steps - 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (st.kind == SLAB_ROW) {
      if (warp >= 2) {
        if (st.row == ROW_GATHER_CRITIC) row_gather(rargs, rc, false);
        else if (st.row == ROW_GATHER_ACTOR) row_gather(rargs, rc, true);
        else if (st.row == ROW_CE) row_ce(rargs, rc);
        else if (st.row == ROW_SAMPLE) row_sample(rargs, rc);
        else if (st.row == ROW_Q) row_q(rargs, rc);
        else row_actor_grad(rargs, rc);
      }
      continue;
    }
    const int N = st.N, epi = st.epi;
    const int ntiles = (N + SB_N - 1) / SB_N;
    const int my_tiles = rank < ntiles ? (ntiles - rank + ncta - 1) / ncta : 0;
    const int nkb = (st.K + SB_K - 1) / SB_K;
    const bool is_norm = (epi == EPI_FWD_NORM || epi == EPI_BWD_NORM);
    float4 pre[8];              // epilogue operands fetched from global memory during the main loop
    float prow[8], prow2[8];

    if (warp == 0) {
      if (lane == 0) {
        // ===== TMA producer =====
        asm volatile("fence.proxy.async;" ::: "memory");
        uint32_t i = it;
        for (int t = 0; t < my_tiles; ++t) {
          const int n0 = (rank + t * ncta) * SB_N;
          for (int kb = 0; kb < nkb; ++kb, ++i) {
            const int sg = i % kSlabStages;
            const uint32_t ph = (i / kSlabStages) & 1;
            mbar_wait(&empty_bar[sg], ph ^ 1);
            uint8_t* sa = smem + sg * kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_expect_tx(&full_bar[sg], kStageBytes);
            tma_load_2d(&maps.ta, &full_bar[sg], sa, kb * SB_K, m0);
            if (!st.b_mn) tma_load_2d(&maps.tb, &full_bar[sg], sb, kb * SB_K, n0);   // [n][k] tile, coordinates {k, n}
            else tma_load_2d(&maps.tb, &full_bar[sg], sb, n0, kb * SB_K);            // [k][n] tile, coordinates {n, k}
          }
        }
        SLAB_STAMP(1);
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(st.b_mn != 0);
        uint32_t i = it;
        for (int t = 0; t < my_tiles; ++t) {
          for (int kb = 0; kb < nkb; ++kb, ++i) {
            const int sg = i % kSlabStages;
            const uint32_t ph = (i / kSlabStages) & 1;
            mbar_wait(&full_bar[sg], ph);
            tc_fence_after();
            if (t == 0 && kb == 0) SLAB_STAMP(2);
            const uint32_t sa = smem_u32(smem + sg * kStageBytes);
            const uint32_t sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < SB_K / SB_UMMA_K; ++k) {
              const uint64_t da = make_smem_desc(sa + k * (SB_UMMA_K * 2), 16, 1024);
              const uint64_t db = st.b_mn ? make_smem_desc(sb + k * (SB_UMMA_K * 128), SB_K * 128, 1024)
                                          : make_smem_desc(sb + k * (SB_UMMA_K * 2), 16, 1024);
              tc_mma_bf16(tmem_base + t * SB_N, da, db, idesc, (kb > 0 || k > 0 || st.accumulate) ? 1u : 0u);
            }
            tc_commit(&empty_bar[sg]);
          }
        }
        if (my_tiles > 0 && epi != EPI_NONE) tc_commit(tmem_full_bar);
        SLAB_STAMP(3);
      }
      __syncwarp();
    } else if (epi != EPI_NONE) {
      // ===== epilogue, part 0: what the epilogue needs from global memory is requested while the MMAs run =====
      // (coalesced layout: this thread's rows rg + 4 i, columns col0 .. col0 + 3 of the CTA's first tile)
      const int col0 = rank * SB_N + cbase + cl;
      if (my_tiles > 0) {
        const int te = ew * 32 + lane;
        if (te < SB_N) {
          const int c = rank * SB_N + te;
          colc[te] = (st.bias != nullptr && c < N) ? st.bias[c] : 0.f;
          colc[64 + te] = (is_norm && st.gamma != nullptr && c < N) ? st.gamma[c] : 1.f;
          colc[128 + te] = (is_norm && st.beta != nullptr && c < N) ? st.beta[c] : 0.f;
        }
      }
      if (my_tiles > 0 && (epi == EPI_BWD_NORM || epi == EPI_BWD_ACT)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = m0 + lane_base + rg + 4 * i;
          pre[i] = make_float4(0.f, 0.f, 0.f, 0.f); prow[i] = 0.f; prow2[i] = 0.f;
          if (row < M && col0 < N) {
            pre[i] = *reinterpret_cast<const float4*>(st.out + static_cast<size_t>(row) * st.ld_out + col0);   // saved z
            if (epi == EPI_BWD_NORM) { prow[i] = st.rstd[row]; if (st.norm == NORM_LAYER) prow2[i] = st.mean[row]; }
          }
        }
      } else if (my_tiles > 0 && epi == EPI_FWD_AUX) {
        const SlabRowArgs& a = rargs;
        int srcs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = m0 + lane_base + rg + 4 * i;
          srcs[i] = (row < M) ? (a.idx ? a.idx[row] : row) : -1;
        }
        const int off = a.reward_head ? 1 : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pre[i] = make_float4(0.f, 0.f, 0.f, 0.f); prow[i] = 0.f; prow2[i] = 0.f;
          if (srcs[i] >= 0) {
            const size_t sr = static_cast<size_t>(srcs[i]);
            prow[i] = a.no_mask ? 1.f : 1.f - a.truncated[sr];
            prow2[i] = a.mask_done ? 1.f - a.done[sr] : 1.f;
            float tg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int cc = col0 + j;
              tg[j] = (cc < N) ? ((a.reward_head && cc == 0) ? a.reward[sr] : a.next_emb[sr * a.H + (cc - off)]) : 0.f;
            }
            pre[i] = make_float4(tg[0], tg[1], tg[2], tg[3]);
          }
        }
      }
      // ===== epilogue, part 1: accumulators -> staging; row statistics of this CTA's columns =====
      if (my_tiles > 0) {
        epi_bar();   // colc is complete
        mbar_wait(tmem_full_bar, tf & 1);
        tc_fence_after();
        if (threadIdx.x == 64) SLAB_STAMP(4);
      }
      if (my_tiles > 0 && is_norm) {
        // one tile per CTA (N <= 64 * cluster size for normalised layers)
        const int n0 = rank * SB_N;
        float s1 = 0.f, s2 = 0.f;
        {
          float v[32];
          tc_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + cbase, v);
          if (epi == EPI_FWD_NORM) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = n0 + cbase + j;
              v[j] = (col < N) ? v[j] + colc[cbase + j] : 0.f;
              s1 += v[j];
              s2 += v[j] * v[j];
            }
          }
          float4* dst = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
        if (epi == EPI_FWD_NORM) {
          part[(0 * 2 + half) * 128 + rrow] = s1;
          part[(1 * 2 + half) * 128 + rrow] = s2;
        } else {
          // BWD_NORM pass A (coalesced): dy = d * swish'(y) and zh replace d in the staging; partial sums of dzh and dzh * zh
          float g[4], be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) { g[j] = colc[64 + cbase + cl + j]; be[j] = colc[128 + cbase + cl + j]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i;
            const float4 d4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, zz[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
            float dy[4], zh[4], a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              zh[j] = (zz[j] - prow2[i]) * prow[i];
              dy[j] = dd[j] * fswish_grad(zh[j] * g[j] + be[j]);   // rows past the end: d = 0 (TMA zero fill), z = 0
              const float dzh = dy[j] * g[j];
              a1 += dzh; a2 += dzh * zh[j];
            }
            *reinterpret_cast<float4*>(stg + r * kStgLd + cl) = make_float4(dy[0], dy[1], dy[2], dy[3]);
            *reinterpret_cast<float4*>(stg2 + r * kStgLd + cl) = make_float4(zh[0], zh[1], zh[2], zh[3]);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
            if ((lane & 7) == 0) {
              part[(0 * 2 + half) * 128 + lane_base + r] = a1;
              part[(1 * 2 + half) * 128 + lane_base + r] = a2;
            }
          }
        }
        epi_bar();
        if (half == 0) {
          // push this CTA's per-row partial statistics into every peer (remote stores: fire and forget; the cluster
          // barrier below publishes them) -- pulling them after the barrier costs one remote round trip per peer
          const float u1 = part[0 * 128 + rrow] + part[1 * 128 + rrow];
          const float u2 = part[2 * 128 + rrow] + part[3 * 128 + rrow];
          for (int r = 0; r < ncta; ++r) {
            st_dsmem_f32(allstat + (rank * 2 + 0) * 128 + rrow, r, u1);
            st_dsmem_f32(allstat + (rank * 2 + 1) * 128 + rrow, r, u2);
          }
        }
      }
      tc_fence_before();
      if (threadIdx.x == 64) SLAB_STAMP(5);
    }
    if (is_norm) cluster_sync_all();   // every CTA's row statistics are published
    if (threadIdx.x == 64) SLAB_STAMP(6);

    if (warp >= 2 && epi != EPI_NONE && my_tiles > 0) {
      // ===== epilogue, part 2 =====
      for (int t = 0; t < my_tiles; ++t) {
        const int n0 = (rank + t * ncta) * SB_N;
        const int col = n0 + cbase + cl;
        if (is_norm) {
          // totals over the cluster (pushed into this CTA's shared memory by the owners)
          const int nsrc = min(ncta, ntiles);   // CTAs that own a tile of this layer
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (r < nsrc) { t1 += allstat[(r * 2 + 0) * 128 + rrow]; t2 += allstat[(r * 2 + 1) * 128 + rrow]; }
          }
          if (threadIdx.x == 64) SLAB_STAMP(8);
          const float inv_n = 1.f / static_cast<float>(N);
          if (epi == EPI_FWD_NORM) {
            float mean = 0.f, rstd;
            if (st.norm == NORM_RMS) {
              rstd = rsqrtf(t2 * inv_n + st.eps);
            } else {
              mean = t1 * inv_n;
              rstd = rsqrtf(t2 * inv_n - mean * mean + st.eps);   // flax use_fast_variance
            }
            rowa[lane] = rstd; rowb[lane] = mean;
            if (rank == 0 && half == 0 && m0 + rrow < M) {
              st.rstd[m0 + rrow] = rstd;
              if (st.norm == NORM_LAYER) st.mean[m0 + rrow] = mean;
            }
          } else {
            rowa[lane] = (st.norm == NORM_LAYER) ? t1 * inv_n : 0.f;   // mean(dzh)
            rowb[lane] = t2 * inv_n;                                   // mean(dzh * zh)
          }
          __syncwarp();
        } else {
          // plain tiles: accumulators (+ bias) -> staging
          float v[32];
          tc_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + t * SB_N + cbase, v);
          if (threadIdx.x == 64) SLAB_STAMP(8);
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += colc[cbase + j];
          } else if (st.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { const int c = n0 + cbase + j; v[j] += (c < N) ? st.bias[c] : 0.f; }
          }
          __syncwarp();   // the previous tile's readers are done with the staging
          float4* dst = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
        }
        if (threadIdx.x == 64) SLAB_STAMP(9);
        float* const out_p = st.out;
        __nv_bfloat16* const twin_p = st.twin;
        const int ld_out = st.ld_out, ld_twin = st.ld_twin;
        const bool full4 = (col + 3 < N);
        float pg[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f}, pd[4] = {0.f, 0.f, 0.f, 0.f};
        float mv0 = 0.f, mv1 = 0.f, mv2 = 0.f;
        if (epi == EPI_FWD_NORM) {
          float g[4], be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) { g[j] = colc[64 + cbase + cl + j]; be[j] = colc[128 + cbase + cl + j]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
            const float rr = rowa[r], mm = rowb[r];
            const float y0 = fswish((x4.x - mm) * rr * g[0] + be[0]), y1 = fswish((x4.y - mm) * rr * g[1] + be[1]);
            const float y2 = fswish((x4.z - mm) * rr * g[2] + be[2]), y3 = fswish((x4.w - mm) * rr * g[3] + be[3]);
            if (row < M) {
              *reinterpret_cast<float4*>(out_p + static_cast<size_t>(row) * ld_out + col) = x4;
              store_bf16x4(twin_p + static_cast<size_t>(row) * ld_twin + col, y0, y1, y2, y3);
            }
          }
        } else if (epi == EPI_FWD_ACT) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
            const float y0 = fswish(x4.x), y1 = fswish(x4.y), y2 = fswish(x4.z), y3 = fswish(x4.w);
            if (row < M && col < N) {
              *reinterpret_cast<float4*>(out_p + static_cast<size_t>(row) * ld_out + col) = x4;
              store_bf16x4(twin_p + static_cast<size_t>(row) * ld_twin + col, y0, y1, y2, y3);
            }
          }
        } else if (epi == EPI_FWD_OUT || epi == EPI_BWD_OUT) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
            const float x[4] = {x4.x, x4.y, x4.z, x4.w};
            float* o = out_p + static_cast<size_t>(row) * ld_out + col;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (row < M && col + j < N) o[j] = x[j];
          }
        } else if (epi == EPI_FWD_AUX) {
          const SlabRowArgs& a = rargs;
          const int off = a.reward_head ? 1 : 0;
          const float inv_P = 1.f / static_cast<float>(a.P);
          const float gscale = 2.f * a.aux_mult * a.inv_rows * inv_P;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const bool okr = row < M && col < N;
            const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
            const float x[4] = {x4.x, x4.y, x4.z, x4.w};
            float mask = prow[i], nd = prow2[i];
            float tg[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
            if (t > 0 && okr) {   // extra tiles (P > 64 * cluster size: the reward column of a full-width head): loaded here
              const size_t sr = static_cast<size_t>(a.idx ? a.idx[row] : row);
              mask = a.no_mask ? 1.f : 1.f - a.truncated[sr];
              nd = a.mask_done ? 1.f - a.done[sr] : 1.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cc = col + j;
                tg[j] = (cc < N) ? ((a.reward_head && cc == 0) ? a.reward[sr] : a.next_emb[sr * a.H + (cc - off)]) : 0.f;
              }
            }
            const float gg = gscale * mask * nd;
            float sq = 0.f, dp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (okr && col + j < N) {
                if (a.reward_head && col + j == 0) mv2 += tg[j] * a.inv_rows;
                const float d = x[j] - tg[j];
                sq += d * d;
                dp[j] = gg * d;
                pd[j] += dp[j];
              }
            }
            __nv_bfloat16* tw = twin_p + static_cast<size_t>(row) * ld_twin + col;
            if (okr) {
              if (full4) store_bf16x4(tw, dp[0], dp[1], dp[2], dp[3]);
              else
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col + j < N) tw[j] = __float2bfloat16_rn(dp[j]);
            }
            const float aux = nd * sq * inv_P;
            mv0 += mask * a.aux_mult * aux * a.inv_rows;
            mv1 += (a.reward_head ? aux : mask * aux) * a.inv_rows;
          }
        } else if (epi == EPI_BWD_NORM) {
          float g[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] = colc[64 + cbase + cl + j];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const float4 y4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);    // dy (0 for rows past the end)
            const float4 h4 = *reinterpret_cast<const float4*>(stg2 + r * kStgLd + cl);   // zh
            const float dy[4] = {y4.x, y4.y, y4.z, y4.w}, zh[4] = {h4.x, h4.y, h4.z, h4.w};
            const float rstd = prow[i], S1 = rowa[r], S2 = rowb[r];
            float dz[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              dz[j] = rstd * (dy[j] * g[j] - S1 - zh[j] * S2);
              pg[j] += dy[j] * zh[j];
              pb[j] += dy[j];
              pd[j] += dz[j];
            }
            if (row < M) store_bf16x4(twin_p + static_cast<size_t>(row) * ld_twin + col, dz[0], dz[1], dz[2], dz[3]);
          }
        } else {   // EPI_BWD_ACT
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rg + 4 * i, row = m0 + lane_base + r;
            const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);   // 0 for rows past the end
            const float dz0 = x4.x * fswish_grad(pre[i].x), dz1 = x4.y * fswish_grad(pre[i].y);
            const float dz2 = x4.z * fswish_grad(pre[i].z), dz3 = x4.w * fswish_grad(pre[i].w);
            pd[0] += dz0; pd[1] += dz1; pd[2] += dz2; pd[3] += dz3;
            if (row < M && col < N) store_bf16x4(twin_p + static_cast<size_t>(row) * ld_twin + col, dz0, dz1, dz2, dz3);
          }
        }
        if (threadIdx.x == 64) SLAB_STAMP(10);
        // ---- column sums (bias / gamma / beta gradients) and metric partials of this tile ----
        const bool need_g = (epi == EPI_BWD_NORM && st.g_gamma != nullptr);
        const bool need_b = (epi == EPI_BWD_NORM && st.norm == NORM_LAYER && st.g_beta != nullptr);
        const bool need_d = ((epi == EPI_BWD_NORM || epi == EPI_BWD_ACT || epi == EPI_FWD_AUX) && st.g_bias != nullptr);
        if (need_g || need_b || need_d || epi == EPI_FWD_AUX) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pg[j] += __shfl_xor_sync(0xffffffffu, pg[j], 8);  pg[j] += __shfl_xor_sync(0xffffffffu, pg[j], 16);
            pb[j] += __shfl_xor_sync(0xffffffffu, pb[j], 8);  pb[j] += __shfl_xor_sync(0xffffffffu, pb[j], 16);
            pd[j] += __shfl_xor_sync(0xffffffffu, pd[j], 8);  pd[j] += __shfl_xor_sync(0xffffffffu, pd[j], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              colpart[(0 * 8 + ew) * 32 + cl + j] = pg[j];
              colpart[(1 * 8 + ew) * 32 + cl + j] = pb[j];
              colpart[(2 * 8 + ew) * 32 + cl + j] = pd[j];
            }
          }
          if (epi == EPI_FWD_AUX) {
            mv0 = warp_sum(mv0); mv1 = warp_sum(mv1); mv2 = warp_sum(mv2);
            if (lane == 0) { metpart[ew * 4 + 0] = mv0; metpart[ew * 4 + 1] = mv1; metpart[ew * 4 + 2] = mv2; }
          }
          epi_bar();
          const int te = ew * 32 + lane;
          if (te < 64) {
            const int c = n0 + te, h = te >> 5, cc = te & 31;
            if (c < N) {
              float sg = 0.f, sb = 0.f, sd = 0.f;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                sg += colpart[(0 * 8 + h * 4 + w) * 32 + cc];
                sb += colpart[(1 * 8 + h * 4 + w) * 32 + cc];
                sd += colpart[(2 * 8 + h * 4 + w) * 32 + cc];
              }
              if (need_g) atomicAdd(st.g_gamma + c, sg);
              if (need_b) atomicAdd(st.g_beta + c, sb);
              if (need_d) atomicAdd(st.g_bias + c, sd);
            }
          } else if (te == 64 && epi == EPI_FWD_AUX) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int w = 0; w < 8; ++w) { a0 += metpart[w * 4 + 0]; a1 += metpart[w * 4 + 1]; a2 += metpart[w * 4 + 2]; }
            atomicAdd(rargs.metrics + M_CRITIC_LOSS, a0);
            atomicAdd(rargs.metrics + M_AUX_MEAN, a1);
            if (a2 != 0.f) atomicAdd(rargs.metrics + M_REWARD_MEAN, a2);
          }
          epi_bar();
        }
      }
      tc_fence_before();
      if (threadIdx.x == 64) SLAB_STAMP(7);
    }
    it += static_cast<uint32_t>(my_tiles * nkb);
    if (my_tiles > 0 && epi != EPI_NONE) ++tf;
  }
  // nobody leaves (and frees its shared memory / TMEM) while a peer may still read its statistics
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (timed && threadIdx.x == 64) timing[num_steps] = clock64();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prog.tmem_cols));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

bool slab_encode_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  if ((ld & 7) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_slab(const SlabProgram& prog, int cluster_ctas, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlabSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (prog.rows <= 0 || prog.num_steps <= 0) return cudaSuccess;
  if (cluster_ctas < 1 || cluster_ctas > 8 || prog.num_steps > kSlabMaxSteps) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster_ctas, (prog.rows + SB_M - 1) / SB_M, 1);
  cfg.blockDim = dim3(kSlabThreads);
  cfg.dynamicSmemBytes = kSlabSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, slab_kernel, prog);
}

}  // namespace reppo
