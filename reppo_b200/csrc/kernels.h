// Internal launcher interface between api.cu (host orchestration / C ABI) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reppo {

enum { GEMM_STORE = 0, GEMM_ACCUM = 1, GEMM_ATOMIC = 2 };
enum { NORM_NONE = 0, NORM_RMS = 1, NORM_LAYER = 2 };
enum { KL_CLIPPED = 0, KL_FULL = 1, KL_VALUE = 2 };
constexpr int kMaxPartials = 512;

// metric slots: must mirror include/reppo_b200.h
enum {
  M_CRITIC_LOSS = 0, M_CE_MEAN, M_AUX_MEAN, M_VALUE_LOSS, M_Q_MEAN, M_TARGET_MEAN, M_TARGET_MAX, M_TARGET_MIN,
  M_CRITIC_GRAD_NORM, M_ACTOR_LOSS, M_POLICY_LOSS, M_KL, M_ENTROPY, M_TEMPERATURE, M_LAGRANGIAN, M_ENTROPY_LOSS,
  M_LAGRANGIAN_LOSS, M_ACTOR_Q_MEAN, M_ABS_PRED_ACTION, M_ACTOR_GRAD_NORM, M_REWARD_MEAN, M_ABS_BATCH_ACTION,
  M_NUM = 24
};
constexpr uint32_t kCriticMetricMask = 0x1FFu | (1u << M_REWARD_MEAN) | (1u << M_ABS_BATCH_ACTION);
constexpr uint32_t kActorMetricMask = ((1u << (M_ACTOR_GRAD_NORM + 1)) - 1u) & ~0x1FFu;

struct HLConsts {           // categorical head + HL-Gauss constants
  const float* edges;       // [B+1] fp32 linspace(vmin - bw/2, vmax + bw/2)
  const float* support;     // [B]   fp32 linspace(vmin, vmax)
  float vmin, vmax;
  float den;                // sqrt(2) * sigma (+1e-6 torch)
  float z_eps;              // 0 (jax) | 1e-6 (torch)
  float bias_scale;         // 40.0 (jax_models.py:298) | 40.9 (torch_models.py:282)
  int num_bins;
};

struct ActorConsts {
  float min_std, action_clip, kl_bound, target_entropy, reduce_kl;
  int clip_policy_sample, old_logp_at_clipped, kl_mode, update_entropy, update_kl;
  int reverse_kl;   // KL(pi || pi_old) from samples of the current policy (src/jaxrl/reppo.py:543-553)
};

struct AdamConsts {
  float lr, b1, b2, eps, weight_decay, max_grad_norm;
  int torch_clip;
  int anneal_steps;   // > 0: optax.linear_schedule(lr, 0, anneal_steps) evaluated at the device step counter
};

cudaError_t launch_td_lambda(const float* sr, const float* val, const float* done, float* trunc, const float* iw,
                             float* out, int T, long long N, double gamma, double lmbda, int convention, int mutate,
                             cudaStream_t st);

cudaError_t launch_gemm_f32(bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                            float* C, int ldc, const float* bias, int mode, int split_k, cudaStream_t st);
// tcgen05 GEMM on bf16 operands (gemm_bf16_tc.cu); see the file header for the operand arrangements
cudaError_t launch_gemm_bf16_tc(bool a_mn, bool b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                float* C, int ldc, const float* bias, int mode, int split_k, void* act_out, int act_ld,
                                cudaStream_t st);
// the same kernel on fp32 operands read as tf32 (reference-precision mode on tensor cores); cudaErrorNotSupported = an operand is
// not TMA-addressable (alignment / leading dimension): run launch_gemm_f32 instead
cudaError_t launch_gemm_tf32_tc(bool a_mn, bool b_mn, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, const float* bias, int mode, int split_k, cudaStream_t st);
// persistent CTA-pair (cta_group::2) variant for large shapes (gemm_bf16_tc2.cu); cudaErrorNotSupported = not eligible
cudaError_t launch_gemm_bf16_tc2(bool a_mn, bool b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                 float* C, int ldc, const float* bias, int mode, int split_k, void* act_out, int act_ld,
                                 cudaStream_t st);
bool gemm_bf16_tc2_eligible(bool a_mn, bool b_mn, int M, int N, int K, int mode, int split_k);
// fused forward Linear + bias + norm + swish (cluster / DSMEM row statistics); cudaErrorNotSupported = shape not eligible
struct NormArgs { const float* gamma; const float* beta; float eps; float* rstd_out; float* mean_out; };
cudaError_t launch_gemm_bf16_tc_norm(int norm, int M, int N, int K, const void* A, int lda, const void* B, int ldb, float* C,
                                     int ldc, const float* bias, void* act_out, int act_ld, const NormArgs& na, cudaStream_t st);
// fp32 -> bf16 copies: plain (row-padded) and, for weights, the [out,in] matrices of one network in one launch
struct PackEntry { const float* w; void* wb; int out, in, ld_b; };
constexpr int kMaxPack = 16;
struct PackTable { int n; int fp32; PackEntry e[kMaxPack]; };   // fp32: the shadow keeps fp32 (tf32 GEMM operands), else bf16
cudaError_t launch_pack_weights(const PackTable& t, cudaStream_t st);
cudaError_t launch_pad_rows_f32(const float* src, int rows, int cols, int ld_src, float* dst, int ld_dst, cudaStream_t st);
cudaError_t launch_cast_bf16(const float* src, int rows, int cols, int ld_src, void* dst, int ld_dst, cudaStream_t st);
cudaError_t launch_colsum(const float* X, int M, int N, int ld, float* out, float* out2, float scale2, cudaStream_t st);

cudaError_t launch_gather_concat(const float* a, int da, const float* b, int db, const int* idx, int b_is_local,
                                 int rows, float* out, int out_ld, void* out_h, int out_ldh, cudaStream_t st);
// two gathers in one launch: out_a[r, :] = a[idx[r], 0:da], out_b[r, 0:db] = b[idx[r], 0:db]  (fp32 + bf16 twins, all optional)
cudaError_t launch_gather_pair(const float* a, int da, const float* b, int db, const int* idx, int rows, float* out_a, int lda,
                               void* out_ah, int ldah, float* out_b, int ldb, void* out_bh, int ldbh, cudaStream_t st);
cudaError_t launch_norm_act_fwd(int norm, const float* z, const float* gamma, const float* beta, int rows, int H,
                                float eps, float* x, float* rstd, float* mean, void* x_h, int ldh, cudaStream_t st);
cudaError_t launch_norm_act_bwd(int norm, const float* dx, const float* dx2, const float* z, const float* gamma, const float* beta,
                                const float* rstd, const float* mean, int rows, int H, float* dz, float* dgamma,
                                float* dbeta, float* dbias, void* dz_h, int ldh, cudaStream_t st);

cudaError_t launch_critic_ce(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, const float* target,
                             const float* truncated, const int* idx, int rows, float inv_rows, int no_mask, float* dlogits,
                             float* metrics, void* dl_h, int ldh, float* dbias, float* dzero, cudaStream_t st);
cudaError_t launch_critic_aux(const float* pred, int P, int H, int reward_head, int mask_done, const float* next_emb,
                              const float* reward, const float* done, const float* truncated, const int* idx, int rows,
                              float inv_rows, int no_mask, float aux_mult, float* dpred, float* metrics, void* dp_h, int ldh,
                              float* dbias, cudaStream_t st);
cudaError_t launch_actor_sample(const float* out, const float* out_old, const int* old_idx, const float* eps_pi,
                                const float* eps_kl, int rows, int A, int kl_samples, const ActorConsts& ac, float* act_out,
                                int act_ld, float* logp, float* kl, float* gmu, float* gsig, void* act_h, int act_ldh,
                                cudaStream_t st);
cudaError_t launch_actor_q(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, const float* kl, int rows,
                           float inv_rows, int kl_mode, float kl_bound, float* q_out, float* dlogits, void* dl_h, int ldh,
                           cudaStream_t st);
cudaError_t launch_actor_grad(const float* out, const float* eps_pi, const float* act, int act_ld, const float* dact,
                              int dact_ld, const float* logp, const float* kl, const float* q, const float* gmu,
                              const float* gsig, const float* actor_params, int rows, int A, float inv_rows,
                              const ActorConsts& ac, float* dout, float* actor_grads, float* metrics, void* dout_h, int ldh,
                              float* dbias, cudaStream_t st);
cudaError_t launch_value_head(const float* head_out, int ld, const float* zero_dist, const HLConsts& hl, int rows,
                              float* value, float* logits, cudaStream_t st);
cudaError_t launch_mean_std(const float* out, int rows, int A, float min_std, float* mean, float* std, cudaStream_t st);

// rollout side (rollout.cu)
cudaError_t launch_policy_sample(const float* out, const float* eps, const float* scale, int rows, int A, float min_std, float clip,
                                 int store_clipped, float* action, int act_ld, void* act_h, int act_ldh, float* logp,
                                 float* logp_scaled, cudaStream_t st);
cudaError_t launch_soft_reward(const float* reward, const float* logp, const float* log_temp, float gamma, int rows, float* out,
                               cudaStream_t st);
cudaError_t launch_obs_normalize(const float* x, long long rows, int D, float* mean, float* var, float* stdv, long long* count,
                                 float eps, int update, int center, long long until, float* out, cudaStream_t st);

// bf16 shadows of the weight matrices, refreshed by the Adam kernel itself (REPPO_GEMM_BF16)
struct ShadowEntry { long long start, end; int in, ld; void* dst; };
struct ShadowTable { int n; int fp32; ShadowEntry e[kMaxPack]; };
// Data-parallel gradient all-reduce over NVSwitch multicast memory, fused into the clip+Adam launch (optim.cu).  The gradient
// arena lives in a symmetric allocation that every rank maps three ways: its own copy (what the gradient kernels write),
// every peer's copy (flags only) and the MULTICAST mapping, on which `multimem.ld_reduce` returns the switch-side sum over
// all ranks and `multimem.st` writes every rank's copy at once.  world <= 1: no exchange.
constexpr int kMcMaxRanks = 16;
constexpr int kMcFlagBlocks = 256;    // flag rows per chain (>= grid of clip_adam_kernel)
constexpr int kMcChains = 3;
constexpr int kMcStampWords = 8 + 8192 * 8;
struct McArgs {
  float* g_mc;                          // multicast address of this launch's gradient arena
  unsigned int* flags[kMcMaxRanks];     // every rank's flag block of this chain: [kMcFlagBlocks][world] words
  unsigned int* epoch;                  // this rank's epoch word of this chain (+2 per launch; flags carry epoch + 1 / + 2)
  int* status;                          // device word set to 1 when a cross-GPU wait timed out (reppo_dp_status)
  unsigned long long* stamps;           // REPPO_MC_STAMPS=1: [0] = launches recorded, then 8 words per block (tools/mc_stamps.py)
  int world, rank, chain;
};
// bar: two device uint32 {arrival count, generation} of the kernel's grid barrier (zero-initialised once; one pair per
// concurrently running chain)
// true: one launch with a grid barrier (multicast all-reduce, REPPO_ADAM_FUSED=1); false: two launches (partials, then clip + Adam)
bool clip_adam_fused(const McArgs* mc);
cudaError_t launch_clip_adam(const float* p_in, float* p, float* g, float* m, float* v, long long n, int* step, float* partials,
                             unsigned int* bar, const AdamConsts& c, const ShadowTable* shadow, float* grad_norm_out,
                             int zero_grads, cudaStream_t st, const McArgs* mc = nullptr);
cudaError_t launch_metrics_mean(const float* step_metrics, int first, int count, int nm, float* out, cudaStream_t st);
cudaError_t launch_metrics_init(float* m, uint32_t mask, cudaStream_t st);
cudaError_t launch_metrics_init_all(float* m, int steps, cudaStream_t st);

}  // namespace reppo
