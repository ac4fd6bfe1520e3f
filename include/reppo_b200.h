/* reppo_b200.h -- C ABI of libreppo_b200.so: REPPO's update step on B200 (sm_100a).
 *
 * The reference (cvoelcker/reppo) has no FFI: its "plugin API" for this path is a set of Python
 * factory closures.  Every entry point below names the reference code it replaces
 * (paths relative to the reference root).  Plain pointers and sizes only -- no torch types.
 *
 * Ownership: the CALLER owns all device memory (parameter / gradient / Adam arenas, the rollout
 * batch, the workspace, metric buffers); the library owns only host-side descriptors.  Calls are
 * asynchronous on the stream passed in, never synchronise, never allocate, and are CUDA-graph
 * capturable.  Return value: 0 on success, negative code on error (message via
 * reppo_last_error).  There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef REPPO_B200_H
#define REPPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct reppo_ctx reppo_ctx;
typedef void* reppo_stream; /* cudaStream_t */

enum { REPPO_OK = 0, REPPO_ERR_INVALID = -1, REPPO_ERR_UNSUPPORTED = -2, REPPO_ERR_CUDA = -3,
       REPPO_ERR_WORKSPACE = -4, REPPO_ERR_NCCL = -5 };

/* which of the reference's two sets of constants to follow */
enum { REPPO_SEM_JAX = 0,    /* src/jaxrl/reppo.py + src/networks/jax_models.py (parity target) */
       REPPO_SEM_TORCH = 1   /* src/torchrl/reppo.py + src/networks/torch_models.py             */ };
enum { REPPO_NORM_NONE = 0, REPPO_NORM_RMS = 1 /* jax_models.py:46, torch_models.py:76 */,
       REPPO_NORM_LAYER = 2 /* src/reppo_mj_playground.py:65-72 */ };
enum { REPPO_KL_CLIPPED = 0, REPPO_KL_FULL = 1, REPPO_KL_VALUE = 2 }; /* actor_kl_clip_mode, config/reppo.yaml:64 */
enum { REPPO_GEMM_FP32 = 0,  /* fp32 FFMA GEMMs: reference-precision mode ("highest", src/jaxrl/__init__.py:3) */
       REPPO_GEMM_BF16 = 1,  /* tcgen05 bf16 operands / fp32 TMEM accumulators ("medium", src/torchrl/reppo.py:35) */
       REPPO_GEMM_TF32 = 2   /* tcgen05.mma.kind::tf32 on the fp32 operands themselves (10-bit mantissa products, fp32 accumulation):
                                reference precision on tensor cores; everything outside the GEMMs as in REPPO_GEMM_FP32.
                                Activation operands whose leading dimension TMA cannot address (not a multiple of 4 floats) are
                                re-laid into row-padded scratch first; the FFMA kernel remains the fall-back. */ };
enum { REPPO_NET_CRITIC = 0, REPPO_NET_ACTOR = 1 };

/* Static shape of the two networks (config/reppo.yaml:38-55) and of the workspace. */
typedef struct reppo_shape {
  int32_t obs_dim, critic_obs_dim, action_dim;
  int32_t critic_hidden, actor_hidden, num_bins;
  int32_t enc_layers, head_layers, pred_layers, actor_layers; /* each >= 2 */
  int32_t critic_norm, actor_norm;                             /* REPPO_NORM_*  */
  int32_t semantics;                                           /* REPPO_SEM_*   */
  int32_t gemm_mode;                                           /* REPPO_GEMM_*  */
  int32_t max_rows;   /* largest minibatch / forward batch the workspace must hold */
  int32_t kl_samples; /* 16: src/jaxrl/reppo.py:557, src/torchrl/reppo.py:308 */
  float vmin, vmax;   /* config/env/<env>.yaml vmin / vmax */
  int32_t max_batch_rows; /* optional (0 = off): rows one epoch of reppo_update visits (num_mini_batches * minibatch).  When
                             set, the workspace also holds two epoch-sized sets of pre-gathered first-layer inputs and the
                             minibatch gathers leave the per-step chains (one gather per epoch instead of two per step) */
} reppo_shape;

/* Per-call hyper-parameters (config/reppo.yaml:16-70). */
typedef struct reppo_hyper {
  float gamma, lmbda, lr, max_grad_norm; /* max_grad_norm <= 0 disables clipping */
  float kl_bound, ent_target_mult, aux_loss_mult, actor_min_std;
  int32_t kl_clip_mode;                  /* REPPO_KL_* */
  int32_t reduce_kl, update_entropy_lagrangian, update_kl_lagrangian; /* JAX semantics only */
  int32_t partial_reset;                 /* cfg.env.partial_reset, src/torchrl/reppo.py:237 (Torch semantics only) */
  int32_t world_size;                    /* data-parallel ranks sharing each minibatch (>= 1); losses are means over
                                            world_size * mb rows and gradients are all-reduced when a communicator is attached */
  int32_t lr_anneal_steps;               /* cfg.anneal_lr (src/jaxrl/reppo.py:239-244): > 0 = optax.linear_schedule(lr, 0,
                                            lr_anneal_steps) evaluated on the device at each network's own optimizer step
                                            counter (lr is NOT re-read from the host per step, so a captured graph stays
                                            valid across updates); 0 = constant lr */
  int32_t reverse_kl;                    /* cfg.reverse_kl (src/jaxrl/reppo.py:543-553, JAX semantics only): the KL estimate is
                                            E_{a ~ pi}[log pi(a) - log pi_old(a)] over 16 reparameterised samples of the CURRENT
                                            policy (gradient through the samples) instead of the forward KL from pi_old */
} reppo_hyper;

/* One network's flat fp32 arenas (layout: reppo_param_* below) plus its Adam step counter. */
typedef struct reppo_net_state {
  float* params;
  float* grads;
  float* adam_m;
  float* adam_v;
  int32_t* step; /* device int32, number of optimizer steps taken */
} reppo_net_state;

typedef struct reppo_state {
  reppo_net_state critic;
  reppo_net_state actor;
  const float* actor_old; /* parameters of the behaviour policy (actor_target, src/jaxrl/reppo.py:457-464) */
} reppo_state;

/* Flattened rollout batch, row r = t * num_envs + n (src/jaxrl/reppo.py:452-455; src/torchrl/reppo.py:219).
 * All float32; `target` is the TD(lambda) return written by reppo_td_lambda. */
typedef struct reppo_batch {
  const float* obs;        /* [S, obs_dim]        */
  const float* critic_obs; /* [S, critic_obs_dim] */
  const float* action;     /* [S, action_dim]     */
  const float* reward;     /* [S] raw reward (JAX aux head, src/jaxrl/reppo.py:497); may be NULL in Torch semantics */
  const float* done;       /* [S] */
  const float* truncated;  /* [S] */
  const float* next_emb;   /* [S, critic_hidden] */
  const float* target;     /* [S] */
  int64_t num_rows;        /* S */
} reppo_batch;

/* Slots of the per-step metric vector (float32[REPPO_NUM_METRICS]); sums are already divided by the row count. */
enum {
  REPPO_M_CRITIC_LOSS = 0,   /* JAX `loss` (:510) / Torch `qf_loss` (:264) */
  REPPO_M_CE_MEAN,           /* mean_i ce_i (unmasked) */
  REPPO_M_AUX_MEAN,          /* JAX mean aux_loss (:498-502) / Torch `embedding_loss` (:255-262) */
  REPPO_M_VALUE_LOSS,        /* JAX `value_loss` (:505-509) */
  REPPO_M_Q_MEAN,            /* JAX `q` (:519) */
  REPPO_M_TARGET_MEAN,       /* `target_values` / `qf_mean` */
  REPPO_M_TARGET_MAX,        /* `qf_max` */
  REPPO_M_TARGET_MIN,        /* `qf_min` */
  REPPO_M_CRITIC_GRAD_NORM,
  REPPO_M_ACTOR_LOSS,        /* total actor loss */
  REPPO_M_POLICY_LOSS,       /* mean of the pathwise / KL-clipped term */
  REPPO_M_KL,
  REPPO_M_ENTROPY,           /* mean(-log_prob) */
  REPPO_M_TEMPERATURE,
  REPPO_M_LAGRANGIAN,
  REPPO_M_ENTROPY_LOSS,
  REPPO_M_LAGRANGIAN_LOSS,
  REPPO_M_ACTOR_Q_MEAN,
  REPPO_M_ABS_PRED_ACTION,
  REPPO_M_ACTOR_GRAD_NORM,
  REPPO_M_REWARD_MEAN,
  REPPO_M_ABS_BATCH_ACTION,
  REPPO_NUM_METRICS = 24
};

/* Schedule of one whole update = src/jaxrl/reppo.py:417-673 (learn_step) / src/torchrl/reppo.py:614-632. */
typedef struct reppo_plan {
  int32_t num_epochs, num_mini_batches, minibatch; /* minibatch = rows per step on THIS rank */
  const int32_t* perms;   /* [num_epochs, num_mini_batches * minibatch] row indices into the batch   */
  const float* eps_pi;    /* [num_epochs, minibatch, A]  standard normals, reused by every minibatch
                             of the epoch (late-bound key, src/jaxrl/reppo.py:535) */
  const float* eps_kl;    /* [num_epochs, kl_samples, minibatch, A] */
  float* step_metrics;    /* [num_epochs * num_mini_batches, REPPO_NUM_METRICS] or NULL */
  float* metrics;         /* [REPPO_NUM_METRICS] mean over the minibatches of the LAST epoch (:660,671) */
  int32_t use_graph;      /* capture the step sequence in a CUDA graph and replay it */
  int32_t noise_per_minibatch; /* 0: eps_* indexed by epoch only (JAX, above); 1: eps_pi [E, n_mb, mb, A] and
                                  eps_kl [E, n_mb, kl_samples, mb, A] -- fresh draws for every step like
                                  src/torchrl/reppo.py:300,308 */
  float* old_policy_out;  /* optional scratch [batch.num_rows, 2A]: when given, the behaviour policy (actor_old) is
                             evaluated ONCE per update on every row instead of once per minibatch step -- the same
                             values, since actor_old is frozen during the update (src/jaxrl/reppo.py:457-464) */
} reppo_plan;

/* ---- context ------------------------------------------------------------------------------ */
/* Replaces the network construction in src/torchrl/reppo.py:500-525 (shapes only; no allocation). */
int reppo_ctx_create(const reppo_shape* shape, reppo_ctx** out);
void reppo_ctx_destroy(reppo_ctx* ctx);
const char* reppo_last_error(const reppo_ctx* ctx); /* ctx may be NULL: error of the last failed create */
int reppo_version(void);

/* Flat parameter arena layout = the reference Torch state_dict order (SURVEY.md A.7;
 * src/networks/torch_models.py:209-335), weights stored [out, in] row-major. */
int64_t reppo_param_count(const reppo_ctx* ctx, int net);
int32_t reppo_param_num_tensors(const reppo_ctx* ctx, int net);
int reppo_param_tensor(const reppo_ctx* ctx, int net, int32_t index, char* name, int32_t name_cap,
                       int64_t* offset, int32_t* rows, int32_t* cols /* 0 for vectors, rows=0 for scalars */);

/* Scratch the caller must provide (device, 256-byte aligned), sized for shape.max_rows. */
size_t reppo_workspace_bytes(const reppo_ctx* ctx);
int reppo_set_workspace(reppo_ctx* ctx, void* workspace, size_t bytes);
/* Optional: override the library's own fp32 linspace tables with host arrays computed exactly like the
 * reference does (torch.linspace): HL-Gauss bin edges [num_bins+1] (src/torchrl/reppo_util.py:760-765,
 * src/jaxrl/utils.py:48-50) and the value support [num_bins] (torch_models.py:265-267). NULL keeps the default. */
int reppo_set_bins(reppo_ctx* ctx, const float* edges_host, const float* support_host);

/* ---- K1: soft TD(lambda) targets ------------------------------------------------------------
 * convention REPPO_SEM_JAX  : compute_nstep_lambda + reverse scan, src/jaxrl/reppo.py:422-450
 *            (truncation flag / log importance weight of step t+1, seed value[T-1]).
 * convention REPPO_SEM_TORCH: compute_gve, src/torchrl/reppo.py:175-189 (truncated[t], seed 0, no
 *            importance weights; row T-1 of `truncated` is treated as 1 -- when mutate_truncated
 *            is non-zero it is also WRITTEN to 1 like the reference does at :178).
 * All arrays [T, N] float32 device pointers; importance_weight may be NULL (== 0).  gamma / lmbda
 * are the config's doubles: the reference multiplies fp32 tensors by Python scalars, so (1 - lmbda)
 * is rounded to fp32 from the double difference in the Torch convention. */
int reppo_td_lambda(const float* soft_reward, const float* value, const float* done, float* truncated,
                    const float* importance_weight, float* target, int32_t T, int64_t N, double gamma,
                    double lmbda, int32_t convention, int32_t mutate_truncated, reppo_stream stream);

/* ---- network forwards (rollout side / façade modules) ---------------------------------------
 * Critic.forward, src/networks/torch_models.py:278-285 (value, logits, next_pred, features);
 * any output pointer may be NULL.  pred is [rows, H] (Torch) or [rows, H+1] (JAX, slot 0 = reward). */
int reppo_critic_forward(reppo_ctx* ctx, const float* critic_params, const float* critic_obs, const float* action,
                         int32_t rows, float* value, float* logits, float* pred, float* features, reppo_stream stream);
/* Actor.forward, src/networks/torch_models.py:321-335: mean, std = exp(log_std) + min_std. */
int reppo_actor_forward(reppo_ctx* ctx, const float* actor_params, const float* obs, int32_t rows, float min_std,
                        float* mean, float* std, reppo_stream stream);

/* ---- rollout side: what the collection loop runs per environment step --------
 * EmpiricalNormalization, src/torchrl/reppo_util.py:394-471: when update != 0 the running mean / var / std [dim] and the
 * int64 device counter are first merged with the batch (Chan's parallel algorithm, :432-466; skipped once
 * count >= until when until >= 0), then out = (x - mean) / (std + eps) (center != 0) or x / (std + eps).  No ctx needed. */
int reppo_obs_normalize(const float* x, int64_t rows, int32_t dim, float* mean, float* var, float* std, int64_t* count, float eps,
                        int32_t update, int32_t center, int64_t until, float* out, reppo_stream stream);
/* Policy step of collect_fn (src/torchrl/reppo.py:104-109,146) / step_env (src/jaxrl/reppo.py:347-364): actor MLP forward,
 * action = tanh(mu + std * scale * eps) with injected standard normals eps [rows, A] and an optional per-row exploration scale
 * (jax_models.py:362-367), log_prob [rows] of the action clipped to +-clip under the UNscaled policy and (optional)
 * log_prob_scaled under the scaled one; clip <= 0: distrax sample_and_log_prob (log-prob at the pre-tanh sample).
 * store_clipped != 0 writes the clipped action (src/jaxrl/reppo.py:357). */
int reppo_policy_act(reppo_ctx* ctx, const float* actor_params, const float* obs, const float* eps, const float* scale, int32_t rows,
                     float min_std, float clip, int32_t store_clipped, float* action, float* log_prob, float* log_prob_scaled,
                     reppo_stream stream);
/* Bootstrap half of the same step (src/torchrl/reppo.py:117-138; src/jaxrl/reppo.py:367-374): next action + log-prob from the
 * policy at next_obs, Critic.forward(next_critic_obs, next_action) -> next_value [rows], next_emb [rows, H] (= features),
 * soft_reward = reward - gamma * log_prob * exp(log_temp).  Output pointers may be NULL. */
int reppo_bootstrap(reppo_ctx* ctx, const float* actor_params, const float* critic_params, const float* next_obs,
                    const float* next_critic_obs, const float* eps, const float* reward, int32_t rows, float gamma, float min_std,
                    float clip, float* next_value, float* next_emb, float* soft_reward, float* next_action, float* next_log_prob,
                    reppo_stream stream);

/* One Linear layer of FCNN (src/networks/torch_models.py:71-79, jax_models.py:35-49) in isolation -- the
 * GEMM family the MLP forward / backward is built from; used by the kernel tests and bench.py's roofline leg.
 *   op 0  forward : y[rows,out]  = x[rows,in] w[out,in]^T + bias   (bias may be NULL)
 *   op 1  dgrad   : dx[rows,in]  = dy[rows,out] w[out,in]
 *   op 2  wgrad   : dw[out,in]   = dy[rows,out]^T x[rows,in]       (dw must be zeroed by the caller)
 * a / b / c are (x, w, y), (dy, w, dx), (dy, x, dw) respectively.  In REPPO_GEMM_BF16 mode the fp32 operands are first
 * rounded into bf16 staging buffers; op | 0x100 skips that and reuses the staging of the previous identical call
 * (bench.py times the tcgen05 kernel alone this way). */
int reppo_linear(reppo_ctx* ctx, int32_t op, int32_t rows, int32_t in_features, int32_t out_features, const float* a,
                 const float* b, float* c, const float* bias, reppo_stream stream);

/* Roofline probe of the memory-bound loss / sampling kernels of one minibatch step at an arbitrary row count (bench.py's
 * `loss_kernels` leg and the kernel tests; not a reference entry point).  Every buffer is caller-owned device memory.
 *   which 0  HL-Gauss cross-entropy (critic_loss_fn, src/jaxrl/reppo.py:474-513; hl_gauss src/jaxrl/utils.py:43-59):
 *            in0 = head output [rows, B], in1 = target [rows], in2 = truncated [rows], in3 = zero_dist [B];
 *            out0 = d logits [rows, B]; scratch >= 24 + 2 B floats
 *   which 1  next-embedding (+ reward) regression (src/jaxrl/reppo.py:491-507):
 *            in0 = pred [rows, P] (P = H + 1 JAX / H Torch), in1 = next_emb [rows, H], in2 = reward [rows], in3 = done [rows];
 *            out0 = d pred [rows, P]; scratch >= 24 + P floats
 *   which 2  tanh-Gaussian sample, log-prob and 16-sample KL to the behaviour policy (actor_loss, src/jaxrl/reppo.py:526-571):
 *            in0 = policy output [rows, 2A], in1 = behaviour-policy output [rows, 2A], in2 = eps_pi [rows, A],
 *            in3 = eps_kl [kl_samples, rows, A]; out0 = action [rows, A]; scratch >= rows * (2 + 2 A) floats */
int reppo_kernel_probe(reppo_ctx* ctx, int32_t which, int32_t rows, const float* in0, const float* in1, const float* in2,
                       const float* in3, float* out0, float* scratch, const reppo_hyper* hyper, reppo_stream stream);

/* ---- one minibatch step ---------------------------------------------------------------------
 * update_critic, src/torchrl/reppo.py:227-283  == critic_loss_fn + apply_gradients,
 * src/jaxrl/reppo.py:474-524,624-626.  idx: [mb] int32 rows of `batch` (device). */
int reppo_critic_step(reppo_ctx* ctx, const reppo_state* state, const reppo_batch* batch, const int32_t* idx,
                      int32_t mb, const reppo_hyper* hp, float* metrics /* [REPPO_NUM_METRICS] device */,
                      reppo_stream stream);
/* update_actor, src/torchrl/reppo.py:291-358 == actor_loss + apply_gradients, src/jaxrl/reppo.py:526-634.
 * eps_pi [mb, A], eps_kl [kl_samples, mb, A]: injected standard normals (device). */
int reppo_actor_step(reppo_ctx* ctx, const reppo_state* state, const reppo_batch* batch, const int32_t* idx,
                     int32_t mb, const float* eps_pi, const float* eps_kl, const reppo_hyper* hp,
                     float* metrics, reppo_stream stream);
/* Gradients only (no clip / Adam): used by the parity tests to compare against autograd. */
int reppo_critic_grad(reppo_ctx* ctx, const reppo_state* state, const reppo_batch* batch, const int32_t* idx,
                      int32_t mb, const reppo_hyper* hp, float* metrics, reppo_stream stream);
int reppo_actor_grad(reppo_ctx* ctx, const reppo_state* state, const reppo_batch* batch, const int32_t* idx,
                     int32_t mb, const float* eps_pi, const float* eps_kl, const reppo_hyper* hp,
                     float* metrics, reppo_stream stream);
/* clip_by_global_norm -> adam (src/jaxrl/reppo.py:246-257) / clip_grad_norm_ -> AdamW
 * (src/torchrl/reppo.py:270-273,527-534) on one flat arena of n floats. grad_norm_out: device float or NULL. */
int reppo_clip_adam(reppo_ctx* ctx, const reppo_net_state* net, int64_t n, const reppo_hyper* hp,
                    float* grad_norm_out, reppo_stream stream);

/* ---- the whole update ----------------------------------------------------------------------
 * learn_step, src/jaxrl/reppo.py:417-673 minus the scan (call reppo_td_lambda first) and minus the
 * actor_target sync (the caller copies actor -> actor_old, :457-464 / src/torchrl/reppo.py:631-632). */
int reppo_update(reppo_ctx* ctx, const reppo_state* state, const reppo_batch* batch, const reppo_plan* plan,
                 const reppo_hyper* hp, reppo_stream stream);
/* kernels launched by the most recent reppo_update / step call on this ctx (bench.py's gpu_launches) */
int64_t reppo_launch_count(const reppo_ctx* ctx);

/* ---- data-parallel gradient all-reduce (not in the reference) ----------------
 * The library dlopen()s libnccl.so.2; rank 0 creates the 128-byte unique id, the host broadcasts it.
 * reppo_nccl_init may be called twice (with two different ids): the first communicator serves the critic chain, the
 * second the actor chain, which lets the step pipeline of reppo_update keep both chains in flight under DP. */
int reppo_nccl_unique_id(reppo_ctx* ctx, void* id_out_128);
int reppo_nccl_init(reppo_ctx* ctx, const void* id_128, int32_t rank, int32_t world_size);
int reppo_nccl_finalize(reppo_ctx* ctx);

/* The same all-reduce without NCCL: fused into the clip + Adam launch over NVSwitch MULTICAST memory.
 * The host places both gradient arenas and a flag block in ONE symmetric allocation (CUDA VMM + multicast object; from
 * Python: torch.distributed._symmetric_memory) and hands over its three mappings.  Every step's clip+Adam kernel then
 * (1) meets its twin blocks on the other ranks, (2) replaces its slice of the arena by the switch-side sum
 * (multimem.ld_reduce -> multimem.st), (3) meets the twins again and goes on to the global norm, clip and Adam:
 * no separate collective launch, no NCCL kernel in the chain.  world_size must divide 16.  state.{critic,actor}.grads of
 * every later data-parallel call must be local_base + {critic,actor}_grads_off.  flags: >= (3*256*16 + 16)*4 bytes, zeroed, and
 * all ranks past a host barrier before the first update.  When NCCL communicators are attached as well, NCCL serves unless
 * REPPO_ALLREDUCE=mc is set (measured: DESIGN.md section 6). */
typedef struct reppo_multicast {
  void* local_base;        /* this rank's copy */
  void* mc_base;           /* multicast mapping (all ranks' copies at once) */
  void* peer_base[16];     /* rank r's copy as mapped here; peer_base[rank] == local_base */
  int64_t critic_grads_off, actor_grads_off, flags_off;   /* bytes, 16-byte aligned */
  int64_t flags_bytes;
  int32_t rank, world_size;
} reppo_multicast;
int reppo_dp_attach_multicast(reppo_ctx* ctx, const reppo_multicast* m);
/* *timed_out = 1 when a cross-GPU wait of the fused all-reduce gave up (a rank was lost); results are then invalid */
int reppo_dp_status(reppo_ctx* ctx, int32_t* timed_out);
/* profiling aid: with REPPO_MC_STAMPS=1 in the environment at attach time every clip+Adam block records globaltimer stamps
 * of its all-reduce phases; this copies the ring (8 + 8192 x 8 uint64 words) out.  tools/mc_stamps.py reads it. */
int reppo_dp_stamps(reppo_ctx* ctx, uint64_t* out, int64_t words);

#ifdef __cplusplus
}
#endif
#endif /* REPPO_B200_H */
