// Slab programs: one persistent tcgen05 kernel per critic / actor chain of a minibatch step (slab.cu).
//
// A SLAB is 128 consecutive minibatch rows.  Forward and input-gradient passes of the MLPs never mix rows, so one
// thread-block cluster can carry a slab through the WHOLE chain of a step -- gather, every Linear (+bias, norm,
// swish), the loss, and the dgrad / norm-backward chain -- without leaving the kernel: the only cross-row
// reductions of a step (weight gradients, gradient norm, Adam) run afterwards.  The cluster's CTAs split the
// output columns of every layer in 64-wide tiles; between two layers they meet at a cluster barrier instead of a
// kernel boundary.  This header is the host/device contract: api.cu builds a SlabProgram per launch.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "kernels.h"

namespace reppo {

enum { SLAB_GEMM = 0, SLAB_ROW = 1 };
// epilogues of a GEMM step (what happens to the 128 x 64 fp32 accumulator tile)
enum {
  EPI_NONE = 0,   // keep the accumulator: the next step accumulates on top of it (two-branch gradient sum)
  EPI_FWD_NORM,   // z = acc + b (stored fp32), x = swish(norm(z)) -> bf16 twin, rstd / mean per row
  EPI_FWD_ACT,    // y = acc + b (stored fp32), swish(y) -> bf16 twin
  EPI_FWD_OUT,    // y = acc + b (stored fp32)
  EPI_FWD_AUX,    // pred = acc + b; next-embedding / reward MSE and its gradient -> bf16 twin, bias gradient, metrics
  EPI_BWD_NORM,   // acc = dL/dx, x = swish(norm(z)): dz -> bf16 twin; d gamma, d beta, d bias column sums
  EPI_BWD_ACT,    // dz = acc * swish'(z) -> bf16 twin; d bias column sums
  EPI_BWD_OUT     // acc stored fp32 (gradient w.r.t. the network input)
};
// row phases (one warp per row, the slab's rows spread over every warp of the cluster)
enum { ROW_GATHER_CRITIC = 0, ROW_GATHER_ACTOR, ROW_CE, ROW_SAMPLE, ROW_Q, ROW_ACTOR_GRAD };

constexpr int kSlabMaxSteps = 28;
constexpr int kSlabStages = 4;
constexpr int kSlabThreads = 320;

// everything about a step except its two tensor maps (the kernel keeps a shared-memory copy of these)
struct SlabStepInfo {
  int kind, epi, row;
  int K, N, b_mn, accumulate;
  int norm;                  // NORM_RMS | NORM_LAYER for the *_NORM epilogues
  float eps;
  int ld_out, ld_twin;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* out;                // fp32 output [rows, ld_out]; for the BWD epilogues: the saved forward z (input)
  float* rstd;               // FWD_NORM: out; BWD_NORM: in
  float* mean;
  __nv_bfloat16* twin;       // bf16 output [rows, ld_twin]
  float* g_gamma;            // gradient accumulators (column sums, atomics into the zeroed arena); nullable
  float* g_beta;
  float* g_bias;
};
struct alignas(64) SlabStep {
  CUtensorMap ta, tb;        // A: bf16 [rows, K] (box 64 x 128); B: weights, K-major [N, K] or MN-major [K, N] (box 64 x 64)
  SlabStepInfo f;
};

struct SlabRowArgs {
  // gather (minibatch rows through the permutation)
  const float* obs;
  const float* critic_obs;
  const float* action;
  const int* idx;
  int D, Dc, A, K0;
  __nv_bfloat16* x0_twin;  int x0_ld;    // critic input twin [rows, K0]
  __nv_bfloat16* xa0_twin; int xa0_ld;   // policy input twin [rows, D]
  float* x0a;                            // fp32 [rows, K0]: the sampled action lands in columns Dc..
  // categorical head / HL-Gauss cross-entropy (ROW_CE) and Q (ROW_Q)
  const float* head_out; int ld_head;
  const float* zero_dist;
  HLConsts hl;
  const float* target;
  const float* truncated;
  float inv_rows; int no_mask;
  __nv_bfloat16* dl_twin; int dl_ld;
  float* g_bias_head; float* g_zero;
  float* metrics;
  // auxiliary prediction loss (EPI_FWD_AUX)
  const float* next_emb; const float* reward; const float* done;
  int P, H, reward_head, mask_done; float aux_mult;
  // policy sample + KL (ROW_SAMPLE)
  const float* actor_out; const float* old_out; const int* old_idx;
  const float* eps_pi; const float* eps_kl; int kl_samples;
  ActorConsts ac;
  float* logp; float* kl; float* gmu; float* gsig;
  float* q;
  // policy gradient (ROW_ACTOR_GRAD)
  const float* dx0; const float* actor_params; float* actor_grads;
  __nv_bfloat16* dout_twin; int dout_ld; float* g_bias_actor_last;
};

struct SlabProgram {
  int num_steps;
  int rows;        // minibatch rows
  int tmem_cols;   // power of two >= 64 * (largest number of tiles one CTA owns in a step)
  long long* timing;   // optional [num_steps + 1] clock64() stamps of CTA (0,0) (REPPO_SLAB_TIMING=1; profiling aid)
  SlabRowArgs r;
  SlabStep steps[kSlabMaxSteps];
};

// cluster = (cluster_ctas, 1, 1); grid = (cluster_ctas, ceil(rows / 128))
cudaError_t launch_slab(const SlabProgram& prog, int cluster_ctas, cudaStream_t st);
// encode helpers shared with the host builder (bf16 row-major [rows, cols], leading dimension ld elements)
bool slab_encode_2d(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows);

}  // namespace reppo
