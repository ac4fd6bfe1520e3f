// C ABI + host orchestration of the REPPO update step (see include/reppo_b200.h).
//
// The host side only sequences kernels on the caller's stream; all device memory belongs to the
// caller.  One minibatch step = critic forward/backward/clip/Adam followed by actor
// forward/backward/clip/Adam (src/jaxrl/reppo.py:466-643; src/torchrl/reppo.py:622-629), the
// whole update = epochs x minibatches of those, optionally captured once into a CUDA graph.
//
// GEMM modes.  REPPO_GEMM_FP32: every Linear runs the fp32 FFMA kernel (gemm_f32.cu).
// REPPO_GEMM_BF16: every Linear runs the tcgen05 kernel (gemm_bf16_tc.cu).  Its bf16 operands
// are never produced by separate cast passes inside a step: each elementwise kernel that
// produces a GEMM operand (gather, norm+swish forward/backward, the loss kernels) writes a
// bf16 "twin" next to its fp32 output, and the Adam step is followed by one pack kernel that
// refreshes the bf16 shadows W and W^T of the network it updated.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/reppo_b200.h"
#include "kernels.h"
#include "knobs.h"

using namespace reppo;

namespace {

struct Linear {
  int64_t w = -1, b = -1, g = -1, beta = -1;  // offsets into the flat fp32 arena (floats)
  int in = 0, out = 0;
  int64_t sb = 0;                             // offset into the bf16 shadow W [out, ld_b] (elements)
  int ld_b = 0;
};
struct Mlp {
  std::vector<Linear> layers;
  int norm = NORM_NONE;  // applied after every layer but the last
};
struct TensorInfo {
  std::string name;
  int64_t off;
  int rows, cols;  // cols == 0: vector of `rows`; rows == 0: scalar
};
struct Acts {  // saved activations of one MLP evaluation + its private backward scratch
  std::vector<float*> z, x, rstd, mean;
  std::vector<float*> dz;   // gradient w.r.t. z of every hidden layer (own buffer: wgrad reads it concurrently with the chain)
  float* out = nullptr;
  float* tmp = nullptr;     // dgrad output before the norm backward
};
struct Sem {
  float norm_eps, bias_scale, action_clip;
  int clip_policy_sample, old_logp_at_clipped, reward_head, aux_mask_done, torch_opt;
  float force_min_std;  // < 0: use hp.actor_min_std
};
struct Twin {  // bf16 copy of an fp32 activation buffer (REPPO_GEMM_BF16 only)
  uint16_t* p = nullptr;
  int ld = 0;
};
enum { SLOT_CRITIC = 0, SLOT_ACTOR = 1, SLOT_ACTOR_OLD = 2, SLOT_CRITIC_ALT = 3, NUM_SLOTS = 4 };

// minimal NCCL surface (resolved with dlsym from libnccl.so.2)
typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

struct GraphKey {
  reppo_state state;
  reppo_batch batch;
  reppo_plan plan;
  reppo_hyper hp;
};

std::string g_create_error;

inline int round8(int x) { return (x + 7) & ~7; }

}  // namespace

struct reppo_ctx {
  reppo_shape shape;
  Sem sem;
  bool tc = false;  // REPPO_GEMM_BF16
  Mlp feature, head, pred, actor;
  int64_t zero_dist_off = 0, critic_count = 0, actor_count = 0;
  int64_t shadow_b_elems[2] = {0, 0};  // per network (critic, actor)
  int pred_out = 0, K0 = 0, maxdim = 0;
  std::vector<TensorInfo> critic_tensors, actor_tensors;
  std::vector<float> edges_h, support_h;
  HLConsts hl{};
  // workspace
  char* ws = nullptr;
  size_t ws_bytes = 0, ws_need = 0;
  float *edges_d = nullptr, *support_d = nullptr, *partials = nullptr;
  float *x0 = nullptr, *xs = nullptr, *xa0 = nullptr;
  Acts fa, ha, pa, aa, oa;
  float *x0a = nullptr;  // critic input of the ACTOR step ([critic_obs | sampled action]); separate so that it can be
                         // built while the critic step still reads x0
  float *dxs = nullptr, *dxs2 = nullptr, *dfeat = nullptr, *dlogits = nullptr, *dpred = nullptr, *dout = nullptr, *dx0 = nullptr;
  float *logp = nullptr, *kl = nullptr, *q = nullptr, *gmu = nullptr, *gsig = nullptr, *scratch_metrics = nullptr;
  std::unordered_map<const float*, Twin> twins;
  // epoch-sized pre-gathered first-layer inputs (two sets, alternating by epoch): any row slice of them has a twin too
  struct TwinRange { const float* base; size_t floats; int cols; Twin t; };
  std::vector<TwinRange> twin_ranges;
  float *x0_ep[2] = {nullptr, nullptr}, *xa0_ep[2] = {nullptr, nullptr}, *x0a_ep[2] = {nullptr, nullptr};
  cudaEvent_t ev_epoch[2] = {nullptr, nullptr};
  unsigned int* adam_bar = nullptr;   // 3 x {arrival count, generation}: critic chain, actor chain, stand-alone calls
  uint16_t* shadow_b[NUM_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  // REPPO_GEMM_TF32: aligned, row-padded fp32 copies of the weight matrices (same offsets / leading dimensions as the bf16
  // shadows): the parameter arena follows the reference's state_dict order and its matrices are not 16-byte aligned
  bool tf32 = false;
  float* shadow_f[NUM_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  // ... and row-padded copies of the activation operands whose leading dimension is not a multiple of four floats (the
  // obs + act inputs, the 151-bin / H + 1 / 2A head gradients): two per stream that launches GEMMs, so concurrent chains and
  // their weight-gradient side streams never share one
  static constexpr int kTf32Streams = 8;
  float* tf32_pad[kTf32Streams][2] = {};
  cudaStream_t tf32_pad_owner[kTf32Streams] = {};
  int tf32_pad_used = 0;
  // step pipelining: second copy of the critic parameters (Adam ping-pongs between the two), and a private set of
  // critic-trunk buffers for the actor step, so that critic step k+1 can overlap actor step k
  float* cp_alt = nullptr;
  Acts fa2, ha2;
  float *xs2 = nullptr, *dxs_a = nullptr, *dfeat_a = nullptr, *dlogits_a = nullptr, *partials2 = nullptr;
  bool pipeline = true;       // REPPO_NO_PIPELINE=1 disables
  cudaEvent_t ev_c[4] = {nullptr, nullptr, nullptr, nullptr}, ev_a[4] = {nullptr, nullptr, nullptr, nullptr};
  uint16_t *lin_a = nullptr, *lin_b = nullptr, *lin_c = nullptr;  // scratch operands of reppo_linear
  std::string err;
  int64_t launches = 0;
  // graph cache
  // two entries: a host that double-buffers its rollout staging (H2D of batch k+1 under update k) alternates two pointer sets
  static constexpr int kGraphSlots = 2;
  cudaGraphExec_t graph_exec[kGraphSlots] = {nullptr, nullptr};
  cudaStream_t cap_stream = nullptr;  // capture happens here: the caller's stream may be the legacy default stream
  GraphKey graph_key[kGraphSlots] = {};
  bool graph_valid[kGraphSlots] = {false, false};
  int64_t graph_launches[kGraphSlots] = {0, 0};
  int graph_next = 0;   // slot to evict next
  void invalidate_graphs() { for (int i = 0; i < kGraphSlots; ++i) graph_valid[i] = false; }
  // concurrency: independent branches of a step run on context-owned side streams (captured into the graph as
  // parallel branches): [0] pred-head branch, [1] weight-gradient GEMMs, [2] actor-only prefix of the actor step
  bool concurrent = true;
  bool fuse_norm = true;   // REPPO_NO_FUSE_NORM=1: separate GEMM + norm/swish kernels
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // [3]: weight gradients of the actor chain
  std::vector<cudaEvent_t> events;
  size_t ev_next = 0;
  // nccl
  NcclApi nccl;
  ncclComm_t comm[2] = {nullptr, nullptr};   // [0] critic chain, [1] actor chain (optional: enables step pipelining under DP)
  int world = 1, rank = 0;
  // NVSwitch-multicast all-reduce fused into clip+Adam (reppo_dp_attach_multicast; kernels.h McArgs)
  struct McState {
    bool on = false;
    char* local = nullptr;
    char* mcast = nullptr;
    char* peer[kMcMaxRanks] = {};
    int64_t grads_off[2] = {0, 0};   // critic, actor gradient arenas (bytes into the symmetric buffer)
    int64_t flags_off = 0;
    int world = 1, rank = 0;
    int* status = nullptr;           // device word: a cross-GPU wait timed out
    unsigned long long* stamps = nullptr;   // REPPO_MC_STAMPS=1: phase stamps of the fused all-reduce (reppo_dp_stamps)
  } mc;
  bool mc_pending = false;           // allreduce_grads deferred the exchange to the next clip_adam launch

  Twin twin(const float* p) const {
    auto it = twins.find(p);
    if (it != twins.end()) return it->second;
    for (const TwinRange& r : twin_ranges)
      if (p > r.base && p < r.base + r.floats && (p - r.base) % r.cols == 0)
        return Twin{r.t.p + static_cast<size_t>((p - r.base) / r.cols) * r.t.ld, r.t.ld};
    return Twin{};
  }
};

namespace {

int fail(reppo_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    ++c->launches;                                                                                 \
    if (e__ != cudaSuccess)                                                                        \
      return fail(c, REPPO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)
#define RET(expr)                       \
  do {                                  \
    int r__ = (expr);                   \
    if (r__ != REPPO_OK) return r__;    \
  } while (0)

void add_fcnn(std::vector<TensorInfo>& ts, int64_t& off, int64_t& sb, Mlp& mlp, const std::string& prefix,
              int in_f, int hidden, int out_f, int layers, bool input_activation, int norm) {
  // Sequential indices of torch_models.FCNN (:97-139): an input activation occupies slot 0.
  const int shift = input_activation ? 1 : 0;
  mlp.norm = norm;
  for (int i = 0; i < layers; ++i) {
    Linear L;
    L.in = (i == 0) ? in_f : hidden;
    L.out = (i == layers - 1) ? out_f : hidden;
    const std::string blk = prefix + ".net." + std::to_string(shift + i);
    L.w = off; ts.push_back({blk + ".0.weight", off, L.out, L.in}); off += static_cast<int64_t>(L.out) * L.in;
    L.b = off; ts.push_back({blk + ".0.bias", off, L.out, 0}); off += L.out;
    if (norm != NORM_NONE && i < layers - 1) {
      L.g = off; ts.push_back({blk + ".1.weight", off, L.out, 0}); off += L.out;
      if (norm == NORM_LAYER) { L.beta = off; ts.push_back({blk + ".1.bias", off, L.out, 0}); off += L.out; }
    }
    L.ld_b = round8(L.in);
    L.sb = sb; sb += static_cast<int64_t>(L.out) * L.ld_b; sb = (sb + 63) & ~static_cast<int64_t>(63);
    mlp.layers.push_back(L);
  }
}

// torch.linspace(start, end, steps, dtype=float32) on the CPU (symmetric two-sided evaluation)
std::vector<float> linspace_f32(double start_d, double end_d, int steps) {
  const float start = static_cast<float>(start_d), end = static_cast<float>(end_d);
  std::vector<float> v(steps);
  if (steps == 1) { v[0] = start; return v; }
  const float step = (end - start) / static_cast<float>(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    v[i] = (i < half) ? start + step * static_cast<float>(i) : end - step * static_cast<float>(steps - i - 1);
  return v;
}

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t o = off; off += (bytes + 255) & ~static_cast<size_t>(255); return o; }
};

// lays the workspace out; when base == nullptr only sizes it
size_t layout_workspace(reppo_ctx* c, char* base) {
  const reppo_shape& s = c->shape;
  const size_t R = s.max_rows, H = s.critic_hidden, B = s.num_bins, A = s.action_dim;
  const size_t P = c->pred_out, K0 = c->K0, D = s.obs_dim;
  Bump bp;
  c->twins.clear();
  c->twin_ranges.clear();
  auto F = [&](size_t n) -> float* {
    const size_t o = bp.take(n * sizeof(float));
    return base ? reinterpret_cast<float*>(base + o) : nullptr;
  };
  auto Hf = [&](size_t n) -> uint16_t* {
    const size_t o = bp.take(n * sizeof(uint16_t));
    return base ? reinterpret_cast<uint16_t*>(base + o) : nullptr;
  };
  // fp32 buffer [R, cols] that is also a GEMM operand: give it a bf16 twin in tensor-core mode
  auto FT = [&](size_t cols) -> float* {
    float* p = F(R * cols);
    if (c->tc) {
      Twin t;
      t.ld = round8(static_cast<int>(cols));
      t.p = Hf(R * t.ld);
      if (base) c->twins[p] = t;
    }
    return p;
  };
  auto acts = [&](Acts& a, const Mlp& m, size_t out_dim) {
    a.z.clear(); a.x.clear(); a.rstd.clear(); a.mean.clear();
    for (size_t i = 0; i + 1 < m.layers.size(); ++i) {
      const size_t h = m.layers[i].out;
      a.z.push_back(F(R * h)); a.x.push_back(FT(h)); a.rstd.push_back(F(R)); a.mean.push_back(F(R));
    }
    a.out = F(R * out_dim);
  };
  auto bwd_scratch = [&](Acts& a, const Mlp& m) {
    a.dz.clear();
    size_t hmax = 1;
    for (size_t i = 0; i + 1 < m.layers.size(); ++i) { a.dz.push_back(FT(m.layers[i].out)); hmax = std::max<size_t>(hmax, m.layers[i].out); }
    a.tmp = F(R * hmax);
  };
  c->edges_d = F(B + 1); c->support_d = F(B); c->partials = F(kMaxPartials); c->scratch_metrics = F(M_NUM);
  c->adam_bar = reinterpret_cast<unsigned int*>(F(8));
  c->x0 = FT(K0); c->xs = FT(H); c->xa0 = FT(D);
  acts(c->fa, c->feature, H); acts(c->ha, c->head, B); acts(c->pa, c->pred, P);
  acts(c->aa, c->actor, 2 * A); acts(c->oa, c->actor, 2 * A);
  bwd_scratch(c->fa, c->feature); bwd_scratch(c->ha, c->head); bwd_scratch(c->pa, c->pred); bwd_scratch(c->aa, c->actor);
  acts(c->fa2, c->feature, H); acts(c->ha2, c->head, B); bwd_scratch(c->fa2, c->feature); bwd_scratch(c->ha2, c->head);
  c->xs2 = FT(H); c->dxs_a = F(R * H); c->dfeat_a = FT(H); c->dlogits_a = FT(B); c->partials2 = F(kMaxPartials);
  c->cp_alt = F(static_cast<size_t>(c->critic_count) + 4);
  const size_t maxdim = c->maxdim;
  c->x0a = FT(K0);
  c->dxs = F(R * H); c->dxs2 = F(R * H); c->dfeat = FT(H); c->dlogits = FT(B); c->dpred = FT(P); c->dout = FT(2 * A);
  c->dx0 = F(R * K0);
  c->logp = F(R); c->kl = F(R); c->q = F(R); c->gmu = F(R * A); c->gsig = F(R * A);
  // epoch-sized gathered inputs: [critic_obs | action], obs, [critic_obs | sampled action]
  const size_t E = (s.max_batch_rows > 0 && !knobs().no_pregather) ? static_cast<size_t>(s.max_batch_rows) : 0;
  for (int i = 0; i < 2 && E > 0; ++i) {
    auto big = [&](size_t cols) -> float* {
      float* p = F(E * cols);
      reppo_ctx::TwinRange r{p, E * cols, static_cast<int>(cols), Twin{}};
      if (c->tc) { r.t.ld = round8(static_cast<int>(cols)); r.t.p = Hf(E * r.t.ld); }
      if (base) { if (c->tc) c->twins[p] = r.t; c->twin_ranges.push_back(r); }
      return p;
    };
    c->x0_ep[i] = big(K0); c->xa0_ep[i] = big(D); c->x0a_ep[i] = big(K0);
  }
  if (c->tf32) {
    const size_t md4 = (maxdim + 3) & ~static_cast<size_t>(3);
    for (int i = 0; i < reppo_ctx::kTf32Streams; ++i) { c->tf32_pad[i][0] = F(R * md4); c->tf32_pad[i][1] = F(R * md4); }
    c->shadow_f[SLOT_CRITIC] = F(c->shadow_b_elems[0]);
    c->shadow_f[SLOT_ACTOR] = F(c->shadow_b_elems[1]);
    c->shadow_f[SLOT_ACTOR_OLD] = F(c->shadow_b_elems[1]);
    c->shadow_f[SLOT_CRITIC_ALT] = F(c->shadow_b_elems[0]);
  }
  if (c->tc) {
    c->shadow_b[SLOT_CRITIC] = Hf(c->shadow_b_elems[0]);
    c->shadow_b[SLOT_ACTOR] = Hf(c->shadow_b_elems[1]);
    c->shadow_b[SLOT_ACTOR_OLD] = Hf(c->shadow_b_elems[1]);
    c->shadow_b[SLOT_CRITIC_ALT] = Hf(c->shadow_b_elems[0]);
    const size_t md8 = round8(static_cast<int>(maxdim));
    c->lin_a = Hf(R * md8); c->lin_c = Hf(R * md8); c->lin_b = Hf(maxdim * md8);
  }
  return bp.off;
}

int wgrad_split(int out, int in, int rows, bool tc) {
  if (tc) {
    if (knobs().wgrad_split > 0) return std::min(knobs().wgrad_split, std::max(1, rows / 64));
    // measured at C2 (512 x 512 x 2048 rows): split 1 / 2 / 4 / 8 / 16 -> 160.9 / 147.8 / 140.9 / 147.7 / 166.9 ms per update.
    // The weight-gradient GEMMs run beside the dgrad chain: ~64 CTAs keep the chain's SMs free and the atomics few.
    const int tiles = ((out + 127) / 128) * ((in + 127) / 128);
    int split = std::max(1, 64 / std::max(1, tiles));
    return std::min(split, std::max(1, rows / 128));
  }
  const int tiles = ((out + 63) / 64) * ((in + 63) / 64);
  int split = std::max(1, 296 / std::max(1, tiles));
  return std::min(split, std::max(1, rows / 64));
}

// tf32 mode: `p` [rows, cols] with leading dimension cols as an operand TMA can address -- itself when cols is a multiple of four
// floats, else a row-padded copy in this stream's scratch (`which` = 0 / 1: the two operands of one GEMM); ld receives the
// leading dimension to use.  nullptr: no scratch left (more GEMM streams than kTf32Streams) -> the caller runs the FFMA kernel.
const float* tf32_operand(reppo_ctx* c, const float* p, int rows, int cols, int which, int* ld, cudaStream_t st, cudaError_t* err) {
  *err = cudaSuccess;
  *ld = cols;
  if ((cols & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return p;
  if (rows > c->shape.max_rows || cols > c->maxdim) return nullptr;
  int slot = -1;
  for (int i = 0; i < c->tf32_pad_used; ++i) if (c->tf32_pad_owner[i] == st) slot = i;
  if (slot < 0) {
    if (c->tf32_pad_used >= reppo_ctx::kTf32Streams) return nullptr;
    slot = c->tf32_pad_used++;
    c->tf32_pad_owner[slot] = st;
  }
  float* dst = c->tf32_pad[slot][which];
  if (dst == nullptr) return nullptr;
  *ld = (cols + 3) & ~3;
  *err = launch_pad_rows_f32(p, rows, cols, cols, dst, *ld, st);
  ++c->launches;
  return dst;
}

// y[rows,out] = x[rows,in] W^T + b
// act: optional bf16 twin that receives swish(y) from the GEMM epilogue (tensor-core mode only)
int linear_fwd(reppo_ctx* c, const Linear& L, const float* params, int slot, const float* x, int rows, float* y, Twin act,
               cudaStream_t st) {
  if (c->tc) {
    const Twin tx = c->twin(x);
    if (tx.p == nullptr) return fail(c, REPPO_ERR_INVALID, "internal: forward operand has no bf16 twin");
    CK(launch_gemm_bf16_tc(false, false, rows, L.out, L.in, tx.p, tx.ld, c->shadow_b[slot] + L.sb, L.ld_b, y, L.out,
                           params + L.b, GEMM_STORE, 1, act.p, act.ld, st));
  } else {
    if (c->tf32) {
      int lda = L.in;
      cudaError_t pe;
      const float* xa = tf32_operand(c, x, rows, L.in, 0, &lda, st, &pe);
      if (pe != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("pad_rows_f32: ") + cudaGetErrorString(pe));
      if (xa != nullptr) {
        const cudaError_t e = launch_gemm_tf32_tc(false, false, rows, L.out, L.in, xa, lda, c->shadow_f[slot] + L.sb, L.ld_b, y, L.out,
                                                  params + L.b, GEMM_STORE, 1, st);
        if (e != cudaErrorNotSupported) { CK(e); return REPPO_OK; }
      }
    }
    CK(launch_gemm_f32(false, true, rows, L.out, L.in, x, L.in, params + L.w, L.in, y, L.out, params + L.b, GEMM_STORE, 1, st));
  }
  return REPPO_OK;
}
// dx[rows,in] (=|+=) dy[rows,out] W
int linear_dgrad(reppo_ctx* c, const Linear& L, const float* params, int slot, const float* dy, int rows, float* dx, int mode,
                 cudaStream_t st) {
  if (c->tc) {
    const Twin td = c->twin(dy);
    if (td.p == nullptr) return fail(c, REPPO_ERR_INVALID, "internal: dgrad operand has no bf16 twin");
    // B = W [out, in] read MN-major (reduction dim = out is the outer dim): no transposed shadow needed
    CK(launch_gemm_bf16_tc(false, true, rows, L.in, L.out, td.p, td.ld, c->shadow_b[slot] + L.sb, L.ld_b, dx, L.in, nullptr, mode,
                           1, nullptr, 0, st));
  } else {
    if (c->tf32) {
      int lda = L.out;
      cudaError_t pe;
      const float* da = tf32_operand(c, dy, rows, L.out, 0, &lda, st, &pe);
      if (pe != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("pad_rows_f32: ") + cudaGetErrorString(pe));
      if (da != nullptr) {
        const cudaError_t e = launch_gemm_tf32_tc(false, true, rows, L.in, L.out, da, lda, c->shadow_f[slot] + L.sb, L.ld_b, dx, L.in,
                                                  nullptr, mode, 1, st);
        if (e != cudaErrorNotSupported) { CK(e); return REPPO_OK; }
      }
    }
    CK(launch_gemm_f32(false, false, rows, L.in, L.out, dy, L.out, params + L.w, L.in, dx, L.in, nullptr, mode, 1, st));
  }
  return REPPO_OK;
}
// dW[out,in] += dy[rows,out]^T x[rows,in]   (dW zeroed at the start of the step)
int linear_wgrad(reppo_ctx* c, const Linear& L, const float* dy, const float* x, int rows, float* dw, cudaStream_t st) {
  const int split = wgrad_split(L.out, L.in, rows, c->tc || c->tf32);
  if (c->tf32) {
    int lda = L.out, ldb = L.in;
    cudaError_t pe, pe2;
    const float* da = tf32_operand(c, dy, rows, L.out, 0, &lda, st, &pe);
    const float* xb = tf32_operand(c, x, rows, L.in, 1, &ldb, st, &pe2);
    if (pe != cudaSuccess || pe2 != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "pad_rows_f32 failed");
    if (da != nullptr && xb != nullptr) {
      const cudaError_t e = launch_gemm_tf32_tc(true, true, L.out, L.in, rows, da, lda, xb, ldb, dw, L.in, nullptr, GEMM_ATOMIC, split, st);
      if (e != cudaErrorNotSupported) { CK(e); return REPPO_OK; }
    }
  }
  if (c->tc) {
    const Twin td = c->twin(dy), tx = c->twin(x);
    if (td.p == nullptr || tx.p == nullptr) return fail(c, REPPO_ERR_INVALID, "internal: wgrad operand has no bf16 twin");
    CK(launch_gemm_bf16_tc(true, true, L.out, L.in, rows, td.p, td.ld, tx.p, tx.ld, dw, L.in, nullptr, GEMM_ATOMIC, split, nullptr, 0, st));
  } else {
    CK(launch_gemm_f32(true, false, L.out, L.in, rows, dy, L.out, x, L.in, dw, L.in, nullptr, GEMM_ATOMIC, split, st));
  }
  return REPPO_OK;
}

// refresh the operand shadows of one network's weights: bf16 (tcgen05 bf16 mode) or aligned fp32 (tf32 mode); no-op in fp32 mode
int pack_shadows(reppo_ctx* c, int slot, const float* params, cudaStream_t st) {
  if (!c->tc && !c->tf32) return REPPO_OK;
  PackTable t{};
  t.fp32 = c->tf32 ? 1 : 0;
  auto add = [&](const Mlp& m) {
    for (const Linear& L : m.layers) {
      PackEntry& e = t.e[t.n++];
      e.w = params + L.w;
      e.wb = c->tf32 ? static_cast<void*>(c->shadow_f[slot] + L.sb) : static_cast<void*>(c->shadow_b[slot] + L.sb);
      e.out = L.out; e.in = L.in; e.ld_b = L.ld_b;
    }
  };
  if (slot == SLOT_CRITIC || slot == SLOT_CRITIC_ALT) { add(c->feature); add(c->head); add(c->pred); } else { add(c->actor); }
  CK(launch_pack_weights(t, st));
  return REPPO_OK;
}

int pack_all(reppo_ctx* c, const reppo_state* s, cudaStream_t st) {
  RET(pack_shadows(c, SLOT_CRITIC, s->critic.params, st));
  RET(pack_shadows(c, SLOT_ACTOR, s->actor.params, st));
  if (s->actor_old) RET(pack_shadows(c, SLOT_ACTOR_OLD, s->actor_old, st));
  return REPPO_OK;
}

// out_act: optional bf16 twin that receives swish(output) straight from the last GEMM's epilogue
int mlp_forward(reppo_ctx* c, const Mlp& m, const float* params, int slot, const float* in, int rows, Acts& a, cudaStream_t st,
                Twin out_act = Twin{}) {
  const float* cur = in;
  const int n = static_cast<int>(m.layers.size());
  for (int i = 0; i < n; ++i) {
    const Linear& L = m.layers[i];
    const bool last = (i == n - 1);
    float* dst = last ? a.out : a.z[i];
    if (!last && c->tc && c->fuse_norm && m.norm != NORM_NONE) {
      // Linear + bias + norm + swish in ONE tcgen05 kernel (cluster-wide row statistics); falls through when the
      // shape does not fit the cluster scheme
      const Twin tx = c->twin(cur), tw = c->twin(a.x[i]);
      NormArgs na{L.g >= 0 ? params + L.g : nullptr, L.beta >= 0 ? params + L.beta : nullptr, c->sem.norm_eps, a.rstd[i], a.mean[i]};
      const cudaError_t e = launch_gemm_bf16_tc_norm(m.norm, rows, L.out, L.in, tx.p, tx.ld, c->shadow_b[slot] + L.sb, L.ld_b,
                                                     dst, L.out, params + L.b, tw.p, tw.ld, na, st);
      if (e == cudaSuccess) { ++c->launches; cur = a.x[i]; continue; }
      if (e != cudaErrorNotSupported) return fail(c, REPPO_ERR_CUDA, std::string("launch_gemm_bf16_tc_norm: ") + cudaGetErrorString(e));
    }
    RET(linear_fwd(c, L, params, slot, cur, rows, dst, last ? out_act : Twin{}, st));
    if (!last) {
      const Twin tw = c->twin(a.x[i]);
      CK(launch_norm_act_fwd(m.norm, a.z[i], L.g >= 0 ? params + L.g : nullptr, L.beta >= 0 ? params + L.beta : nullptr,
                             rows, L.out, c->sem.norm_eps, c->tc ? nullptr : a.x[i], a.rstd[i], a.mean[i], tw.p, tw.ld, st));
      cur = a.x[i];
    }
  }
  return REPPO_OK;
}

// `to` waits for everything enqueued so far on `from` (an edge of the step's DAG; captured as a graph dependency)
int dep(reppo_ctx* c, cudaStream_t from, cudaStream_t to) {
  if (from == to) return REPPO_OK;
  cudaEvent_t e = c->events[c->ev_next++ % c->events.size()];
  cudaError_t r = cudaEventRecord(e, from);
  if (r == cudaSuccess) r = cudaStreamWaitEvent(to, e, 0);
  if (r != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("stream dependency: ") + cudaGetErrorString(r));
  return REPPO_OK;
}

// The actor chain (side[2]) is the critical path of the two-chain pipeline (profiles/r01_v8_timeline_summary.txt): its kernels
// are captured from a high-priority stream and the graph is instantiated with cudaGraphInstantiateFlagUseNodePriority, so
// their CTAs are dispatched first whenever both chains have work pending.  Measured at C2 (ms per update): off 103.9, actor
// chain first 98.2; every other assignment tried lost.  REPPO_PRIO=0 disables, REPPO_PRIO=1 forces.
// one minibatch step is a chain of few-microsecond kernels (the named configurations) rather than machine-filling ones
bool small_step(const reppo_ctx* c) {
  return static_cast<long long>(c->shape.max_rows) * std::max(c->shape.critic_hidden, c->shape.actor_hidden) <= (4LL << 20);
}
bool dp_on(const reppo_ctx* c) { return c->comm[0] != nullptr || c->mc.on; }
// Both attached: NCCL serves (it measured faster: 118.0 vs 126.4 ms per C2 update on 2 GPUs, 124.6 vs 128.6 ms on 8) unless
// REPPO_ALLREDUCE=mc asks for the multicast kernel; multicast alone attached: multicast.
bool use_mc(const reppo_ctx* c) { return c->mc.on && (knobs().nccl_own == 1 || c->comm[0] == nullptr); }
// latency-bound steps only: at the wide-net shapes every kernel fills the machine and the priorities cost 4 %
// (C5: 229.7 vs 221.2 ms); with NCCL collectives inside the chains they lose as well (2 GPUs: 115.6 ms off, 120.5 on)
bool want_prio(const reppo_ctx* c) {
  const int k = knobs().prio;
  if (k >= 0) return k != 0;
  return small_step(c) && !dp_on(c);
}

int ensure_streams(reppo_ctx* c) {
  if (!c->events.empty()) return REPPO_OK;
  // side stream i: 0 = critic pred branch, 1 = critic weight gradients, 2 = actor chain, 3 = actor weight gradients
  int least = 0, greatest = 0;
  const bool prio = want_prio(c) && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && greatest < least;
  for (int i = 0; i < 4; ++i) {
    const cudaError_t e = prio ? cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, i == 2 ? greatest : least)
                               : cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaStreamCreate failed");
  }
  for (int i = 0; i < 4; ++i)
    if (cudaEventCreateWithFlags(&c->ev_c[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_a[i], cudaEventDisableTiming) != cudaSuccess)
      return fail(c, REPPO_ERR_CUDA, "cudaEventCreate failed");
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&c->ev_epoch[i], cudaEventDisableTiming) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaEventCreate failed");
  c->events.resize(64);
  for (auto& e : c->events)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaEventCreate failed");
  return REPPO_OK;
}

// d_out: gradient w.r.t. the MLP output [rows, out_last] (read only).  grads == nullptr: frozen network (dgrad only).
// dx_in != nullptr: also produce the gradient w.r.t. the MLP input (mode GEMM_STORE / GEMM_ACCUM).
// last_bias_done: the producer of d_out already accumulated the last layer's bias gradient (column sums of d_out).
// The dgrad / norm-backward chain runs on `st`; every weight-gradient GEMM only depends on one link of the chain and
// runs on `wst` (the caller joins `wst` before the optimizer).
int mlp_backward(reppo_ctx* c, const Mlp& m, const float* params, int slot, float* grads, const float* in, int rows,
                 const Acts& a, const float* d_out, float* dx_in, int dx_mode, float* bias2, float scale2, bool last_bias_done,
                 cudaStream_t st, cudaStream_t wst) {
  const int n = static_cast<int>(m.layers.size());
  const float* d = d_out;
  for (int i = n - 1; i >= 0; --i) {
    const Linear& L = m.layers[i];
    const float* xin = (i == 0) ? in : a.x[i - 1];
    if (grads != nullptr) {
      RET(dep(c, st, wst));
      RET(linear_wgrad(c, L, d, xin, rows, grads + L.w, wst));
      // hidden layers: the norm backward that produced d accumulated its column sums already
      if (i == n - 1 && !last_bias_done) CK(launch_colsum(d, rows, L.out, L.out, grads + L.b, bias2, scale2, wst));
    }
    if (i > 0) {
      const Linear& Lp = m.layers[i - 1];
      float* dz = a.dz[i - 1];
      const Twin tw = c->twin(dz);
      RET(linear_dgrad(c, L, params, slot, d, rows, a.tmp, GEMM_STORE, st));
      CK(launch_norm_act_bwd(m.norm, a.tmp, nullptr, a.z[i - 1], Lp.g >= 0 ? params + Lp.g : nullptr,
                             Lp.beta >= 0 ? params + Lp.beta : nullptr, a.rstd[i - 1], a.mean[i - 1], rows, L.in,
                             c->tc ? nullptr : dz, (grads && Lp.g >= 0) ? grads + Lp.g : nullptr,
                             (grads && Lp.beta >= 0) ? grads + Lp.beta : nullptr, grads ? grads + Lp.b : nullptr, tw.p, tw.ld, st));
      d = dz;
    } else if (dx_in != nullptr) {
      RET(linear_dgrad(c, L, params, slot, d, rows, dx_in, dx_mode, st));
    }
  }
  return REPPO_OK;
}

int check_rows(reppo_ctx* c, int rows) {
  if (c->ws == nullptr) return fail(c, REPPO_ERR_WORKSPACE, "workspace not set (reppo_set_workspace)");
  if (rows <= 0 || rows > c->shape.max_rows) return fail(c, REPPO_ERR_INVALID, "rows outside (0, shape.max_rows]");
  return REPPO_OK;
}

// the buffers one evaluation of the critic trunk (encoder + categorical head) runs in
struct TrunkBufs { Acts* fa; Acts* ha; float* xs; float* dxs; float* dfeat; float* dlogits; };
TrunkBufs trunk_main(reppo_ctx* c) { return TrunkBufs{&c->fa, &c->ha, c->xs, c->dxs, c->dfeat, c->dlogits}; }
TrunkBufs trunk_actor(reppo_ctx* c) { return TrunkBufs{&c->fa2, &c->ha2, c->xs2, c->dxs_a, c->dfeat_a, c->dlogits_a}; }

// encoder + swish(features) -> xs
int critic_trunk_encoder(reppo_ctx* c, const float* cp, int cslot, const float* x0, int rows, const TrunkBufs& tb, cudaStream_t st) {
  const Twin tw = c->twin(tb.xs);
  if (c->tc && !knobs().no_fuse_act) {
    // swish(features), the input of both heads, leaves the encoder's last GEMM epilogue directly as bf16
    RET(mlp_forward(c, c->feature, cp, cslot, x0, rows, *tb.fa, st, tw));
  } else {
    RET(mlp_forward(c, c->feature, cp, cslot, x0, rows, *tb.fa, st));
    CK(launch_norm_act_fwd(NORM_NONE, tb.fa->out, nullptr, nullptr, rows, c->shape.critic_hidden, 0.f, c->tc ? nullptr : tb.xs,
                           nullptr, nullptr, tw.p, tw.ld, st));
  }
  return REPPO_OK;
}
int critic_trunk(reppo_ctx* c, const float* cp, int cslot, const float* x0, int rows, bool with_pred, const TrunkBufs& tb,
                 cudaStream_t st) {
  RET(critic_trunk_encoder(c, cp, cslot, x0, rows, tb, st));
  RET(mlp_forward(c, c->head, cp, cslot, tb.xs, rows, *tb.ha, st));
  if (with_pred) RET(mlp_forward(c, c->pred, cp, cslot, tb.xs, rows, c->pa, st));
  return REPPO_OK;
}

// gradient all-reduce of one chain (0 = critic, 1 = actor); no-op on a single GPU
int allreduce_grads(reppo_ctx* c, float* g, int64_t n, int chain, cudaStream_t st) {
  if (use_mc(c)) {
    // the exchange happens inside the clip+Adam launch that follows (multimem.ld_reduce / multimem.st, optim.cu)
    const char* gb = reinterpret_cast<const char*>(g);
    if (gb != c->mc.local + c->mc.grads_off[0] && gb != c->mc.local + c->mc.grads_off[1])
      return fail(c, REPPO_ERR_INVALID, "data-parallel step: the gradient arena is not the one registered with reppo_dp_attach_multicast");
    c->mc_pending = true;
    return REPPO_OK;
  }
  if (c->comm[0] == nullptr) return REPPO_OK;
  ncclComm_t comm = (chain == 1 && c->comm[1] != nullptr) ? c->comm[1] : c->comm[0];
  const int r = c->nccl.AllReduce(g, g, static_cast<size_t>(n), /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, st);
  ++c->launches;
  if (r != 0) return fail(c, REPPO_ERR_NCCL, std::string("ncclAllReduce: ") + c->nccl.GetErrorString(r));
  return REPPO_OK;
}

float inv_rows(const reppo_hyper* hp, int mb) {
  const int w = hp->world_size > 0 ? hp->world_size : 1;
  return 1.f / (static_cast<float>(mb) * static_cast<float>(w));
}

// cp / cslot: the critic parameters (and their bf16 shadow slot) this step READS
// prepped: the metric vector is already initialised and the gradient arena already zero (whole-update path)
// x0_pre: this minibatch's rows of [critic_obs | action], already gathered (whole-update path); nullptr = gather here
int critic_grad_impl(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int mb,
                     const reppo_hyper* hp, float* metrics, const float* cp, int cslot, cudaStream_t st, bool prepped = false,
                     float* x0_pre = nullptr) {
  RET(check_rows(c, mb));
  const reppo_shape& sh = c->shape;
  float* cg = s->critic.grads;
  const float inv = inv_rows(hp, mb);
  const int no_mask = (sh.semantics == REPPO_SEM_TORCH && hp->partial_reset) ? 1 : 0;
  if (c->sem.reward_head && b->reward == nullptr) return fail(c, REPPO_ERR_INVALID, "batch.reward is required in JAX semantics");
  RET(ensure_streams(c));
  cudaStream_t s1 = c->concurrent ? c->side[0] : st;   // pred-head branch
  cudaStream_t sw = c->concurrent ? c->side[1] : st;   // weight gradients
  const int Hc = sh.critic_hidden;
  if (!prepped) {
    CK(launch_metrics_init(metrics, kCriticMetricMask, st));
    CK(cudaMemsetAsync(cg, 0, static_cast<size_t>(c->critic_count) * sizeof(float), st));
  }
  float* x0 = x0_pre ? x0_pre : c->x0;
  Twin tw = c->twin(x0);
  // encoder, then swish(features) feeds two independent heads
  if (x0_pre == nullptr)
    CK(launch_gather_concat(b->critic_obs, sh.critic_obs_dim, b->action, sh.action_dim, idx, 0, mb, x0, c->K0, tw.p, tw.ld, st));
  RET(critic_trunk_encoder(c, cp, cslot, x0, mb, trunk_main(c), st));
  RET(dep(c, st, s1));
  // --- branch s1: prediction head forward, aux loss, backward down to d swish(features)
  RET(mlp_forward(c, c->pred, cp, cslot, c->xs, mb, c->pa, s1));
  tw = c->twin(c->dpred);
  CK(launch_critic_aux(c->pa.out, c->pred_out, Hc, c->sem.reward_head, c->sem.aux_mask_done, b->next_emb, b->reward, b->done,
                       b->truncated, idx, mb, inv, no_mask, hp->aux_loss_mult, c->dpred, metrics, tw.p, tw.ld,
                       cg + c->pred.layers.back().b, s1));
  RET(mlp_backward(c, c->pred, cp, cslot, cg, c->xs, mb, c->pa, c->dpred, c->dxs2, GEMM_STORE, nullptr, 0.f, true, s1, sw));
  // --- main: categorical head forward, HL-Gauss cross-entropy, backward (the loss kernels also accumulate the
  //     bias / zero_dist gradients of the two last layers)
  RET(mlp_forward(c, c->head, cp, cslot, c->xs, mb, c->ha, st));
  tw = c->twin(c->dlogits);
  CK(launch_critic_ce(c->ha.out, sh.num_bins, cp + c->zero_dist_off, c->hl, b->target, b->truncated, idx, mb, inv, no_mask,
                      c->dlogits, metrics, tw.p, tw.ld, cg + c->head.layers.back().b, cg + c->zero_dist_off, st));
  RET(mlp_backward(c, c->head, cp, cslot, cg, c->xs, mb, c->ha, c->dlogits, c->dxs, GEMM_STORE, nullptr, 0.f, true, st, sw));
  // d feat = (d swish(feat) from both heads) * swish'(feat); its column sums are the bias gradient of the encoder's last Linear
  RET(dep(c, s1, st));
  tw = c->twin(c->dfeat);
  CK(launch_norm_act_bwd(NORM_NONE, c->dxs, c->dxs2, c->fa.out, nullptr, nullptr, nullptr, nullptr, mb, Hc,
                         c->tc ? nullptr : c->dfeat, nullptr, nullptr, cg + c->feature.layers.back().b, tw.p, tw.ld, st));
  RET(mlp_backward(c, c->feature, cp, cslot, cg, x0, mb, c->fa, c->dfeat, nullptr, GEMM_STORE, nullptr, 0.f, true, st, sw));
  RET(dep(c, sw, st));
  return REPPO_OK;
}

ActorConsts actor_consts(const reppo_ctx* c, const reppo_hyper* hp) {
  ActorConsts ac{};
  ac.min_std = c->sem.force_min_std >= 0.f ? c->sem.force_min_std : hp->actor_min_std;
  ac.action_clip = c->sem.action_clip;
  ac.kl_bound = hp->kl_bound;
  ac.target_entropy = static_cast<float>(c->shape.action_dim) * hp->ent_target_mult;
  const bool torch_sem = c->shape.semantics == REPPO_SEM_TORCH;
  ac.reduce_kl = (torch_sem || hp->reduce_kl) ? 1.f : 0.f;
  ac.clip_policy_sample = c->sem.clip_policy_sample;
  ac.old_logp_at_clipped = c->sem.old_logp_at_clipped;
  ac.kl_mode = hp->kl_clip_mode;
  ac.update_entropy = (torch_sem || hp->update_entropy_lagrangian) ? 1 : 0;
  ac.update_kl = (torch_sem || hp->update_kl_lagrangian) ? 1 : 0;
  ac.reverse_kl = (!torch_sem && hp->reverse_kl) ? 1 : 0;
  return ac;
}

// old_table: optional [S, 2A] behaviour-policy outputs for every batch row (precomputed once per update)
int actor_grad_impl(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int mb,
                    const float* eps_pi, const float* eps_kl, const reppo_hyper* hp, float* metrics, const float* old_table,
                    cudaStream_t prefix_stream, const float* cp, int cslot, cudaEvent_t critic_ready, cudaEvent_t critic_read_done,
                    cudaStream_t st, bool prepped = false, float* xa0_pre = nullptr, float* x0a_pre = nullptr) {
  RET(check_rows(c, mb));
  if (hp->kl_clip_mode < 0 || hp->kl_clip_mode > 2) return fail(c, REPPO_ERR_INVALID, "unknown actor_kl_clip_mode");
  const reppo_shape& sh = c->shape;
  const int A = sh.action_dim, Dc = sh.critic_obs_dim;
  const TrunkBufs tb = trunk_actor(c);
  const float* ap = s->actor.params;
  float* ag = s->actor.grads;
  const float inv = inv_rows(hp, mb);
  const ActorConsts ac = actor_consts(c, hp);
  RET(ensure_streams(c));
  cudaStream_t sw = c->concurrent ? c->side[3] : st;
  // The actor-only prefix (policy forward, sampling, KL statistics) does not depend on the critic update that precedes
  // it on `st`: when `pre` is a different stream it overlaps the critic step (the caller forked it before that step).
  cudaStream_t pre = prefix_stream ? prefix_stream : st;
  if (!prepped) {
    CK(launch_metrics_init(metrics, kActorMetricMask, pre));
    CK(cudaMemsetAsync(ag, 0, static_cast<size_t>(c->actor_count) * sizeof(float), pre));
  }
  // xa0_pre / x0a_pre: this minibatch's rows of obs and of [critic_obs | .], already gathered (whole-update path)
  float* xa0 = xa0_pre ? xa0_pre : c->xa0;
  float* x0a = x0a_pre ? x0a_pre : c->x0a;
  Twin tw = c->twin(xa0);
  const Twin tx0 = c->twin(x0a);
  // ---- policy forward: policy input and the observation part of the critic input gathered by one launch
  if (xa0_pre == nullptr)
    CK(launch_gather_pair(b->obs, sh.obs_dim, b->critic_obs, Dc, idx, mb, xa0, sh.obs_dim, tw.p, tw.ld, x0a, c->K0, tx0.p, tx0.ld, pre));
  RET(mlp_forward(c, c->actor, ap, SLOT_ACTOR, xa0, mb, c->aa, pre));
  if (old_table == nullptr) RET(mlp_forward(c, c->actor, s->actor_old, SLOT_ACTOR_OLD, xa0, mb, c->oa, pre));
  const float* old_out = old_table ? old_table : c->oa.out;
  const int32_t* old_idx = old_table ? idx : nullptr;
  CK(launch_actor_sample(c->aa.out, old_out, old_idx, eps_pi, eps_kl, mb, A, sh.kl_samples, ac, x0a + Dc, c->K0, c->logp, c->kl,
                         c->gmu, c->gsig, tx0.p ? tx0.p + Dc : nullptr, tx0.ld, pre));
  RET(dep(c, pre, st));
  // ---- Q(s, a) through the (already updated, frozen) critic and dQ/da
  if (critic_ready != nullptr && cudaStreamWaitEvent(st, critic_ready, 0) != cudaSuccess)
    return fail(c, REPPO_ERR_CUDA, "cudaStreamWaitEvent failed");
  RET(critic_trunk(c, cp, cslot, x0a, mb, false, tb, st));
  tw = c->twin(tb.dlogits);
  CK(launch_actor_q(tb.ha->out, sh.num_bins, cp + c->zero_dist_off, c->hl, c->kl, mb, inv, ac.kl_mode, ac.kl_bound, c->q,
                    tb.dlogits, tw.p, tw.ld, st));
  RET(mlp_backward(c, c->head, cp, cslot, nullptr, tb.xs, mb, *tb.ha, tb.dlogits, tb.dxs, GEMM_STORE, nullptr, 0.f, false, st, st));
  tw = c->twin(tb.dfeat);
  CK(launch_norm_act_bwd(NORM_NONE, tb.dxs, nullptr, tb.fa->out, nullptr, nullptr, nullptr, nullptr, mb, sh.critic_hidden,
                         c->tc ? nullptr : tb.dfeat, nullptr, nullptr, nullptr, tw.p, tw.ld, st));
  RET(mlp_backward(c, c->feature, cp, cslot, nullptr, x0a, mb, *tb.fa, tb.dfeat, c->dx0, GEMM_STORE, nullptr, 0.f, false, st, st));
  // last read of the critic snapshot: the critic chain may overwrite it from here on
  if (critic_read_done != nullptr && cudaEventRecord(critic_read_done, st) != cudaSuccess)
    return fail(c, REPPO_ERR_CUDA, "cudaEventRecord failed");
  // ---- policy gradient
  tw = c->twin(c->dout);
  CK(launch_actor_grad(c->aa.out, eps_pi, x0a + Dc, c->K0, c->dx0 + Dc, c->K0, c->logp, c->kl, c->q, c->gmu, c->gsig, ap,
                       mb, A, inv, ac, c->dout, ag, metrics, tw.p, tw.ld, ag + c->actor.layers.back().b, st));
  RET(mlp_backward(c, c->actor, ap, SLOT_ACTOR, ag, xa0, mb, c->aa, c->dout, nullptr, GEMM_STORE, nullptr, 0.f, true, st, sw));
  RET(dep(c, sw, st));
  return REPPO_OK;
}

// slot < 0: plain arena (no bf16 shadow to refresh)
// p_out: where the updated parameters go (nullptr = in place); partials: scratch of the global-norm reduction
// chain: which grid-barrier words the kernel uses (0 critic chain, 1 actor chain, 2 stand-alone) -- launches that may run
// concurrently must not share them
int clip_adam_impl(reppo_ctx* c, const reppo_net_state* net, int64_t n, const reppo_hyper* hp, float* gn_out, int slot,
                   const float* p_in, float* p_out, float* partials, int chain, cudaStream_t st, bool zero_grads = false) {
  if (c->ws == nullptr) return fail(c, REPPO_ERR_WORKSPACE, "workspace not set (reppo_set_workspace)");
  ShadowTable sh{};
  if ((c->tc || c->tf32) && slot >= 0) {
    sh.fp32 = c->tf32 ? 1 : 0;
    auto add = [&](const Mlp& m) {
      for (const Linear& L : m.layers) {
        ShadowEntry& e = sh.e[sh.n++];
        e.start = L.w; e.end = L.w + static_cast<int64_t>(L.out) * L.in; e.in = L.in; e.ld = L.ld_b;
        e.dst = c->tf32 ? static_cast<void*>(c->shadow_f[slot] + L.sb) : static_cast<void*>(c->shadow_b[slot] + L.sb);
      }
    };
    if (slot == SLOT_CRITIC || slot == SLOT_CRITIC_ALT) { add(c->feature); add(c->head); add(c->pred); } else { add(c->actor); }
  }
  AdamConsts a{};
  a.lr = hp->lr; a.b1 = 0.9f; a.b2 = 0.999f; a.eps = 1e-8f;
  a.weight_decay = c->sem.torch_opt ? 0.01f : 0.f;  // torch.optim.AdamW default (src/torchrl/reppo.py:527-534)
  a.max_grad_norm = hp->max_grad_norm;
  a.torch_clip = c->sem.torch_opt;
  a.anneal_steps = hp->lr_anneal_steps > 0 ? hp->lr_anneal_steps : 0;
  McArgs mc{};
  if (c->mc_pending) {
    c->mc_pending = false;
    const int64_t off = reinterpret_cast<const char*>(net->grads) - c->mc.local;
    mc.g_mc = reinterpret_cast<float*>(c->mc.mcast + off);
    for (int r = 0; r < c->mc.world; ++r)
      mc.flags[r] = reinterpret_cast<unsigned int*>(c->mc.peer[r] + c->mc.flags_off) + static_cast<size_t>(chain) * kMcFlagBlocks * kMcMaxRanks;
    // epoch words sit behind the flag rows of the three chains
    mc.epoch = reinterpret_cast<unsigned int*>(c->mc.local + c->mc.flags_off) + static_cast<size_t>(kMcChains) * kMcFlagBlocks * kMcMaxRanks + chain;
    mc.status = c->mc.status;
    mc.stamps = c->mc.stamps;
    mc.world = c->mc.world; mc.rank = c->mc.rank; mc.chain = chain;
  }
  CK(launch_clip_adam(p_in ? p_in : net->params, p_out ? p_out : net->params, net->grads, net->adam_m, net->adam_v, n, net->step,
                      partials ? partials : c->partials, c->adam_bar + 2 * chain, a, &sh, gn_out, zero_grads ? 1 : 0, st, &mc));
  if (!clip_adam_fused(&mc)) ++c->launches;   // partial sums and the update are two kernels
  return REPPO_OK;
}

int critic_step_impl(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int mb,
                     const reppo_hyper* hp, float* metrics, cudaStream_t st, bool prepped = false, float* x0_pre = nullptr) {
  RET(critic_grad_impl(c, s, b, idx, mb, hp, metrics, s->critic.params, SLOT_CRITIC, st, prepped, x0_pre));
  RET(allreduce_grads(c, s->critic.grads, c->critic_count, 0, st));
  return clip_adam_impl(c, &s->critic, c->critic_count, hp, metrics + M_CRITIC_GRAD_NORM, SLOT_CRITIC, nullptr, nullptr, nullptr, 0, st,
                        prepped);
}

int actor_step_impl(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int mb, const float* eps_pi,
                    const float* eps_kl, const reppo_hyper* hp, float* metrics, const float* old_table,
                    cudaStream_t prefix_stream, cudaStream_t st, bool prepped = false, float* xa0_pre = nullptr,
                    float* x0a_pre = nullptr) {
  RET(actor_grad_impl(c, s, b, idx, mb, eps_pi, eps_kl, hp, metrics, old_table, prefix_stream, s->critic.params, SLOT_CRITIC,
                      nullptr, nullptr, st, prepped, xa0_pre, x0a_pre));
  RET(allreduce_grads(c, s->actor.grads, c->actor_count, 1, st));
  return clip_adam_impl(c, &s->actor, c->actor_count, hp, metrics + M_ACTOR_GRAD_NORM, SLOT_ACTOR, nullptr, nullptr, c->partials2, 1, st,
                        prepped);
}

// All epochs x minibatches.  Software pipeline over steps: the critic update of step k+1 depends only on the critic
// update of step k, the actor update of step k only on the actor update of step k-1 and on the critic parameters
// AFTER critic update k.  The two chains therefore run on two streams; the critic's Adam ping-pongs between two
// copies of the critic parameters (and bf16 shadows) so that actor step k keeps reading snapshot k while critic step
// k+1 already produces snapshot k+1.  Same arithmetic, same dependencies, half the serial chain.
int run_steps(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const reppo_plan* p, const reppo_hyper* hp,
              cudaStream_t st) {
  const int A = c->shape.action_dim, mb = p->minibatch, ks = c->shape.kl_samples;
  const int64_t per_epoch = static_cast<int64_t>(p->num_mini_batches) * mb;
  RET(ensure_streams(c));
  RET(pack_all(c, s, st));
  if (p->old_policy_out != nullptr) {
    // behaviour policy on every batch row, once: max_rows-sized chunks through the actor's forward buffers
    const int D = c->shape.obs_dim, R = c->shape.max_rows;
    const Twin tw = c->twin(c->xa0);
    for (int64_t r0 = 0; r0 < b->num_rows; r0 += R) {
      const int rows = static_cast<int>(std::min<int64_t>(R, b->num_rows - r0));
      CK(launch_gather_concat(b->obs + r0 * D, D, nullptr, 0, nullptr, 0, rows, c->xa0, D, tw.p, tw.ld, st));
      RET(mlp_forward(c, c->actor, s->actor_old, SLOT_ACTOR_OLD, c->xa0, rows, c->oa, st));
      CK(cudaMemcpyAsync(p->old_policy_out + r0 * 2 * A, c->oa.out, sizeof(float) * rows * 2 * A, cudaMemcpyDeviceToDevice, st));
    }
  }
  // one NCCL communicator serves one chain: pipelining under data parallelism needs the second communicator
  const bool pipe = c->concurrent && c->pipeline && (!dp_on(c) || c->comm[1] != nullptr || use_mc(c));
  cudaStream_t sa = pipe ? c->side[2] : st;   // the actor chain
  // snapshot j of the critic parameters lives in copy j % depth
  const int depth = pipe ? 2 : 1;
  float* P[2] = {s->critic.params, pipe ? c->cp_alt : s->critic.params};
  const int S[2] = {SLOT_CRITIC, pipe ? SLOT_CRITIC_ALT : SLOT_CRITIC};
  // With per-step metric vectors the whole update is prepared once: all metric rows initialised by one kernel, both
  // gradient arenas cleared once (every Adam launch leaves its arena zeroed for the next step).
  const bool prepped = p->step_metrics != nullptr;
  if (prepped) {
    CK(launch_metrics_init_all(p->step_metrics, p->num_epochs * p->num_mini_batches, st));
    CK(cudaMemsetAsync(s->critic.grads, 0, static_cast<size_t>(c->critic_count) * sizeof(float), st));
    CK(cudaMemsetAsync(s->actor.grads, 0, static_cast<size_t>(c->actor_count) * sizeof(float), st));
  }
  if (pipe) RET(dep(c, st, sa));
  // Pre-gathered first-layer inputs: one gather per EPOCH through that epoch's permutation (three launches) instead of
  // two per step at the head of both chains; minibatch j is then rows [j mb, (j+1) mb) of the epoch buffers.  Two sets
  // alternate by epoch; set e % 2 was last read by the actor chain in epoch e - 2.
  const bool pregather = c->x0_ep[0] != nullptr && per_epoch <= c->shape.max_batch_rows;
  const int D = c->shape.obs_dim, Dc = c->shape.critic_obs_dim, K0 = c->K0;
  int64_t k = 0;
  for (int e = 0; e < p->num_epochs; ++e) {
    const int set = e & 1;
    if (pregather) {
      const int32_t* eidx = p->perms + e * per_epoch;
      const int er = static_cast<int>(per_epoch);
      if (pipe && e >= 2 && cudaStreamWaitEvent(st, c->ev_epoch[set], 0) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaStreamWaitEvent failed");
      Twin t0 = c->twin(c->x0_ep[set]), t1 = c->twin(c->xa0_ep[set]), t2 = c->twin(c->x0a_ep[set]);
      CK(launch_gather_concat(b->critic_obs, Dc, b->action, A, eidx, 0, er, c->x0_ep[set], K0, t0.p, t0.ld, st));
      CK(launch_gather_pair(b->obs, D, b->critic_obs, Dc, eidx, er, c->xa0_ep[set], D, t1.p, t1.ld, c->x0a_ep[set], K0, t2.p, t2.ld, st));
      if (pipe) RET(dep(c, st, sa));
    }
    for (int j = 0; j < p->num_mini_batches; ++j, ++k) {
      float* x0_pre = pregather ? c->x0_ep[set] + static_cast<size_t>(j) * mb * K0 : nullptr;
      float* xa0_pre = pregather ? c->xa0_ep[set] + static_cast<size_t>(j) * mb * D : nullptr;
      float* x0a_pre = pregather ? c->x0a_ep[set] + static_cast<size_t>(j) * mb * K0 : nullptr;
      const int64_t slot = p->noise_per_minibatch ? static_cast<int64_t>(e) * p->num_mini_batches + j : e;
      const float* eps_pi = p->eps_pi + slot * mb * A;
      const float* eps_kl = p->eps_kl + slot * ks * mb * A;
      const int32_t* idx = p->perms + e * per_epoch + static_cast<int64_t>(j) * mb;
      float* m = p->step_metrics ? p->step_metrics + k * M_NUM : c->scratch_metrics;
      const int cur = static_cast<int>(k % depth), nxt = pipe ? static_cast<int>((k + 1) % depth) : 0;
      if (!pipe) {
        // the actor-only prefix of this step's actor update is forked here and overlaps the critic update
        cudaStream_t pre = nullptr;
        if (c->concurrent) { pre = c->side[2]; RET(dep(c, st, pre)); }
        RET(critic_step_impl(c, s, b, idx, mb, hp, m, st, prepped, x0_pre));
        RET(actor_step_impl(c, s, b, idx, mb, eps_pi, eps_kl, hp, m, p->old_policy_out, pre, st, prepped, xa0_pre, x0a_pre));
        continue;
      }
      // ---- critic chain (stream st): reads snapshot `cur`, Adam writes snapshot `nxt`
      RET(critic_grad_impl(c, s, b, idx, mb, hp, m, P[cur], S[cur], st, prepped, x0_pre));
      RET(allreduce_grads(c, s->critic.grads, c->critic_count, 0, st));
      // the copy that receives snapshot k+1 held snapshot k+1-depth, last read by actor step k-depth
      if (k >= depth && cudaStreamWaitEvent(st, c->ev_a[(k - depth) & 3], 0) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaStreamWaitEvent failed");
      RET(clip_adam_impl(c, &s->critic, c->critic_count, hp, m + M_CRITIC_GRAD_NORM, S[nxt], P[cur], P[nxt], c->partials, 0, st, prepped));
      if (cudaEventRecord(c->ev_c[k & 3], st) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaEventRecord failed");
      // ---- actor chain (stream sa): needs snapshot `nxt`
      RET(actor_grad_impl(c, s, b, idx, mb, eps_pi, eps_kl, hp, m, p->old_policy_out, nullptr, P[nxt], S[nxt], c->ev_c[k & 3],
                          c->ev_a[k & 3], sa, prepped, xa0_pre, x0a_pre));
      RET(allreduce_grads(c, s->actor.grads, c->actor_count, 1, sa));
      RET(clip_adam_impl(c, &s->actor, c->actor_count, hp, m + M_ACTOR_GRAD_NORM, SLOT_ACTOR, nullptr, nullptr, c->partials2, 1, sa, prepped));
    }
    if (pregather && pipe && cudaEventRecord(c->ev_epoch[set], sa) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaEventRecord failed");
  }
  if (pipe) {
    RET(dep(c, sa, st));
    if (k % depth != 0) {  // the final critic parameters sit in an alternate copy
      CK(cudaMemcpyAsync(s->critic.params, P[k % depth], sizeof(float) * c->critic_count, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (p->metrics != nullptr && p->step_metrics != nullptr)
    CK(launch_metrics_mean(p->step_metrics, (p->num_epochs - 1) * p->num_mini_batches, p->num_mini_batches, M_NUM, p->metrics, st));
  return REPPO_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int reppo_version(void) { return 2; }

const char* reppo_last_error(const reppo_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int reppo_ctx_create(const reppo_shape* shape, reppo_ctx** out) {
  if (shape == nullptr || out == nullptr) return fail(nullptr, REPPO_ERR_INVALID, "null argument");
  const reppo_shape& s = *shape;
  if (s.obs_dim <= 0 || s.critic_obs_dim <= 0 || s.action_dim <= 0 || s.critic_hidden <= 0 || s.actor_hidden <= 0 ||
      s.num_bins < 2 || s.max_rows <= 0 || s.kl_samples <= 0)
    return fail(nullptr, REPPO_ERR_INVALID, "non-positive dimension in reppo_shape");
  if (s.enc_layers < 2 || s.head_layers < 2 || s.pred_layers < 2 || s.actor_layers < 2)
    return fail(nullptr, REPPO_ERR_UNSUPPORTED,
                "layers == 1 has divergent JAX / Torch semantics (jax_models.py:81-89 vs torch_models.py:98-107)");
  if (s.enc_layers + s.head_layers + s.pred_layers > kMaxPack || s.actor_layers > kMaxPack)
    return fail(nullptr, REPPO_ERR_UNSUPPORTED, "more than 16 Linear layers in one network");
  if (s.num_bins > 256) return fail(nullptr, REPPO_ERR_UNSUPPORTED, "num_bins > 256");
  if (s.critic_hidden > 2048 || s.actor_hidden > 2048) return fail(nullptr, REPPO_ERR_UNSUPPORTED, "hidden_dim > 2048");
  if (s.semantics != REPPO_SEM_JAX && s.semantics != REPPO_SEM_TORCH) return fail(nullptr, REPPO_ERR_INVALID, "semantics");
  if (s.critic_norm < 0 || s.critic_norm > 2 || s.actor_norm < 0 || s.actor_norm > 2) return fail(nullptr, REPPO_ERR_INVALID, "norm type");
  if (s.gemm_mode != REPPO_GEMM_FP32 && s.gemm_mode != REPPO_GEMM_BF16 && s.gemm_mode != REPPO_GEMM_TF32)
    return fail(nullptr, REPPO_ERR_INVALID, "gemm_mode");
  if (!(s.vmax > s.vmin)) return fail(nullptr, REPPO_ERR_INVALID, "vmax <= vmin");

  reppo_ctx* c = new reppo_ctx();
  c->shape = s;
  c->tc = (s.gemm_mode == REPPO_GEMM_BF16);
  c->tf32 = (s.gemm_mode == REPPO_GEMM_TF32);
  c->concurrent = !knobs().serial;
  c->fuse_norm = !knobs().no_fuse_norm;
  c->pipeline = !knobs().no_pipeline;
  if (s.semantics == REPPO_SEM_JAX) {
    // flax RMSNorm/LayerNorm eps 1e-6; logits + 40.0 * zero_dist (jax_models.py:298); clip 1-1e-4
    // (jaxrl/reppo.py:558); min_std 0.1 because the constructor ignores cfg (jax_models.py:336);
    // reward head + (1-done) aux mask (jaxrl/reppo.py:493-502); optax clip + adam.
    c->sem = Sem{1e-6f, 40.0f, 1.0f - 1e-4f, 0, 0, 1, 1, 0, 0.1f};
  } else {
    // nn.RMSNorm(eps=None) -> finfo(float32).eps; 40.9 (torch_models.py:282); clip 1-1e-6
    // (torchrl/reppo.py:301,308); AdamW(wd=0.01) + clip_grad_norm_ (:270-273,527-534).
    c->sem = Sem{1.1920928955078125e-07f, 40.9f, 1.0f - 1e-6f, 1, 1, 0, 0, 1, -1.f};
  }
  if (s.critic_norm == REPPO_NORM_LAYER || s.actor_norm == REPPO_NORM_LAYER) c->sem.norm_eps = 1e-6f;  // reppo_mj_playground.py:65-72
  c->K0 = s.critic_obs_dim + s.action_dim;
  c->pred_out = s.critic_hidden + (c->sem.reward_head ? 1 : 0);
  c->maxdim = std::max(std::max(s.critic_hidden, s.actor_hidden),
                       std::max(std::max(s.num_bins, c->pred_out), std::max(std::max(c->K0, s.obs_dim), 2 * s.action_dim)));

  int64_t off = 0, sb = 0;
  c->zero_dist_off = 0;
  c->critic_tensors.push_back({"zero_dist", 0, s.num_bins, 0});
  off += s.num_bins;
  add_fcnn(c->critic_tensors, off, sb, c->feature, "feature_module", c->K0, s.critic_hidden, s.critic_hidden, s.enc_layers, false, s.critic_norm);
  add_fcnn(c->critic_tensors, off, sb, c->head, "critic_module", s.critic_hidden, s.critic_hidden, s.num_bins, s.head_layers, true, s.critic_norm);
  add_fcnn(c->critic_tensors, off, sb, c->pred, "pred_module", s.critic_hidden, s.critic_hidden, c->pred_out, s.pred_layers, true, s.critic_norm);
  c->critic_count = off; c->shadow_b_elems[0] = sb;
  c->actor_tensors.push_back({"log_temp", 0, 0, 0});
  c->actor_tensors.push_back({"log_lagrange", 1, 0, 0});
  off = 2; sb = 0;
  add_fcnn(c->actor_tensors, off, sb, c->actor, "model", s.obs_dim, s.actor_hidden, 2 * s.action_dim, s.actor_layers, false, s.actor_norm);
  c->actor_count = off; c->shadow_b_elems[1] = sb;

  const double bw = (static_cast<double>(s.vmax) - static_cast<double>(s.vmin)) / (s.num_bins - 1);
  const double sigma = bw * 0.75;
  c->edges_h = linspace_f32(s.vmin - bw / 2, s.vmax + bw / 2, s.num_bins + 1);
  c->support_h = linspace_f32(s.vmin, s.vmax, s.num_bins);
  c->hl.vmin = s.vmin; c->hl.vmax = s.vmax; c->hl.num_bins = s.num_bins; c->hl.bias_scale = c->sem.bias_scale;
  if (s.semantics == REPPO_SEM_JAX) {
    c->hl.den = static_cast<float>(static_cast<double>(sqrtf(2.f)) * sigma);  // utils.py:52
    c->hl.z_eps = 0.f;
  } else {
    c->hl.den = sqrtf(2.f) * static_cast<float>(sigma) + 1e-6f;               // reppo_util.py:767-769
    c->hl.z_eps = 1e-6f;                                                      // :773
  }
  c->ws_need = layout_workspace(c, nullptr);
  *out = c;
  return REPPO_OK;
}

void reppo_ctx_destroy(reppo_ctx* c) {
  if (c == nullptr) return;
  for (int i = 0; i < reppo_ctx::kGraphSlots; ++i) if (c->graph_exec[i]) cudaGraphExecDestroy(c->graph_exec[i]);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  for (auto& e : c->events) cudaEventDestroy(e);
  for (int i = 0; i < 4; ++i) { if (c->side[i]) cudaStreamDestroy(c->side[i]); if (c->ev_c[i]) cudaEventDestroy(c->ev_c[i]); if (c->ev_a[i]) cudaEventDestroy(c->ev_a[i]); }
  for (int i = 0; i < 2; ++i) if (c->ev_epoch[i]) cudaEventDestroy(c->ev_epoch[i]);
  for (int i = 0; i < 2; ++i) if (c->comm[i] && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comm[i]);
  if (c->mc.status) cudaFree(c->mc.status);
  if (c->mc.stamps) cudaFree(c->mc.stamps);
  delete c;
}

int64_t reppo_param_count(const reppo_ctx* c, int net) { return net == REPPO_NET_CRITIC ? c->critic_count : c->actor_count; }
int32_t reppo_param_num_tensors(const reppo_ctx* c, int net) {
  return static_cast<int32_t>(net == REPPO_NET_CRITIC ? c->critic_tensors.size() : c->actor_tensors.size());
}
int reppo_param_tensor(const reppo_ctx* c, int net, int32_t index, char* name, int32_t name_cap, int64_t* offset,
                       int32_t* rows, int32_t* cols) {
  const auto& ts = net == REPPO_NET_CRITIC ? c->critic_tensors : c->actor_tensors;
  if (index < 0 || index >= static_cast<int32_t>(ts.size())) return REPPO_ERR_INVALID;
  const TensorInfo& t = ts[index];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", t.name.c_str());
  if (offset) *offset = t.off;
  if (rows) *rows = t.rows;
  if (cols) *cols = t.cols;
  return REPPO_OK;
}

size_t reppo_workspace_bytes(const reppo_ctx* c) { return c->ws_need; }

int reppo_set_bins(reppo_ctx* c, const float* edges_host, const float* support_host) {
  // optional: host-computed torch.linspace arrays (bit-identical to the reference's own support / bin edges)
  if (edges_host) c->edges_h.assign(edges_host, edges_host + c->shape.num_bins + 1);
  if (support_host) c->support_h.assign(support_host, support_host + c->shape.num_bins);
  if (c->ws) {
    if (cudaMemcpy(c->edges_d, c->edges_h.data(), c->edges_h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->support_d, c->support_h.data(), c->support_h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
      return fail(c, REPPO_ERR_CUDA, "uploading bin edges failed");
  }
  return REPPO_OK;
}

int reppo_set_workspace(reppo_ctx* c, void* workspace, size_t bytes) {
  if (workspace == nullptr || bytes < c->ws_need) return fail(c, REPPO_ERR_WORKSPACE, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return fail(c, REPPO_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  c->ws = static_cast<char*>(workspace);
  c->ws_bytes = bytes;
  layout_workspace(c, c->ws);
  c->hl.edges = c->edges_d;
  c->hl.support = c->support_d;
  c->invalidate_graphs();
  // bf16 twins / shadows are read through TMA with their padded leading dimension (clear the padding once), and the
  // optimizer's grid-barrier words start at zero
  if (cudaMemset(workspace, 0, c->ws_need) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "clearing the workspace failed");
  return reppo_set_bins(c, nullptr, nullptr);
}

int reppo_td_lambda(const float* soft_reward, const float* value, const float* done, float* truncated,
                    const float* importance_weight, float* target, int32_t T, int64_t N, double gamma, double lmbda,
                    int32_t convention, int32_t mutate_truncated, reppo_stream stream) {
  if (!soft_reward || !value || !done || !truncated || !target || T <= 0 || N <= 0) return REPPO_ERR_INVALID;
  if (convention != REPPO_SEM_JAX && convention != REPPO_SEM_TORCH) return REPPO_ERR_INVALID;
  const cudaError_t e = launch_td_lambda(soft_reward, value, done, truncated, importance_weight, target, T, N, gamma, lmbda,
                                         convention, mutate_truncated, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { g_create_error = std::string("td_lambda: ") + cudaGetErrorString(e); return REPPO_ERR_CUDA; }
  return REPPO_OK;
}

int reppo_critic_forward(reppo_ctx* c, const float* cp, const float* critic_obs, const float* action, int32_t rows,
                         float* value, float* logits, float* pred, float* features, reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RET(check_rows(c, rows));
  c->launches = 0;
  const reppo_shape& sh = c->shape;
  RET(pack_shadows(c, SLOT_CRITIC, cp, st));
  const Twin tw = c->twin(c->x0);
  CK(launch_gather_concat(critic_obs, sh.critic_obs_dim, action, sh.action_dim, nullptr, 0, rows, c->x0, c->K0, tw.p, tw.ld, st));
  RET(critic_trunk(c, cp, SLOT_CRITIC, c->x0, rows, pred != nullptr, trunk_main(c), st));
  if (value || logits) CK(launch_value_head(c->ha.out, sh.num_bins, cp + c->zero_dist_off, c->hl, rows, value, logits, st));
  if (pred) CK(cudaMemcpyAsync(pred, c->pa.out, sizeof(float) * rows * c->pred_out, cudaMemcpyDeviceToDevice, st));
  if (features) CK(cudaMemcpyAsync(features, c->fa.out, sizeof(float) * rows * sh.critic_hidden, cudaMemcpyDeviceToDevice, st));
  return REPPO_OK;
}

int reppo_actor_forward(reppo_ctx* c, const float* ap, const float* obs, int32_t rows, float min_std, float* mean, float* std,
                        reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RET(check_rows(c, rows));
  c->launches = 0;
  RET(pack_shadows(c, SLOT_ACTOR, ap, st));
  const Twin tw = c->twin(c->xa0);
  CK(launch_gather_concat(obs, c->shape.obs_dim, nullptr, 0, nullptr, 0, rows, c->xa0, c->shape.obs_dim, tw.p, tw.ld, st));
  RET(mlp_forward(c, c->actor, ap, SLOT_ACTOR, c->xa0, rows, c->aa, st));
  const float ms = c->sem.force_min_std >= 0.f ? c->sem.force_min_std : min_std;
  CK(launch_mean_std(c->aa.out, rows, c->shape.action_dim, ms, mean, std, st));
  return REPPO_OK;
}

int reppo_linear(reppo_ctx* c, int32_t op, int32_t rows, int32_t in_f, int32_t out_f, const float* a, const float* b,
                 float* out, const float* bias, reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool staged = (op & 0x100) != 0;   // bf16 operands of the previous identical call are still staged: GEMM only
  op &= 0xff;
  if (!a || !b || !out || rows <= 0 || in_f <= 0 || out_f <= 0 || op < 0 || op > 2)
    return fail(c, REPPO_ERR_INVALID, "reppo_linear: bad argument");
  c->launches = 0;
  if (c->tf32) {
    // fp32 operands read as tf32 straight from the caller's buffers when TMA can address them, else from row-padded copies
    cudaError_t e = cudaErrorNotSupported, p1 = cudaSuccess, p2 = cudaSuccess;
    int la = 0, lb = 0;
    if (op == 0) {          // a = x [rows, in], b = w [out, in]
      const float* xa = tf32_operand(c, a, rows, in_f, 0, &la, st, &p1);
      const float* wb = tf32_operand(c, b, out_f, in_f, 1, &lb, st, &p2);
      if (xa && wb) e = launch_gemm_tf32_tc(false, false, rows, out_f, in_f, xa, la, wb, lb, out, out_f, bias, GEMM_STORE, 1, st);
    } else if (op == 1) {   // a = dy [rows, out], b = w [out, in]
      const float* da = tf32_operand(c, a, rows, out_f, 0, &la, st, &p1);
      const float* wb = tf32_operand(c, b, out_f, in_f, 1, &lb, st, &p2);
      if (da && wb) e = launch_gemm_tf32_tc(false, true, rows, in_f, out_f, da, la, wb, lb, out, in_f, nullptr, GEMM_STORE, 1, st);
    } else {                // a = dy [rows, out], b = x [rows, in]
      const float* da = tf32_operand(c, a, rows, out_f, 0, &la, st, &p1);
      const float* xb = tf32_operand(c, b, rows, in_f, 1, &lb, st, &p2);
      if (da && xb) e = launch_gemm_tf32_tc(true, true, out_f, in_f, rows, da, la, xb, lb, out, in_f, nullptr, GEMM_ATOMIC,
                                            wgrad_split(out_f, in_f, rows, true), st);
    }
    if (p1 != cudaSuccess || p2 != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "pad_rows_f32 failed");
    if (e != cudaErrorNotSupported) { CK(e); return REPPO_OK; }
  }
  if (!c->tc) {
    if (op == 0) CK(launch_gemm_f32(false, true, rows, out_f, in_f, a, in_f, b, in_f, out, out_f, bias, GEMM_STORE, 1, st));
    else if (op == 1) CK(launch_gemm_f32(false, false, rows, in_f, out_f, a, out_f, b, in_f, out, in_f, nullptr, GEMM_STORE, 1, st));
    else CK(launch_gemm_f32(true, false, out_f, in_f, rows, a, out_f, b, in_f, out, in_f, nullptr, GEMM_ATOMIC,
                            wgrad_split(out_f, in_f, rows, false), st));
    return REPPO_OK;
  }
  RET(check_rows(c, rows));
  if (in_f > c->maxdim || out_f > c->maxdim) return fail(c, REPPO_ERR_INVALID, "reppo_linear (bf16): dimension exceeds the context's largest layer");
  const int ldi = round8(in_f), ldo = round8(out_f);
  if (op == 0) {
    if (!staged) {
      CK(launch_cast_bf16(a, rows, in_f, in_f, c->lin_a, ldi, st));
      CK(launch_cast_bf16(b, out_f, in_f, in_f, c->lin_b, ldi, st));
    }
    CK(launch_gemm_bf16_tc(false, false, rows, out_f, in_f, c->lin_a, ldi, c->lin_b, ldi, out, out_f, bias, GEMM_STORE, 1, nullptr, 0, st));
  } else if (op == 1) {
    if (!staged) {
      CK(launch_cast_bf16(a, rows, out_f, out_f, c->lin_a, ldo, st));
      CK(launch_cast_bf16(b, out_f, in_f, in_f, c->lin_b, ldi, st));
    }
    CK(launch_gemm_bf16_tc(false, true, rows, in_f, out_f, c->lin_a, ldo, c->lin_b, ldi, out, in_f, nullptr, GEMM_STORE, 1, nullptr, 0, st));
  } else {
    if (!staged) {
      CK(launch_cast_bf16(a, rows, out_f, out_f, c->lin_a, ldo, st));
      CK(launch_cast_bf16(b, rows, in_f, in_f, c->lin_c, ldi, st));
    }
    CK(launch_gemm_bf16_tc(true, true, out_f, in_f, rows, c->lin_a, ldo, c->lin_c, ldi, out, in_f, nullptr, GEMM_ATOMIC,
                           wgrad_split(out_f, in_f, rows, true), nullptr, 0, st));
  }
  return REPPO_OK;
}


int reppo_kernel_probe(reppo_ctx* c, int32_t which, int32_t rows, const float* in0, const float* in1, const float* in2,
                       const float* in3, float* out0, float* scratch, const reppo_hyper* hp, reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!c || !in0 || !in1 || !in2 || !in3 || !out0 || !scratch || !hp || rows <= 0 || which < 0 || which > 2)
    return fail(c, REPPO_ERR_INVALID, "reppo_kernel_probe: bad argument");
  const reppo_shape& sh = c->shape;
  const float inv = 1.f / static_cast<float>(rows);
  c->launches = 0;
  if (which == 0) {
    const int B = sh.num_bins;
    CK(launch_critic_ce(in0, B, in3, c->hl, in1, in2, nullptr, rows, inv, 0, out0, scratch, nullptr, 0, scratch + M_NUM,
                        scratch + M_NUM + B, st));
  } else if (which == 1) {
    const int H = sh.critic_hidden;
    CK(launch_critic_aux(in0, c->pred_out, H, c->sem.reward_head, c->sem.aux_mask_done, in1, in2, in3, in3, nullptr, rows, inv, 0,
                         hp->aux_loss_mult, out0, scratch, nullptr, 0, scratch + M_NUM, st));
  } else {
    const int A = sh.action_dim;
    const ActorConsts ac = actor_consts(c, hp);
    float* logp = scratch;
    float* kl = scratch + rows;
    float* gmu = scratch + 2 * static_cast<size_t>(rows);
    float* gsig = gmu + static_cast<size_t>(rows) * A;
    CK(launch_actor_sample(in0, in1, nullptr, in2, in3, rows, A, sh.kl_samples, ac, out0, A, logp, kl, gmu, gsig, nullptr, 0, st));
  }
  ++c->launches;
  return REPPO_OK;
}

// ---- rollout side -----------------------------------------------------------------------------
int reppo_obs_normalize(const float* x, int64_t rows, int32_t dim, float* mean, float* var, float* std, int64_t* count, float eps,
                        int32_t update, int32_t center, int64_t until, float* out, reppo_stream stream) {
  if (!x || !mean || !var || !std || !out || rows <= 0 || dim <= 0 || (update && !count)) return REPPO_ERR_INVALID;
  const cudaError_t e = launch_obs_normalize(x, rows, dim, mean, var, std, reinterpret_cast<long long*>(count), eps, update, center,
                                             until, out, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { g_create_error = std::string("obs_normalize: ") + cudaGetErrorString(e); return REPPO_ERR_CUDA; }
  return REPPO_OK;
}

static int policy_act_impl(reppo_ctx* c, const float* ap, const float* obs, const float* eps, const float* scale, int rows,
                           float min_std, float clip, int store_clipped, float* action, int act_ld, void* act_h, int act_ldh,
                           float* logp, float* logp_scaled, cudaStream_t st) {
  RET(pack_shadows(c, SLOT_ACTOR, ap, st));
  const Twin tw = c->twin(c->xa0);
  CK(launch_gather_concat(obs, c->shape.obs_dim, nullptr, 0, nullptr, 0, rows, c->xa0, c->shape.obs_dim, tw.p, tw.ld, st));
  RET(mlp_forward(c, c->actor, ap, SLOT_ACTOR, c->xa0, rows, c->aa, st));
  const float ms = c->sem.force_min_std >= 0.f ? c->sem.force_min_std : min_std;
  CK(launch_policy_sample(c->aa.out, eps, scale, rows, c->shape.action_dim, ms, clip, store_clipped, action, act_ld, act_h, act_ldh,
                          logp, logp_scaled, st));
  return REPPO_OK;
}

int reppo_policy_act(reppo_ctx* c, const float* actor_params, const float* obs, const float* eps, const float* scale, int32_t rows,
                     float min_std, float clip, int32_t store_clipped, float* action, float* log_prob, float* log_prob_scaled,
                     reppo_stream stream) {
  if (!actor_params || !obs || !eps || !action) return fail(c, REPPO_ERR_INVALID, "reppo_policy_act: null argument");
  RET(check_rows(c, rows));
  c->launches = 0;
  return policy_act_impl(c, actor_params, obs, eps, scale, rows, min_std, clip, store_clipped, action, c->shape.action_dim, nullptr, 0,
                         log_prob, log_prob_scaled, static_cast<cudaStream_t>(stream));
}

int reppo_bootstrap(reppo_ctx* c, const float* actor_params, const float* critic_params, const float* next_obs,
                    const float* next_critic_obs, const float* eps, const float* reward, int32_t rows, float gamma, float min_std,
                    float clip, float* next_value, float* next_emb, float* soft_reward, float* next_action, float* next_log_prob,
                    reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!actor_params || !critic_params || !next_obs || !next_critic_obs || !eps)
    return fail(c, REPPO_ERR_INVALID, "reppo_bootstrap: null argument");
  if (soft_reward != nullptr && reward == nullptr) return fail(c, REPPO_ERR_INVALID, "reppo_bootstrap: soft_reward needs reward");
  RET(check_rows(c, rows));
  c->launches = 0;
  const reppo_shape& sh = c->shape;
  const int Dc = sh.critic_obs_dim, A = sh.action_dim;
  // critic input [next_critic_obs | next_action]: the observation part is gathered, the sampled action lands beside it
  const Twin tx = c->twin(c->x0);
  CK(launch_gather_concat(next_critic_obs, Dc, nullptr, 0, nullptr, 0, rows, c->x0, c->K0, tx.p, tx.ld, st));
  RET(policy_act_impl(c, actor_params, next_obs, eps, nullptr, rows, min_std, clip, 0, c->x0 + Dc, c->K0,
                      tx.p ? tx.p + Dc : nullptr, tx.ld, c->logp, nullptr, st));
  RET(pack_shadows(c, SLOT_CRITIC, critic_params, st));
  RET(critic_trunk(c, critic_params, SLOT_CRITIC, c->x0, rows, false, trunk_main(c), st));
  if (next_value) CK(launch_value_head(c->ha.out, sh.num_bins, critic_params + c->zero_dist_off, c->hl, rows, next_value, nullptr, st));
  if (next_emb) CK(cudaMemcpyAsync(next_emb, c->fa.out, sizeof(float) * rows * sh.critic_hidden, cudaMemcpyDeviceToDevice, st));
  if (soft_reward) CK(launch_soft_reward(reward, c->logp, actor_params, gamma, rows, soft_reward, st));
  if (next_action) CK(cudaMemcpy2DAsync(next_action, sizeof(float) * A, c->x0 + Dc, sizeof(float) * c->K0, sizeof(float) * A, rows,
                                        cudaMemcpyDeviceToDevice, st));
  if (next_log_prob) CK(cudaMemcpyAsync(next_log_prob, c->logp, sizeof(float) * rows, cudaMemcpyDeviceToDevice, st));
  return REPPO_OK;
}

int reppo_critic_grad(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int32_t mb,
                      const reppo_hyper* hp, float* metrics, reppo_stream stream) {
  c->launches = 0;
  RET(pack_all(c, s, static_cast<cudaStream_t>(stream)));
  return critic_grad_impl(c, s, b, idx, mb, hp, metrics, s->critic.params, SLOT_CRITIC, static_cast<cudaStream_t>(stream));
}
int reppo_actor_grad(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int32_t mb,
                     const float* eps_pi, const float* eps_kl, const reppo_hyper* hp, float* metrics, reppo_stream stream) {
  c->launches = 0;
  RET(pack_all(c, s, static_cast<cudaStream_t>(stream)));
  return actor_grad_impl(c, s, b, idx, mb, eps_pi, eps_kl, hp, metrics, nullptr, nullptr, s->critic.params, SLOT_CRITIC, nullptr,
                         nullptr, static_cast<cudaStream_t>(stream));
}
int reppo_critic_step(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int32_t mb,
                      const reppo_hyper* hp, float* metrics, reppo_stream stream) {
  c->launches = 0;
  RET(pack_all(c, s, static_cast<cudaStream_t>(stream)));
  return critic_step_impl(c, s, b, idx, mb, hp, metrics, static_cast<cudaStream_t>(stream));
}
int reppo_actor_step(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const int32_t* idx, int32_t mb,
                     const float* eps_pi, const float* eps_kl, const reppo_hyper* hp, float* metrics, reppo_stream stream) {
  c->launches = 0;
  RET(pack_all(c, s, static_cast<cudaStream_t>(stream)));
  return actor_step_impl(c, s, b, idx, mb, eps_pi, eps_kl, hp, metrics, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}
int reppo_clip_adam(reppo_ctx* c, const reppo_net_state* net, int64_t n, const reppo_hyper* hp, float* gn, reppo_stream stream) {
  c->launches = 0;
  return clip_adam_impl(c, net, n, hp, gn, -1, nullptr, nullptr, nullptr, 2, static_cast<cudaStream_t>(stream));
}

int reppo_update(reppo_ctx* c, const reppo_state* s, const reppo_batch* b, const reppo_plan* p, const reppo_hyper* hp,
                 reppo_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!s || !b || !p || !hp) return fail(c, REPPO_ERR_INVALID, "null argument");
  if (p->num_epochs <= 0 || p->num_mini_batches <= 0 || !p->perms || !p->eps_pi || !p->eps_kl)
    return fail(c, REPPO_ERR_INVALID, "incomplete reppo_plan");
  RET(check_rows(c, p->minibatch));
  if (!p->use_graph) {
    c->launches = 0;
    return run_steps(c, s, b, p, hp, st);
  }
  GraphKey key{};
  key.state = *s; key.batch = *b; key.plan = *p; key.hp = *hp;
  int slot = -1;
  for (int i = 0; i < reppo_ctx::kGraphSlots; ++i)
    if (c->graph_valid[i] && memcmp(&key, &c->graph_key[i], sizeof(GraphKey)) == 0) slot = i;
  if (slot < 0) {
    slot = c->graph_next;
    c->graph_next = (c->graph_next + 1) % reppo_ctx::kGraphSlots;
    if (c->graph_exec[slot]) { cudaGraphExecDestroy(c->graph_exec[slot]); c->graph_exec[slot] = nullptr; }
    c->graph_valid[slot] = false;
    c->launches = 0;
    cudaError_t e = cudaSuccess;
    if (c->cap_stream == nullptr) {
      e = cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    e = cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
    const int r = run_steps(c, s, b, p, hp, c->cap_stream);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(c->cap_stream, &graph);
    if (r != REPPO_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    const bool use_prio = want_prio(c);   // the side streams were created with the same decision
    e = cudaGraphInstantiate(&c->graph_exec[slot], graph, use_prio ? cudaGraphInstantiateFlagUseNodePriority : 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    memcpy(&c->graph_key[slot], &key, sizeof(GraphKey));
    c->graph_valid[slot] = true;
    c->graph_launches[slot] = c->launches;
  }
  const cudaError_t e = cudaGraphLaunch(c->graph_exec[slot], st);
  if (e != cudaSuccess) return fail(c, REPPO_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
  c->launches = c->graph_launches[slot];
  return REPPO_OK;
}

int64_t reppo_launch_count(const reppo_ctx* c) { return c->launches; }

static int nccl_load(reppo_ctx* c) {
  if (c->nccl.lib) return REPPO_OK;
  const char* names[] = {knobs().nccl_lib, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (n == nullptr) continue;
    c->nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (c->nccl.lib) break;
  }
  if (!c->nccl.lib) return fail(c, REPPO_ERR_NCCL, "cannot dlopen libnccl.so.2 (set REPPO_NCCL_LIB)");
  c->nccl.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(c->nccl.lib, "ncclGetUniqueId"));
  c->nccl.CommInitRank = reinterpret_cast<int (*)(ncclComm_t*, int, NcclId, int)>(dlsym(c->nccl.lib, "ncclCommInitRank"));
  c->nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t)>(dlsym(c->nccl.lib, "ncclAllReduce"));
  c->nccl.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(dlsym(c->nccl.lib, "ncclCommDestroy"));
  c->nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(c->nccl.lib, "ncclGetErrorString"));
  if (!c->nccl.GetUniqueId || !c->nccl.CommInitRank || !c->nccl.AllReduce || !c->nccl.CommDestroy || !c->nccl.GetErrorString)
    return fail(c, REPPO_ERR_NCCL, "libnccl is missing required symbols");
  return REPPO_OK;
}

int reppo_nccl_unique_id(reppo_ctx* c, void* id_out_128) {
  RET(nccl_load(c));
  NcclId id;
  const int r = c->nccl.GetUniqueId(&id);
  if (r != 0) return fail(c, REPPO_ERR_NCCL, std::string("ncclGetUniqueId: ") + c->nccl.GetErrorString(r));
  memcpy(id_out_128, &id, sizeof(id));
  return REPPO_OK;
}

int reppo_nccl_init(reppo_ctx* c, const void* id_128, int32_t rank, int32_t world_size) {
  RET(nccl_load(c));
  NcclId id;
  memcpy(&id, id_128, sizeof(id));
  // first call: the communicator of the critic chain (and of everything when it stays the only one);
  // second call (with a fresh unique id): the actor chain's, which lets the two chains all-reduce concurrently
  const int which = (c->comm[0] == nullptr) ? 0 : 1;
  if (which == 1 && c->comm[1] != nullptr) return fail(c, REPPO_ERR_INVALID, "both communicators are already attached");
  const int r = c->nccl.CommInitRank(&c->comm[which], world_size, id, rank);
  if (r != 0) { c->comm[which] = nullptr; return fail(c, REPPO_ERR_NCCL, std::string("ncclCommInitRank: ") + c->nccl.GetErrorString(r)); }
  c->world = world_size; c->rank = rank;
  c->invalidate_graphs();
  return REPPO_OK;
}

int reppo_dp_attach_multicast(reppo_ctx* c, const reppo_multicast* m) {
  if (m == nullptr || m->local_base == nullptr || m->mc_base == nullptr || m->world_size < 2 || m->world_size > kMcMaxRanks ||
      m->rank < 0 || m->rank >= m->world_size)
    return fail(c, REPPO_ERR_INVALID, "reppo_dp_attach_multicast: bad argument");
  if (16 % m->world_size != 0)
    return fail(c, REPPO_ERR_INVALID, "reppo_dp_attach_multicast: world_size must divide 16 (warps of a clip+Adam block are dealt to ranks)");
  const int64_t need = (static_cast<int64_t>(kMcChains) * kMcFlagBlocks * kMcMaxRanks + 16) * sizeof(unsigned int);
  if (m->flags_bytes < need || (m->critic_grads_off & 15) || (m->actor_grads_off & 15) || (m->flags_off & 15))
    return fail(c, REPPO_ERR_INVALID, "reppo_dp_attach_multicast: flag block too small or offsets not 16-byte aligned");
  for (int r = 0; r < m->world_size; ++r)
    if (m->peer_base[r] == nullptr) return fail(c, REPPO_ERR_INVALID, "reppo_dp_attach_multicast: missing peer mapping");
  if (m->peer_base[m->rank] != m->local_base) return fail(c, REPPO_ERR_INVALID, "reppo_dp_attach_multicast: peer_base[rank] != local_base");
  reppo_ctx::McState& s = c->mc;
  if (s.status == nullptr) {
    if (cudaMalloc(&s.status, sizeof(int)) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaMalloc failed");
  }
  cudaMemset(s.status, 0, sizeof(int));
  if (s.stamps == nullptr && getenv("REPPO_MC_STAMPS") != nullptr) {
    if (cudaMalloc(&s.stamps, kMcStampWords * sizeof(unsigned long long)) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaMalloc failed");
    cudaMemset(s.stamps, 0, kMcStampWords * sizeof(unsigned long long));
  }
  s.local = static_cast<char*>(m->local_base); s.mcast = static_cast<char*>(m->mc_base);
  for (int r = 0; r < m->world_size; ++r) s.peer[r] = static_cast<char*>(m->peer_base[r]);
  s.grads_off[0] = m->critic_grads_off; s.grads_off[1] = m->actor_grads_off; s.flags_off = m->flags_off;
  s.world = m->world_size; s.rank = m->rank; s.on = true;
  c->world = m->world_size; c->rank = m->rank;
  c->invalidate_graphs();
  return REPPO_OK;
}

// debugging aid (REPPO_MC_STAMPS=1 at attach time): copies the stamp ring [count, 7 pad, 8192 x 8 words] to the host
int reppo_dp_stamps(reppo_ctx* c, uint64_t* out, int64_t words) {
  if (c->mc.stamps == nullptr) return fail(c, REPPO_ERR_INVALID, "reppo_dp_stamps: set REPPO_MC_STAMPS=1 before reppo_dp_attach_multicast");
  const int64_t n = std::min<int64_t>(words, kMcStampWords);
  if (cudaMemcpy(out, c->mc.stamps, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaMemcpy failed");
  return REPPO_OK;
}

int reppo_dp_status(reppo_ctx* c, int32_t* timed_out) {
  if (timed_out == nullptr) return fail(c, REPPO_ERR_INVALID, "reppo_dp_status: null argument");
  *timed_out = 0;
  if (c->mc.status == nullptr) return REPPO_OK;
  int v = 0;
  if (cudaMemcpy(&v, c->mc.status, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(c, REPPO_ERR_CUDA, "cudaMemcpy failed");
  *timed_out = v;
  return REPPO_OK;
}

int reppo_nccl_finalize(reppo_ctx* c) {
  // The captured graphs hold references on the communicators (NCCL counts graph-captured collectives as persistent
  // operations and ncclCommDestroy waits for them): the executable graphs go first, the communicators after.
  cudaDeviceSynchronize();
  for (int i = 0; i < reppo_ctx::kGraphSlots; ++i)
    if (c->graph_exec[i]) { cudaGraphExecDestroy(c->graph_exec[i]); c->graph_exec[i] = nullptr; }
  c->invalidate_graphs();
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) if (c->comm[i]) { c->nccl.CommDestroy(c->comm[i]); c->comm[i] = nullptr; }
  c->mc.on = false;
  c->world = 1;
  c->rank = 0;
  return REPPO_OK;
}

}  // extern "C"
