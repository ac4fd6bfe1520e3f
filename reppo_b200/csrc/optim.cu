// K12 -- global-norm clip + Adam / AdamW on one flat fp32 arena.
//   JAX  : optax.chain(clip_by_global_norm(max_grad_norm), adam(lr))  src/jaxrl/reppo.py:246-257
//          lr = optax.linear_schedule(cfg.lr, 0, num_updates) when cfg.anneal_lr (:239-244)
//   Torch: clip_grad_norm_ (+1e-6) -> AdamW(weight_decay=0.01)         src/torchrl/reppo.py:270-273,527-534
// One kernel template, two ways to run it:
//  * SPLIT (default): launch 1 (PHASE 1) reduces every block's slice of the gradient to one partial sum of squares and advances
//    the device step counter; launch 2 (PHASE 2, programmatic dependent launch) re-adds the <= kMaxPartials partials in a fixed
//    order (deterministic global norm, no float atomics), derives the clip scale, the bias corrections and the (optionally
//    annealed) learning rate, and streams its slice: 28 B per parameter (read g, p, m, v; write p, m, v) + the operand shadow
//    + the zeroed gradient.
//  * FUSED (PHASE 0; REPPO_ADAM_FUSED=1, and always with the multicast all-reduce whose cross-GPU barriers live in this kernel):
//    both phases in one launch around a grid-wide barrier (arrival counter + generation word, replayable inside a CUDA graph).
//    The barrier needs every block resident at once.  The grid never exceeds the SM count and two 512-thread blocks fit on an
//    SM, so the critic chain's and the actor chain's launches can always coexist -- but that holds only while nothing else
//    occupies the SMs: with the input pipeline's kernels (torch.randperm / normal_ on the copy stream) running beside the graph
//    the resident blocks were seen spinning for seconds (1 of 20 end-to-end runs; 4 of 18 in a variant), which is why the fused
//    form is no longer the default, and why its barrier gives up after three seconds and says so instead of hanging the device.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

constexpr int kAdamKeep = 4;   // float4 gradient chunks a thread keeps in registers across the barrier

// ---- data-parallel phase: all-reduce of the gradient arena through NVSwitch multicast memory -------------------------------
// Every rank launches this kernel with the same grid.  Block b of rank R exchanges only with block b of the other ranks
// ("twins"): the float4 chunks of block b's own slice (the ones it will clip and apply below) are split over the ranks by
// warp -- warp w of every twin reduces the chunks of warp w iff w % world == R -- so after the second twin barrier block b
// holds the summed gradient of exactly its slice and needs no device-wide wait beyond the norm barrier it has anyway.
//   barrier A (twins)  : every rank's gradient kernels of this step have finished (stream order on each rank)
//   multimem.ld_reduce : the switch returns the fp32 sum over all ranks' copies of 16 bytes; multimem.st writes it to all
//   barrier B (twins)  : the twins' sums have landed in this rank's copy
// One flag word per (block, source rank) in every rank's flag block.  Every launch of a chain advances that chain's EPOCH
// word (local memory, written by the last block to reach the norm barrier, i.e. after every block has read it) by two;
// barrier A publishes epoch + 1, barrier B epoch + 2.  All ranks replay the same launches, so their epochs agree, flags only
// grow, and the kernel is replayable inside a CUDA graph.  A wait that exceeds two seconds raises *status and proceeds
// (no device hang on a lost rank).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// One twin barrier.  `seq` is the value this barrier instance publishes: flag words only ever grow (epoch counter below), so a
// signal is ONE one-way store into the peer's memory and a wait is a poll of local memory -- no atomic round trips over
// NVLink.  publish: the block's multimem stores must be performed before the flag (release by the signalling threads after
// the block barrier; cumulativity carries the other warps' stores).  Returns the time every warp had arrived (stamps only).
__device__ __forceinline__ unsigned long long mc_twin_barrier(const McArgs& mc, unsigned int seq, bool publish) {
  __syncthreads();
  const unsigned long long arrived = (mc.stamps != nullptr && threadIdx.x == 0) ? global_ns() : 0ull;
  if (threadIdx.x < mc.world) {
    const int peer = threadIdx.x;
    unsigned int* put = mc.flags[peer] + (blockIdx.x * mc.world + mc.rank);
    const unsigned int* get = mc.flags[mc.rank] + (blockIdx.x * mc.world + peer);
    if (publish) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(put), "r"(seq) : "memory");
    else asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(put), "r"(seq) : "memory");
    if (*reinterpret_cast<volatile int*>(mc.status) == 0) {   // after a first time-out the rest of the graph drains without waiting
      const unsigned long long t0 = global_ns();
      unsigned int seen;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(get) : "memory");
        if (static_cast<int>(seen - seq) >= 0) break;
        if (global_ns() - t0 > 2000000000ull) { atomicExch(mc.status, 1); break; }
      }
    }
  }
  __syncthreads();
  return arrived;
}
__device__ __forceinline__ float4 multimem_sum(const float4* p) {
  float4 q;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p) : "memory");
  return q;
}
__device__ __forceinline__ void multimem_store(float4* p, const float4& q) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w) : "memory");
}

// PHASE: 0 = one launch with the grid barrier (fused), 1 = partial sums of squares only, 2 = clip + Adam from the partials
template <int THREADS, int PHASE>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
clip_adam_kernel(const float* p_in, float* p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 float* __restrict__ partials, unsigned int* __restrict__ bar, int* __restrict__ step, AdamConsts c,
                 const ShadowTable sh, float* __restrict__ grad_norm_out, int zero_grads, const McArgs mc) {
  REPPO_PDL_PROLOGUE();
  __shared__ float red[32];
  __shared__ float s_scale, s_bc1, s_bc2, s_lr;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* pi4 = reinterpret_cast<const float4*>(p_in);   // may alias p (in-place update)
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* g4 = reinterpret_cast<float4*>(g);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned long long ts[6] = {0, 0, 0, 0, 0, 0};
  if (PHASE == 0 && mc.world > 1) {
    // ---- phase 0 (data parallel): this block's slice of the gradient becomes the sum over all ranks
    const bool stamp = mc.stamps != nullptr && threadIdx.x == 0;
    if (stamp) ts[0] = global_ns();
    unsigned int epoch;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch) : "l"(mc.epoch) : "memory");
    ts[1] = mc_twin_barrier(mc, epoch + 1u, false);
    if (stamp) ts[2] = global_ns();
    float4* gm = reinterpret_cast<float4*>(mc.g_mc);
    if (static_cast<int>(threadIdx.x >> 5) % mc.world == mc.rank) {
      // four switch-side sums in flight per thread before the first store (a load + store pair per iteration would cost
      // one NVLink round trip each: measured 11.5 us for the critic arena on 2 GPUs)
      long long i = first;
      for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 q0 = multimem_sum(gm + i), q1 = multimem_sum(gm + i + stride), q2 = multimem_sum(gm + i + 2 * stride),
                     q3 = multimem_sum(gm + i + 3 * stride);
        multimem_store(gm + i, q0); multimem_store(gm + i + stride, q1);
        multimem_store(gm + i + 2 * stride, q2); multimem_store(gm + i + 3 * stride, q3);
      }
      if (i + stride < n4) {
        const float4 q0 = multimem_sum(gm + i), q1 = multimem_sum(gm + i + stride);
        multimem_store(gm + i, q0); multimem_store(gm + i + stride, q1);
        i += 2 * stride;
      }
      for (; i < n4; i += stride) multimem_store(gm + i, multimem_sum(gm + i));
    }
    // the ragged tail (the arena is padded to a multiple of four floats) belongs to block 0, the only one that reads it
    if (blockIdx.x == 0 && threadIdx.x == 0 && mc.rank == 0 && (n & 3)) multimem_store(gm + n4, multimem_sum(gm + n4));
    ts[3] = mc_twin_barrier(mc, epoch + 2u, true);
    if (stamp) ts[4] = global_ns();
  }
  // ---- phase 1: this block's partial sum of squares
  float4 gk[kAdamKeep];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kAdamKeep; ++k) {
    const long long i = first + k * stride;
    gk[k] = (i < n4) ? __ldcg(g4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < kAdamKeep; ++k) s += gk[k].x * gk[k].x + gk[k].y * gk[k].y + gk[k].z * gk[k].z + gk[k].w * gk[k].w;
  for (long long i = first + kAdamKeep * stride; i < n4; i += stride) {   // wide networks only
    const float4 q = __ldcg(g4 + i);
    s += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float q = __ldcg(g + (n4 << 2) + threadIdx.x); s += q * q; }
  if constexpr (PHASE != 2) s = block_sum(s, red);
  if constexpr (PHASE == 1) {
    if (threadIdx.x == 0) {
      partials[blockIdx.x] = s;
      if (blockIdx.x == 0) *step = *step + 1;   // the update launch only reads the counter: no block can see it half-way
    }
    return;
  }
  // fused: every block reads the counter before the barrier, block 0 writes it after; split: already advanced by launch 1
  const int t = (PHASE == 2) ? *step : *step + 1;
  unsigned int my_gen = 0;
  bool last = false;
  if (PHASE == 0 && threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(my_gen) : "l"(bar + 1) : "memory");   // before arriving
    __threadfence();
    last = (atomicAdd(bar, 1u) == gridDim.x - 1);
  }
  // phase-2 operands of the first pass are requested before the wait
  float4 pp = make_float4(0.f, 0.f, 0.f, 0.f), mm = pp, vv = pp;
  if (first < n4) { pp = pi4[first]; mm = m4[first]; vv = v4[first]; }
  if (PHASE == 0 && threadIdx.x == 0) {
    if (last) {
      if (mc.world > 1) *mc.epoch += 2u;   // every block read the epoch before it arrived here
      bar[0] = 0u;   // nobody touches the counter again before the generation moves
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(my_gen + 1u) : "memory");
    } else {
      unsigned int seen;
      unsigned long long t0 = 0ull;
      unsigned int spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory");
        if (seen == my_gen && (++spins & 0xfffu) == 0u) {
          // a grid barrier that cannot complete (not every block resident) must not hang the device: say so and go on
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0ull) t0 = now;
          else if (now - t0 > 3000000000ull) {
            unsigned int arrived;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(arrived) : "l"(bar) : "memory");
            printf("reppo clip_adam_kernel: grid barrier timed out (block %d of %d, %u arrived, generation %u, n %lld)\n",
                   static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), arrived, my_gen, n);
            break;
          }
        }
      } while (seen == my_gen);
    }
  }
  __syncthreads();
  // ---- phase 2
  float tot = 0.f;
  for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x) tot += __ldcg(partials + i);
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) {
    const float gn = sqrtf(tot);
    float scale = 1.f;
    if (c.max_grad_norm > 0.f) {
      if (c.torch_clip) scale = fminf(c.max_grad_norm / (gn + 1e-6f), 1.f);   // clip_grad_norm_
      else scale = (gn < c.max_grad_norm) ? 1.f : c.max_grad_norm / gn;       // optax.clip_by_global_norm
    }
    s_scale = scale;
    s_bc1 = static_cast<float>(1.0 - pow(static_cast<double>(c.b1), static_cast<double>(t)));
    s_bc2 = static_cast<float>(1.0 - pow(static_cast<double>(c.b2), static_cast<double>(t)));
    float lr = c.lr;
    if (c.anneal_steps > 0) {
      // optax.linear_schedule(lr, 0, T) at count = t - 1 (polynomial_schedule, power 1): lr * (1 - min(count, T) / T)
      const float cnt = static_cast<float>(min(t - 1, c.anneal_steps));
      lr = c.lr * (1.f - cnt / static_cast<float>(c.anneal_steps));
    }
    s_lr = lr;
    if (blockIdx.x == 0) {
      if (PHASE == 0) *step = t;
      if (grad_norm_out != nullptr) *grad_norm_out = gn;
    }
  }
  __syncthreads();
  const float scale = s_scale, bc1 = s_bc1, bc2 = s_bc2, lr = s_lr;
  const float decay = 1.f - lr * c.weight_decay;
  auto upd = [&](float& p1, float g1, float& m1, float& v1) {
    g1 *= scale;
    m1 = c.b1 * m1 + (1.f - c.b1) * g1;
    v1 = c.b2 * v1 + (1.f - c.b2) * g1 * g1;
    p1 = p1 * decay - lr * (m1 / bc1) / (sqrtf(v1 / bc2) + c.eps);
  };
  // re-round updated master weights into the bf16 operand shadow W[out, ld] of the tcgen05 GEMMs (one element)
  auto shadow1 = [&](float p1, long long idx) {
    for (int e = 0; e < sh.n; ++e) {
      if (idx >= sh.e[e].start && idx < sh.e[e].end) {
        const int off = static_cast<int>(idx - sh.e[e].start);
        const int r = off / sh.e[e].in, col = off - r * sh.e[e].in;
        if (sh.fp32) static_cast<float*>(sh.e[e].dst)[static_cast<size_t>(r) * sh.e[e].ld + col] = p1;
        else static_cast<__nv_bfloat16*>(sh.e[e].dst)[static_cast<size_t>(r) * sh.e[e].ld + col] = __float2bfloat16_rn(p1);
        break;
      }
    }
  };
  // one table search per float4: the four elements almost always sit in one row of one weight matrix
  auto shadow4 = [&](const float4& q, long long idx) {
    bool done = false;
    for (int e = 0; e < sh.n; ++e) {
      if (idx >= sh.e[e].start && idx + 3 < sh.e[e].end) {
        const int off = static_cast<int>(idx - sh.e[e].start);
        const int r = off / sh.e[e].in, col = off - r * sh.e[e].in;
        if (col + 3 < sh.e[e].in && !sh.fp32) {
          __nv_bfloat16* d = static_cast<__nv_bfloat16*>(sh.e[e].dst) + static_cast<size_t>(r) * sh.e[e].ld + col;
          const __nv_bfloat162 lo = __floats2bfloat162_rn(q.x, q.y), hi = __floats2bfloat162_rn(q.z, q.w);
          if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) {
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&lo);
            u.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(d) = u;
          } else {
            d[0] = __low2bfloat16(lo); d[1] = __high2bfloat16(lo); d[2] = __low2bfloat16(hi); d[3] = __high2bfloat16(hi);
          }
          done = true;
        }
        break;
      }
    }
    if (!done) { shadow1(q.x, idx); shadow1(q.y, idx + 1); shadow1(q.z, idx + 2); shadow1(q.w, idx + 3); }
  };
  int k = 0;
  for (long long i = first; i < n4; i += stride, ++k) {
    float4 gg;
    if (k < kAdamKeep) {
      // static indexing keeps gk in registers
      gg = (k == 0) ? gk[0] : (k == 1) ? gk[1] : (k == 2) ? gk[2] : gk[3];
    } else {
      gg = __ldcg(g4 + i);
    }
    if (k > 0) { pp = pi4[i]; mm = m4[i]; vv = v4[i]; }
    if (zero_grads) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // the next step accumulates into a clean arena (no memset node)
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
    if (sh.n > 0) shadow4(pp, 4 * i);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float q = p_in[i];
    upd(q, __ldcg(g + i), m[i], v[i]);
    shadow1(q, i);
    p[i] = q;
    if (zero_grads) g[i] = 0.f;
  }
  if (PHASE == 0 && mc.world > 1 && mc.stamps != nullptr && threadIdx.x == 0) {
    // entry, block arrived at A, A done, reduce done + fenced (arrived at B), B done, exit
    ts[5] = global_ns();
    const unsigned long long slot = atomicAdd(mc.stamps, 1ull) & 8191ull;
    unsigned long long* o = mc.stamps + 8 + slot * 8;
    o[0] = (static_cast<unsigned long long>(mc.chain) << 32) | blockIdx.x;
    for (int k = 0; k < 6; ++k) o[1 + k] = ts[k];
  }
}

bool clip_adam_fused(const McArgs* mc) { return (mc != nullptr && mc->world > 1) || knobs().adam_fused; }

cudaError_t launch_clip_adam(const float* p_in, float* p, float* g, float* m, float* v, long long n, int* step, float* partials,
                             unsigned int* bar, const AdamConsts& c, const ShadowTable* shadow, float* grad_norm_out,
                             int zero_grads, cudaStream_t st, const McArgs* mc) {
  if (n <= 0) return cudaSuccess;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
  }
  // all blocks must be co-resident (grid barrier): at most one per SM
  // 256-thread blocks (16 K registers) find room on an SM that already runs two GEMM CTAs of the other chain, but measured
  // slower (101.7 vs 96.2 ms per C2 update: twice the elements per thread after the barrier); 512 stays the default
  const int threads = knobs().adam_threads == 256 ? 256 : 512;
  const long long want = (n / 4 + threads - 1) / threads;
  int cap = std::min(sm_count, kMaxPartials);
  if (knobs().adam_blocks > 0 && knobs().adam_blocks < cap) cap = knobs().adam_blocks;
  int blocks = static_cast<int>(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  ShadowTable sh{};
  if (shadow != nullptr) sh = *shadow;
  McArgs ma{};
  if (mc != nullptr && mc->world > 1) {
    // warps of a block are dealt to ranks round-robin (16 warps), flag rows are per block
    if ((threads / 32) % mc->world != 0 || mc->world > kMcMaxRanks || blocks > kMcFlagBlocks) return cudaErrorInvalidValue;
    ma = *mc;
  }
  // Two launches by default: the fused kernel's grid barrier needs every block resident at once, and other work on the device
  // (the input pipeline's kernels on the copy stream were enough) can keep the last blocks out while the resident ones spin --
  // observed as a multi-second stall in 1 of 20 end-to-end runs.  The fused launch remains for the multicast all-reduce (its
  // cross-GPU barriers live in the same kernel) and behind REPPO_ADAM_FUSED=1.
  const bool fused = clip_adam_fused(mc);
#define REPPO_ADAM_LAUNCH(T, PH)                                                                                      \
  launch_k((clip_adam_kernel<T, PH>), dim3(blocks), dim3(T), 0, st, p_in, p, g, m, v, n, partials, bar, step, c, sh, \
           grad_norm_out, zero_grads, ma)
  if (threads == 512) {
    if (fused) { REPPO_ADAM_LAUNCH(512, 0); } else { REPPO_ADAM_LAUNCH(512, 1); REPPO_ADAM_LAUNCH(512, 2); }
  } else {
    if (fused) { REPPO_ADAM_LAUNCH(256, 0); } else { REPPO_ADAM_LAUNCH(256, 1); REPPO_ADAM_LAUNCH(256, 2); }
  }
#undef REPPO_ADAM_LAUNCH
  return cudaGetLastError();
}

// mean of the step metrics of the last epoch (src/jaxrl/reppo.py:660,671)
__global__ void metrics_mean_kernel(const float* __restrict__ step_metrics, int first, int count, int nm, float* __restrict__ out) {
  REPPO_PDL_PROLOGUE();
  const int k = threadIdx.x;
  if (k >= nm) return;
  float s = 0.f;
  for (int i = 0; i < count; ++i) s += step_metrics[static_cast<size_t>(first + i) * nm + k];
  out[k] = s / static_cast<float>(count);
}

cudaError_t launch_metrics_mean(const float* step_metrics, int first, int count, int nm, float* out, cudaStream_t st) {
  launch_k((metrics_mean_kernel), dim3(1), dim3(32), 0, st, step_metrics, first, count, nm, out);
  return cudaGetLastError();
}

// metric vector prologue: zero the sums, seed the min / max slots
__global__ void metrics_init_kernel(float* __restrict__ m, uint32_t mask) {
  REPPO_PDL_PROLOGUE();
  const int k = threadIdx.x;
  if (k < M_NUM && ((mask >> k) & 1u)) m[k] = (k == M_TARGET_MAX) ? -INFINITY : (k == M_TARGET_MIN ? INFINITY : 0.f);
}
// every step's metric vector of a whole update at once
__global__ void metrics_init_all_kernel(float* __restrict__ m, int steps) {
  REPPO_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= steps * M_NUM) return;
  const int k = i % M_NUM;
  m[i] = (k == M_TARGET_MAX) ? -INFINITY : (k == M_TARGET_MIN ? INFINITY : 0.f);
}
cudaError_t launch_metrics_init_all(float* m, int steps, cudaStream_t st) {
  launch_k((metrics_init_all_kernel), dim3((steps * M_NUM + 255) / 256), dim3(256), 0, st, m, steps);
  return cudaGetLastError();
}
cudaError_t launch_metrics_init(float* m, uint32_t mask, cudaStream_t st) {
  launch_k((metrics_init_kernel), dim3(1), dim3(32), 0, st, m, mask);
  return cudaGetLastError();
}

}  // namespace reppo
