// K12 -- global-norm clip + Adam / AdamW on one flat fp32 arena.
//   JAX  : optax.chain(clip_by_global_norm(max_grad_norm), adam(lr))  src/jaxrl/reppo.py:246-257
//   Torch: clip_grad_norm_ (+1e-6) -> AdamW(weight_decay=0.01)         src/torchrl/reppo.py:270-273,527-534
// Two launches: (1) per-block partial sums of squares (deterministic, no atomics) + step counter
// increment; (2) every block re-reduces the <= kMaxPartials partials, derives the clip scale and
// the bias corrections, and streams its slice: 28 B per parameter (read g,p,m,v; write p,m,v).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace reppo {

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials, int* __restrict__ step) {
  REPPO_PDL_PROLOGUE();
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s += v * v; }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    if (blockIdx.x == 0 && step != nullptr) *step = *step + 1;
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(const float* p_in, float* p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            const float* __restrict__ partials, int num_partials, const int* __restrict__ step, AdamConsts c,
            const ShadowTable sh, float* __restrict__ grad_norm_out, int zero_grads) {
  REPPO_PDL_PROLOGUE();
  __shared__ float red[32];
  __shared__ float s_scale, s_bc1, s_bc2;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* pi4 = reinterpret_cast<const float4*>(p_in);   // may alias p (in-place update)
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  float4* g4 = reinterpret_cast<float4*>(g);
  // Two float4 per thread per pass, and the first pass's loads are issued BEFORE the global-norm prologue below (they do
  // not depend on it): the kernel is load-latency bound (ncu: 54 % of the stall samples sit on the first use of the loaded
  // values, 28 warps per SM), so the prologue's reduction / pow and the second float4 hide behind the first one's latency.
  constexpr int U = 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long base = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float4 pp[U], mm[U], vv[U], gg[U];
  auto load = [&](long long b0) {
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long i = b0 + k * stride;
      if (i < n4) { pp[k] = pi4[i]; mm[k] = m4[i]; vv[k] = v4[i]; gg[k] = g4[i]; }
    }
  };
  load(base);
  float s = 0.f;
  for (int i = threadIdx.x; i < num_partials; i += blockDim.x) s += partials[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float gn = sqrtf(s);
    float scale = 1.f;
    if (c.max_grad_norm > 0.f) {
      if (c.torch_clip) scale = fminf(c.max_grad_norm / (gn + 1e-6f), 1.f);   // clip_grad_norm_
      else scale = (gn < c.max_grad_norm) ? 1.f : c.max_grad_norm / gn;       // optax.clip_by_global_norm
    }
    const int t = *step;
    s_scale = scale;
    s_bc1 = static_cast<float>(1.0 - pow(static_cast<double>(c.b1), static_cast<double>(t)));
    s_bc2 = static_cast<float>(1.0 - pow(static_cast<double>(c.b2), static_cast<double>(t)));
    if (blockIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = gn;
  }
  __syncthreads();
  const float scale = s_scale, bc1 = s_bc1, bc2 = s_bc2;
  const float decay = 1.f - c.lr * c.weight_decay;
  auto upd = [&](float& p1, float g1, float& m1, float& v1) {
    g1 *= scale;
    m1 = c.b1 * m1 + (1.f - c.b1) * g1;
    v1 = c.b2 * v1 + (1.f - c.b2) * g1 * g1;
    p1 = p1 * decay - c.lr * (m1 / bc1) / (sqrtf(v1 / bc2) + c.eps);
  };
  // re-round updated master weights into the bf16 operand shadow W[out, ld] of the tcgen05 GEMMs (one element)
  auto shadow1 = [&](float p1, long long idx) {
    for (int t = 0; t < sh.n; ++t) {
      if (idx >= sh.e[t].start && idx < sh.e[t].end) {
        const int off = static_cast<int>(idx - sh.e[t].start);
        const int r = off / sh.e[t].in, col = off - r * sh.e[t].in;
        static_cast<__nv_bfloat16*>(sh.e[t].dst)[static_cast<size_t>(r) * sh.e[t].ld + col] = __float2bfloat16_rn(p1);
        if (sh.e[t].dst_t != nullptr)   // thin first layers also keep W^T (thin.cu)
          static_cast<__nv_bfloat16*>(sh.e[t].dst_t)[static_cast<size_t>(col) * sh.e[t].ld_t + r] = __float2bfloat16_rn(p1);
        break;
      }
    }
  };
  // one table search per float4: the four elements almost always sit in one row of one weight matrix
  auto shadow4 = [&](const float4& q, long long idx) {
    bool done = false;
    for (int t = 0; t < sh.n; ++t) {
      if (idx >= sh.e[t].start && idx + 3 < sh.e[t].end) {
        const int off = static_cast<int>(idx - sh.e[t].start);
        const int r = off / sh.e[t].in, col = off - r * sh.e[t].in;
        if (col + 3 < sh.e[t].in && sh.e[t].dst_t == nullptr) {
          __nv_bfloat16* d = static_cast<__nv_bfloat16*>(sh.e[t].dst) + static_cast<size_t>(r) * sh.e[t].ld + col;
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
  for (; base < n4; base += U * stride) {
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long i = base + k * stride;
      if (i < n4) {
        if (zero_grads) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // the next step accumulates into a clean arena (no memset node)
        upd(pp[k].x, gg[k].x, mm[k].x, vv[k].x); upd(pp[k].y, gg[k].y, mm[k].y, vv[k].y);
        upd(pp[k].z, gg[k].z, mm[k].z, vv[k].z); upd(pp[k].w, gg[k].w, mm[k].w, vv[k].w);
        p4[i] = pp[k]; m4[i] = mm[k]; v4[i] = vv[k];
        if (sh.n > 0) shadow4(pp[k], 4 * i);
      }
    }
    if (base + U * stride < n4) load(base + U * stride);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float pp = p_in[i];
    upd(pp, g[i], m[i], v[i]);
    shadow1(pp, i);
    p[i] = pp;
    if (zero_grads) g[i] = 0.f;
  }
}

cudaError_t launch_clip_adam(const float* p_in, float* p, float* g, float* m, float* v, long long n, int* step, float* partials,
                             const AdamConsts& c, const ShadowTable* shadow, float* grad_norm_out, int zero_grads, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long want = (n / 4 + 255) / 256;
  static const int cap_env = getenv("REPPO_ADAM_BLOCKS") ? atoi(getenv("REPPO_ADAM_BLOCKS")) : 0;   // experiment knob
  const int cap = (cap_env > 0 && cap_env < kMaxPartials) ? cap_env : kMaxPartials;
  int blocks = static_cast<int>(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  launch_k((sumsq_partial_kernel), dim3(blocks), dim3(256), 0, st, g, n, partials, step);
  ShadowTable sh{};
  if (shadow != nullptr) sh = *shadow;
  launch_k((adam_kernel), dim3(blocks), dim3(256), 0, st, p_in, p, g, m, v, n, partials, blocks, step, c, sh, grad_norm_out, zero_grads);
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
