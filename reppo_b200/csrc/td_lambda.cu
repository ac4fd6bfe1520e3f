// K1 -- soft TD(lambda) return: reverse scan over [T, N] with per-env state in registers.
//
// Restates compute_nstep_lambda (src/jaxrl/reppo.py:422-450) and compute_gve
// (src/torchrl/reppo.py:175-189).  One thread owns VEC adjacent env columns; every row access
// of a warp is one contiguous 128 B (VEC=1) or 512 B (VEC=4) segment, the T loop is unrolled so
// that several rows of loads are in flight before the dependent arithmetic, and loads/stores
// bypass L1 (streamed once).  HBM-bound: 24 B per (t, n) element with importance weights, 20 B
// without.
//
// Arithmetic is written with explicit round-to-nearest mul/add intrinsics (no FMA contraction)
// so that the recurrence reproduces the reference's fp32 op-by-op evaluation bit for bit
// whenever exp(iw) is exact (iw == 0, the default exploration setting).
#include "common.cuh"
#include "kernels.h"

namespace reppo {

template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
REPPO_DEVINL void ld_stream(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
  }
}
template <int VEC>
REPPO_DEVINL void st_stream(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  } else {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
  }
}

// one recurrence step for one env column (reference op order, no contraction)
REPPO_DEVINL float td_step(float G, float w, float omw, float r, float v, float d, bool trunc, float gamma) {
  // lambda_sum = w * G + (1 - w) * v
  const float ls = __fadd_rn(__fmul_rn(w, G), __fmul_rn(omw, v));
  // delta = gamma * where(trunc, v, (1 - done) * lambda_sum)
  const float inner = trunc ? v : __fmul_rn(__fsub_rn(1.f, d), ls);
  return __fadd_rn(r, __fmul_rn(gamma, inner));
}

template <int VEC, int UNROLL, bool JAX, bool HAS_IW>
__global__ void __launch_bounds__(128)
td_lambda_kernel(const float* __restrict__ soft_reward, const float* __restrict__ value,
                 const float* __restrict__ done, float* __restrict__ truncated,
                 const float* __restrict__ iw, float* __restrict__ target,
                 int T, long long N, float gamma, float lmbda, float one_minus_lmbda, int mutate) {
  const long long col = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
  if (col >= N) return;

  // JAX: (1 - w) is an fp32 subtraction of the fp32 weight; Torch: the Python double (1.0 - lmbda)
  // rounded to fp32 (scalar * tensor), passed in by the host.
  float G[VEC], w[VEC], omw[VEC];
  bool tr[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) omw[j] = JAX ? __fsub_rn(1.f, lmbda) : one_minus_lmbda;
  if (JAX) {
    // carry seeded with (value[T-1], trunc = 1, iw = 0)   (src/jaxrl/reppo.py:444-446)
    float v_last[VEC];
    ld_stream<VEC>(value + static_cast<long long>(T - 1) * N + col, v_last);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { G[j] = v_last[j]; tr[j] = true; w[j] = lmbda; }
  } else {
    // last_gve = 0 (src/torchrl/reppo.py:177)
#pragma unroll
    for (int j = 0; j < VEC; ++j) { G[j] = 0.f; tr[j] = false; w[j] = lmbda; }
    if (mutate) {
      float one[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) one[j] = 1.f;
      st_stream<VEC>(truncated + static_cast<long long>(T - 1) * N + col, one);  // :178
    }
  }

  int t = T - 1;
  while (t >= 0) {
    const int nb = (t + 1 < UNROLL) ? (t + 1) : UNROLL;
    float r[UNROLL][VEC], v[UNROLL][VEC], d[UNROLL][VEC], q[UNROLL][VEC], x[UNROLL][VEC];
    // issue all loads of the next UNROLL rows first
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (u < nb) {
        const long long off = static_cast<long long>(t - u) * N + col;
        ld_stream<VEC>(soft_reward + off, r[u]);
        ld_stream<VEC>(value + off, v[u]);
        ld_stream<VEC>(done + off, d[u]);
        // the row T-1 of `truncated` may be written above by this very thread: plain load
        if constexpr (VEC == 4) {
          const float4 tq = *reinterpret_cast<const float4*>(truncated + off);
          q[u][0] = tq.x; q[u][1] = tq.y; q[u][2] = tq.z; q[u][3] = tq.w;
        } else {
          q[u][0] = truncated[off];
        }
        if (HAS_IW) ld_stream<VEC>(iw + off, x[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (u < nb) {
        float out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          bool trunc_now;
          if (JAX) trunc_now = tr[j];                              // flag of step t+1
          else trunc_now = (q[u][j] != 0.f) || (t - u == T - 1);   // truncated[t], last row forced
          G[j] = td_step(G[j], w[j], omw[j], r[u][j], v[u][j], d[u][j], trunc_now, gamma);
          out[j] = G[j];
          if (JAX) {
            tr[j] = (q[u][j] != 0.f);
            if (HAS_IW) {                                          // weight used at step t-1
              w[j] = __fmul_rn(expf(x[u][j]), lmbda);
              omw[j] = __fsub_rn(1.f, w[j]);
            }
          }
        }
        st_stream<VEC>(target + static_cast<long long>(t - u) * N + col, out);
      }
    }
    t -= nb;
  }
}

template <int VEC, int UNROLL>
static void launch(const float* sr, const float* val, const float* done, float* trunc, const float* iw, float* out,
                   int T, long long N, float gamma, float lmbda, float oml, int convention, int mutate, cudaStream_t st) {
  const long long threads = (N + VEC - 1) / VEC;
  const int block = 128;
  const unsigned grid = static_cast<unsigned>((threads + block - 1) / block);
  if (convention == 0) {
    if (iw) td_lambda_kernel<VEC, UNROLL, true, true><<<grid, block, 0, st>>>(sr, val, done, trunc, iw, out, T, N, gamma, lmbda, oml, 0);
    else td_lambda_kernel<VEC, UNROLL, true, false><<<grid, block, 0, st>>>(sr, val, done, trunc, nullptr, out, T, N, gamma, lmbda, oml, 0);
  } else {
    td_lambda_kernel<VEC, UNROLL, false, false><<<grid, block, 0, st>>>(sr, val, done, trunc, nullptr, out, T, N, gamma, lmbda, oml, mutate);
  }
}

cudaError_t launch_td_lambda(const float* sr, const float* val, const float* done, float* trunc, const float* iw,
                             float* out, int T, long long N, double gamma_d, double lmbda_d, int convention, int mutate,
                             cudaStream_t st) {
  const float gamma = static_cast<float>(gamma_d), lmbda = static_cast<float>(lmbda_d);
  const float oml = static_cast<float>(1.0 - lmbda_d);
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec_ok = (N % 4 == 0) && aligned16(sr) && aligned16(val) && aligned16(done) && aligned16(trunc) &&
                      aligned16(out) && (iw == nullptr || aligned16(iw));
  // small N: one column per thread keeps more SMs busy; large N: 16 B per thread per access
  if (vec_ok && N >= 4LL * 128 * 148 * 4) launch<4, 4>(sr, val, done, trunc, iw, out, T, N, gamma, lmbda, oml, convention, mutate, st);
  else launch<1, 8>(sr, val, done, trunc, iw, out, T, N, gamma, lmbda, oml, convention, mutate, st);
  return cudaGetLastError();
}

}  // namespace reppo
