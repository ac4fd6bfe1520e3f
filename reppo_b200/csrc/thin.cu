// Thin layers on the SIMT pipes (REPPO_GEMM_BF16 mode).
//
// The first Linear of every MLP has K = obs (+ action) dims (17 / 23 at CheetahRun), the policy's last Linear has
// N = 2 A outputs (12).  As tcgen05 GEMMs they are a full kernel of the step's dependent chain (TMA fill, TMEM, 8-warp
// epilogue: 7-10 us) for a few MFLOP.  Here they ride inside the row kernels next to them, one warp per row:
//
//  thin_first_layer_kernel : minibatch gather + Linear(K <= 64) + bias + RMS/LayerNorm + swish   (replaces gather + GEMM)
//  actor_head_sample_kernel: Linear(H -> 2A) + tanh-Gaussian sample + log-prob + forward KL      (replaces GEMM + sample)
//  actor_grad_thin_kernel  : dQ/da through the critic's first Linear, the policy-gradient row math, and the policy's
//                            last-layer dgrad                                                   (replaces 2 GEMMs + actor_grad)
//
// Arithmetic matches the tensor-core path: operands rounded to bf16 (the same weight shadows / activation twins),
// products and sums in fp32.  Reference: FCNN src/networks/jax_models.py:57-124, torch_models.py:82-142;
// actor loss src/jaxrl/reppo.py:526-622.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "loss_math.cuh"

namespace reppo {

REPPO_DEVINL float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
REPPO_DEVINL float2 bf16x2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---- first layer -------------------------------------------------------------------------------------
// x[r, :] = [ a[ia(r), 0:da] | b[ib(r), 0:db] ],  ia(r) = idx ? idx[r] : r,  ib(r) = b_local ? r : ia(r)
// z = x W^T + bias (stored fp32), y = swish(norm(z)) -> bf16 twin, rstd / mean per row, x -> bf16 twin (weight gradient)
// Lane owns columns 64 j + 2 lane + {0, 1}; W^T (bf16 [K][H]) is staged in shared memory.
template <int NJ, int NORM>
__global__ void __launch_bounds__(256)
thin_first_layer_kernel(const float* __restrict__ a, int da, int lda, const float* __restrict__ b, int db, int ldb,
                        const int* __restrict__ idx, int b_local, int rows, const __nv_bfloat16* __restrict__ wT, int ldw,
                        const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        float* __restrict__ z, float* __restrict__ rstd_out, float* __restrict__ mean_out,
                        __nv_bfloat16* __restrict__ y_h, int ldy, __nv_bfloat16* __restrict__ x_h, int ldx) {
  REPPO_PDL_PROLOGUE();
  constexpr int H = 64 * NJ;
  extern __shared__ uint32_t wt[];   // [K][H / 2] bf16x2
  const int K = da + db;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // wT: the layer's TRANSPOSED bf16 shadow [K][ldw >= H] (refreshed by the Adam kernel), copied with 16-byte loads
  for (int i = threadIdx.x; i < K * (H / 8); i += blockDim.x) {
    const int k = i / (H / 8), c8 = i - k * (H / 8);
    reinterpret_cast<uint4*>(wt)[i] = reinterpret_cast<const uint4*>(wT + static_cast<size_t>(k) * ldw)[c8];
  }
  const int row = blockIdx.x * 8 + warp;
  // this lane's input elements k = lane, lane + 32 (K <= 64), rounded to bf16 like the GEMM operand
  float xv[2] = {0.f, 0.f};
  if (row < rows) {
    const long long ia = idx ? idx[row] : row;
    const long long ib = b_local ? row : ia;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = lane + 32 * h;
      if (k < K) {
        const float v = k < da ? a[ia * lda + k] : b[ib * ldb + (k - da)];
        xv[h] = bf16_round(v);
        if (x_h != nullptr) x_h[static_cast<size_t>(row) * ldx + k] = __float2bfloat16_rn(v);
      }
    }
  }
  __syncthreads();
  if (row >= rows) return;
  float acc[NJ][2];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    acc[j][0] = bias[64 * j + 2 * lane];       // flat parameter arena: only 4-byte aligned
    acc[j][1] = bias[64 * j + 2 * lane + 1];
  }
  for (int k = 0; k < K; ++k) {
    const float xk = __shfl_sync(0xffffffffu, k < 32 ? xv[0] : xv[1], k & 31);
    const uint32_t* wr = wt + k * (H / 2) + lane;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float2 ww = bf16x2_to_f2(wr[32 * j]);
      acc[j][0] = fmaf(xk, ww.x, acc[j][0]);
      acc[j][1] = fmaf(xk, ww.y, acc[j][1]);
    }
  }
  float rstd = 1.f, mean = 0.f;
  if (NORM != NORM_NONE) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { s1 += acc[j][0] + acc[j][1]; s2 += acc[j][0] * acc[j][0] + acc[j][1] * acc[j][1]; }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (NORM == NORM_RMS) {
      rstd = rsqrtf(s2 / H + eps);
    } else {
      mean = s1 / H;
      rstd = rsqrtf(s2 / H - mean * mean + eps);   // flax use_fast_variance
    }
    if (lane == 0) { rstd_out[row] = rstd; if (NORM == NORM_LAYER) mean_out[row] = mean; }
  }
  float* zr = z + static_cast<size_t>(row) * H;
  __nv_bfloat16* yr = y_h + static_cast<size_t>(row) * ldy;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 64 * j + 2 * lane;
    *reinterpret_cast<float2*>(zr + c) = make_float2(acc[j][0], acc[j][1]);
    float y0 = acc[j][0], y1 = acc[j][1];
    if (NORM != NORM_NONE) {
      y0 = (y0 - mean) * rstd * gamma[c] + (NORM == NORM_LAYER ? beta[c] : 0.f);
      y1 = (y1 - mean) * rstd * gamma[c + 1] + (NORM == NORM_LAYER ? beta[c + 1] : 0.f);
    }
    const __nv_bfloat162 p = __floats2bfloat162_rn(swishf_(y0), swishf_(y1));
    *reinterpret_cast<__nv_bfloat162*>(yr + c) = p;
  }
}

template <int NJ>
static cudaError_t launch_thin_first_nj(int norm, dim3 grid, size_t smem, cudaStream_t st, const float* a, int da, int lda,
                                        const float* b, int db, int ldb, const int* idx, int b_local, int rows, const void* w, int ldw,
                                        const float* bias, const float* gamma, const float* beta, float eps, float* z, float* rstd,
                                        float* mean, void* y_h, int ldy, void* x_h, int ldx) {
  const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y_h);
  __nv_bfloat16* xp = static_cast<__nv_bfloat16*>(x_h);
#define REPPO_THIN_LAUNCH(NRM)                                                                                              \
  do {                                                                                                                       \
    auto kern = thin_first_layer_kernel<NJ, NRM>;                                                                            \
    if (smem > 48 * 1024) {                                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));       \
      if (e != cudaSuccess) return e;                                                                                        \
    }                                                                                                                        \
    launch_k((kern), grid, dim3(256), smem, st, a, da, lda, b, db, ldb, idx, b_local, rows, wp, ldw, bias, gamma, beta, eps, \
             z, rstd, mean, yp, ldy, xp, ldx);                                                                               \
  } while (0)
  if (norm == NORM_NONE) REPPO_THIN_LAUNCH(NORM_NONE);
  else if (norm == NORM_RMS) REPPO_THIN_LAUNCH(NORM_RMS);
  else REPPO_THIN_LAUNCH(NORM_LAYER);
#undef REPPO_THIN_LAUNCH
  return cudaGetLastError();
}

// REPPO_THIN = bit mask of the SIMT thin-layer kernels to use: 1 = first layer, 2 = policy head + sampling, 4 = policy gradient.
// Default 0: each of them measured slower in the update than the tcgen05 GEMM + row kernel it replaces.
// Read per call: launches are captured into the update graph once.
bool thin_enabled(int bit) {
  const char* e = getenv("REPPO_THIN");
  return e != nullptr && (atoi(e) & bit) != 0;
}

// cudaErrorNotSupported: shape outside the thin path (caller runs gather + tcgen05 GEMM)
cudaError_t launch_thin_first_layer(int norm, const float* a, int da, int lda, const float* b, int db, int ldb, const int* idx,
                                    int b_local, int rows, int H, const void* w_bf16, int ldw, const float* bias, const float* gamma,
                                    const float* beta, float eps, float* z, float* rstd, float* mean, void* y_h, int ldy, void* x_h,
                                    int ldx, cudaStream_t st) {
  const int K = da + db;
  if (!thin_enabled(1) || rows <= 0 || K <= 0 || K > 64 || (ldy & 1) || (ldw & 7) || bias == nullptr || z == nullptr ||
      y_h == nullptr || w_bf16 == nullptr)
    return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(z) & 7) || (reinterpret_cast<uintptr_t>(w_bf16) & 15)) return cudaErrorNotSupported;
  const size_t smem = static_cast<size_t>(K) * (H / 2) * sizeof(uint32_t);
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  const dim3 grid((rows + 7) / 8);
#define REPPO_THIN_NJ(N) \
  return launch_thin_first_nj<N>(norm, grid, smem, st, a, da, lda, b, db, ldb, idx, b_local, rows, w_bf16, ldw, bias, gamma, beta, eps, z, rstd, mean, y_h, ldy, x_h, ldx)
  switch (H) {
    case 128: REPPO_THIN_NJ(2);
    case 256: REPPO_THIN_NJ(4);
    case 512: REPPO_THIN_NJ(8);
    case 1024: REPPO_THIN_NJ(16);
    default: return cudaErrorNotSupported;
  }
#undef REPPO_THIN_NJ
}

// ---- policy head + sampling --------------------------------------------------------------------------
// out[r, 0:2A] = x[r, 0:H] W3^T + b3 (x: bf16 twin of the last hidden activation), then exactly actor_sample_kernel
// (losses.cu): half-warp per row, lane sl owns KL sample sl (+16 ...), action dimension sl (+16).
__global__ void __launch_bounds__(256)
actor_head_sample_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int H, const __nv_bfloat16* __restrict__ w, int ldw,
                         const float* __restrict__ bias, float* __restrict__ out, const float* __restrict__ out_old,
                         const int* __restrict__ old_idx, const float* __restrict__ eps_pi, const float* __restrict__ eps_kl, int rows,
                         int A, int kl_samples, ActorConsts ac, float* __restrict__ act_out, int act_ld,
                         float* __restrict__ logp_out, float* __restrict__ kl_out, float* __restrict__ gmu_kl,
                         float* __restrict__ gsig_kl, __nv_bfloat16* __restrict__ act_h, int act_ldh) {
  REPPO_PDL_PROLOGUE();
  extern __shared__ uint32_t sm_u32[];
  const int O = 2 * A;
  uint32_t* w3 = sm_u32;                                          // [O][H / 2] bf16x2 (row o = output o)
  float* so = reinterpret_cast<float*>(sm_u32 + O * (H / 2));     // [16 half-warps][O] head outputs
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hw = lane >> 4, sl = lane & 15;
  for (int i = threadIdx.x; i < O * (H / 2); i += blockDim.x) {
    const int o = i / (H / 2), c2 = i - o * (H / 2);
    w3[i] = reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(o) * ldw)[c2];
  }
  const int rl = (blockIdx.x * 8 + warp) * 2 + hw;
  const bool ok = rl < rows;
  const int row = ok ? rl : rows - 1;
  __syncthreads();
  // ---- head: every lane of the half-warp covers H / 16 inputs (pairs 2 (sl + 16 i), +1)
  float* myo = so + (warp * 2 + hw) * O;
  {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + static_cast<size_t>(row) * ldx);
    for (int o = 0; o < O; ++o) {
      float s = 0.f;
      for (int c2 = sl; c2 < H / 2; c2 += 16) {
        const float2 xx = bf16x2_to_f2(xr[c2]), ww = bf16x2_to_f2(w3[o * (H / 2) + c2]);
        s = fmaf(xx.x, ww.x, s);
        s = fmaf(xx.y, ww.y, s);
      }
#pragma unroll
      for (int of = 8; of > 0; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
      if (sl == 0) {
        const float v = s + bias[o];
        myo[o] = v;
        if (ok) out[static_cast<size_t>(row) * O + o] = v;
      }
    }
  }
  __syncwarp();
  const float* o = myo;
  const float* oo = out_old + static_cast<size_t>(old_idx ? old_idx[row] : row) * O;
  const float inv_s = 1.f / static_cast<float>(kl_samples);
  float logp = 0.f, lp_old = 0.f, lp_new = 0.f;
  for (int k = sl; k < A; k += 16) {
    const float mu = o[k], sig = expf(o[A + k]) + ac.min_std;
    const float e = eps_pi[static_cast<size_t>(row) * A + k];
    const float xs = mu + sig * e;
    const float a = tanhf(xs);
    if (ok) {
      act_out[static_cast<size_t>(row) * act_ld + k] = a;
      if (act_h != nullptr) act_h[static_cast<size_t>(row) * act_ldh + k] = __float2bfloat16_rn(a);
    }
    const float log_sig = logf(sig);
    if (ac.clip_policy_sample) {
      const float xc = atanhf(fminf(fmaxf(a, -ac.action_clip), ac.action_clip));
      const float zz = (xc - mu) / sig;
      logp += -0.5f * zz * zz - log_sig - kHalfLog2Pi - fldj_(xc);
    } else {
      logp += -0.5f * e * e - log_sig - kHalfLog2Pi - fldj_(xs);
    }
  }
  for (int k = 0; k < A; ++k) {
    const float mu = o[k], sig = expf(o[A + k]) + ac.min_std, log_sig = logf(sig), inv_sig = 1.f / sig;
    const float mu_o = oo[k], sig_o = expf(oo[A + k]) + ac.min_std, log_sig_o = logf(sig_o);
    float gm = 0.f, gs = 0.f;
    for (int s = sl; s < kl_samples; s += 16) {
      const float eo = eps_kl[(static_cast<size_t>(s) * rows + row) * A + k];
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
      gmu_kl[static_cast<size_t>(row) * A + k] = gm * inv_s;
      gsig_kl[static_cast<size_t>(row) * A + k] = gs * inv_s;
    }
  }
#pragma unroll
  for (int of = 8; of > 0; of >>= 1) {
    logp += __shfl_xor_sync(0xffffffffu, logp, of);
    lp_old += __shfl_xor_sync(0xffffffffu, lp_old, of);
    lp_new += __shfl_xor_sync(0xffffffffu, lp_new, of);
  }
  if (sl == 0 && ok) {
    logp_out[row] = logp;
    kl_out[row] = (lp_old - lp_new) * inv_s;
  }
}

cudaError_t launch_actor_head_sample(const void* x_h, int ldx, int H, const void* w_bf16, int ldw, const float* bias, float* out,
                                     const float* out_old, const int* old_idx, const float* eps_pi, const float* eps_kl, int rows,
                                     int A, int kl_samples, const ActorConsts& ac, float* act_out, int act_ld, float* logp,
                                     float* kl, float* gmu, float* gsig, void* act_h, int act_ldh, cudaStream_t st) {
  if (!thin_enabled(2) || rows <= 0 || 2 * A > 64 || (H & 31) || (ldx & 1) || (ldw & 1) || x_h == nullptr) return cudaErrorNotSupported;
  const size_t smem = (static_cast<size_t>(2 * A) * (H / 2) + 16 * 2 * A) * 4;
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(actor_head_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  launch_k((actor_head_sample_kernel), dim3((rows + 15) / 16), dim3(256), smem, st, static_cast<const __nv_bfloat16*>(x_h), ldx, H,
           static_cast<const __nv_bfloat16*>(w_bf16), ldw, bias, out, out_old, old_idx, eps_pi, eps_kl, rows, A, kl_samples, ac,
           act_out, act_ld, logp, kl, gmu, gsig, static_cast<__nv_bfloat16*>(act_h), act_ldh);
  return cudaGetLastError();
}

// ---- policy gradient with the two thin dgrads around it --------------------------------------------------
// (1) dQ/da[r, k] = dz1[r, 0:Hc] . W1[0:Hc, Dc + k]      dz1: bf16 twin of the critic first layer's dz, W1 its shadow
// (2) actor_grad_kernel's row math (losses.cu)             -> dout [r, 2A] (fp32 + bf16 twin), metrics, bias gradient
// (3) dx[r, 0:Ha] = dout[r, 0:2A] . W3[0:2A, 0:Ha]         -> fp32 (input of the norm backward that follows)
// One warp per row; lane k owns action dimension k (A <= 32) in (2).
__global__ void __launch_bounds__(256)
actor_grad_thin_kernel(const __nv_bfloat16* __restrict__ dz1, int lddz, int Hc, const __nv_bfloat16* __restrict__ w1, int ldw1,
                       int Dc, const __nv_bfloat16* __restrict__ w3, int ldw3, int Ha, const float* __restrict__ out,
                       const float* __restrict__ eps_pi, const float* __restrict__ act, int act_ld,
                       const float* __restrict__ logp_in, const float* __restrict__ kl_in, const float* __restrict__ q_in,
                       const float* __restrict__ gmu_kl, const float* __restrict__ gsig_kl, const float* __restrict__ actor_params,
                       int rows, int A, float inv_rows, ActorConsts ac, float* __restrict__ dout, float* __restrict__ actor_grads,
                       float* __restrict__ metrics, __nv_bfloat16* __restrict__ dout_h, int ldh, float* __restrict__ dbias,
                       float* __restrict__ dx) {
  REPPO_PDL_PROLOGUE();
  extern __shared__ float smf[];
  const int O = 2 * A;
  float* sw1 = smf;                       // [A][Hc]   W1[c, Dc + k] as fp32
  float* sw3 = sw1 + A * Hc;              // [O][Ha]   W3[o, c] as fp32
  float* scol = sw3 + O * Ha;             // [8][O]
  float* smet = scol + 8 * O;             // [8][8]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < A * Hc; i += blockDim.x) {
    const int k = i / Hc, c = i - k * Hc;
    sw1[i] = __bfloat162float(w1[static_cast<size_t>(c) * ldw1 + Dc + k]);
  }
  for (int i = threadIdx.x; i < O * Ha; i += blockDim.x) {
    const int o = i / Ha, c = i - o * Ha;
    sw3[i] = __bfloat162float(w3[static_cast<size_t>(o) * ldw3 + c]);
  }
  const int row = blockIdx.x * 8 + warp;
  const float alpha = expf(actor_params[0]), beta = expf(actor_params[1]);
  float m[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = lane; j < O; j += 32) scol[warp * O + j] = 0.f;
  __syncthreads();
  if (row < rows) {
    // (1) this lane's slice of the row of dz1, then A dot products
    float dact = 0.f;   // lane k < A ends up with dQ/da_k
    {
      const __nv_bfloat16* dr = dz1 + static_cast<size_t>(row) * lddz;
      for (int k = 0; k < A; ++k) {
        float s = 0.f;
        for (int c = lane; c < Hc; c += 32) s = fmaf(__bfloat162float(dr[c]), sw1[k * Hc + c], s);
        s = warp_sum(s);
        if (lane == k) dact = s;
      }
    }
    // (2)
    const float logp = logp_in[row], kl = kl_in[row], q = q_in[row];
    float w_path, w_kl;
    if (ac.kl_mode == KL_CLIPPED) { w_path = kl < ac.kl_bound ? 1.f : 0.f; w_kl = 1.f - w_path; }
    else if (ac.kl_mode == KL_FULL) { w_path = 1.f; w_kl = 1.f; }
    else { w_path = 1.f; w_kl = 0.f; }
    const float g_logp = inv_rows * w_path * alpha;
    const float g_lpnew = -inv_rows * w_kl * beta * ac.reduce_kl;
    const float* o = out + static_cast<size_t>(row) * O;
    float dmu = 0.f, dls = 0.f, abs_a = 0.f;
    if (lane < A) {
      const int k = lane;
      const float mu = o[k], ex = expf(o[A + k]), sig = ex + ac.min_std;
      const float e = eps_pi[static_cast<size_t>(row) * A + k];
      const float a = act[static_cast<size_t>(row) * act_ld + k];
      const float da = dact * (1.f - a * a);
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
      dmu = da + g_logp * dlp_mu + g_lpnew * gmu_kl[static_cast<size_t>(row) * A + k];
      dls = (da * e + g_logp * dlp_sig + g_lpnew * gsig_kl[static_cast<size_t>(row) * A + k]) * ex;
      dout[static_cast<size_t>(row) * O + k] = dmu;
      dout[static_cast<size_t>(row) * O + A + k] = dls;
      if (dout_h != nullptr) {
        dout_h[static_cast<size_t>(row) * ldh + k] = __float2bfloat16_rn(dmu);
        dout_h[static_cast<size_t>(row) * ldh + A + k] = __float2bfloat16_rn(dls);
      }
      scol[warp * O + k] = dmu;
      scol[warp * O + A + k] = dls;
      abs_a = fabsf(a);
    }
    abs_a = warp_sum(abs_a);
    m[0] = w_path * (alpha * logp - q) + w_kl * beta * ac.reduce_kl * kl;
    m[1] = kl; m[2] = -logp; m[3] = q; m[4] = abs_a / static_cast<float>(A); m[5] = ac.target_entropy - logp; m[6] = 1.f;
    // (3) dx = dout . W3 with dout rounded to bf16 like the GEMM operand
    const float dmu_b = bf16_round(dmu), dls_b = bf16_round(dls);
    float* dxr = dx + static_cast<size_t>(row) * Ha;
    for (int c0 = 0; c0 < Ha; c0 += 32) {
      const int c = c0 + lane;
      float s = 0.f;
      for (int k = 0; k < A; ++k) {
        const float a_mu = __shfl_sync(0xffffffffu, dmu_b, k), a_ls = __shfl_sync(0xffffffffu, dls_b, k);
        if (c < Ha) {
          s = fmaf(a_mu, sw3[k * Ha + c], s);
          s = fmaf(a_ls, sw3[(A + k) * Ha + c], s);
        }
      }
      if (c < Ha) dxr[c] = s;
    }
  }
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < 7; ++j) smet[warp * 8 + j] = m[j];
  __syncthreads();
  if (dbias != nullptr) {
    for (int j = threadIdx.x; j < O; j += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += scol[w * O + j];
      atomicAdd(dbias + j, t);
    }
  }
  if (threadIdx.x == 0) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < 8; ++w)
#pragma unroll
      for (int j = 0; j < 7; ++j) v[j] += smet[w * 8 + j];
    const float s_path = v[0] * inv_rows, s_kl = v[1] * inv_rows, s_ent = v[2] * inv_rows, s_q = v[3] * inv_rows;
    const float s_abs = v[4] * inv_rows, s_ent_t = v[5] * inv_rows, frac = v[6] * inv_rows;
    const float ent_loss = alpha * s_ent_t;
    const float lag_loss = -beta * (s_kl - ac.kl_bound * frac);
    float total = s_path;
    if (ac.update_entropy) { total += ent_loss; atomicAdd(actor_grads + 0, ent_loss); }
    if (ac.update_kl) { total += lag_loss; atomicAdd(actor_grads + 1, lag_loss); }
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

cudaError_t launch_actor_grad_thin(const void* dz1_h, int lddz, int Hc, const void* w1_bf16, int ldw1, int Dc, const void* w3_bf16,
                                   int ldw3, int Ha, const float* out, const float* eps_pi, const float* act, int act_ld,
                                   const float* logp, const float* kl, const float* q, const float* gmu, const float* gsig,
                                   const float* actor_params, int rows, int A, float inv_rows, const ActorConsts& ac, float* dout,
                                   float* actor_grads, float* metrics, void* dout_h, int ldh, float* dbias, float* dx,
                                   cudaStream_t st) {
  if (!thin_enabled(4) || rows <= 0 || A > 32 || dz1_h == nullptr || dx == nullptr) return cudaErrorNotSupported;
  const size_t smem = (static_cast<size_t>(A) * Hc + static_cast<size_t>(2 * A) * Ha + 8 * 2 * A + 64) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(actor_grad_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  launch_k((actor_grad_thin_kernel), dim3((rows + 7) / 8), dim3(256), smem, st, static_cast<const __nv_bfloat16*>(dz1_h), lddz, Hc,
           static_cast<const __nv_bfloat16*>(w1_bf16), ldw1, Dc, static_cast<const __nv_bfloat16*>(w3_bf16), ldw3, Ha, out, eps_pi,
           act, act_ld, logp, kl, q, gmu, gsig, actor_params, rows, A, inv_rows, ac, dout, actor_grads, metrics,
           static_cast<__nv_bfloat16*>(dout_h), ldh, dbias, dx);
  return cudaGetLastError();
}

}  // namespace reppo
