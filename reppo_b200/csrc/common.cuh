// Shared device helpers for the REPPO update kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define REPPO_DEVINL __device__ __forceinline__

namespace reppo {

constexpr int kWarp = 32;

REPPO_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
REPPO_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
REPPO_DEVINL float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; every thread gets the result
REPPO_DEVINL float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

REPPO_DEVINL float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
REPPO_DEVINL float swishf_(float x) { return x * sigmoidf_(x); }
// d/dx [x * sigmoid(x)]
REPPO_DEVINL float swish_grad_(float x) {
  const float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
// softplus as torch / jax.nn.softplus compute it (log1p(exp(x)), linear above the threshold)
REPPO_DEVINL float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }

REPPO_DEVINL void atomic_max_float(float* addr, float v) {
  // valid for any sign: monotone integer views
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
REPPO_DEVINL void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

}  // namespace reppo
