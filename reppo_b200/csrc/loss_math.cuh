// Row-wise math shared by the loss kernels (losses.cu) and the rollout kernels (rollout.cu).
#pragma once
#include "common.cuh"

namespace reppo {

constexpr int kMaxBinsPerLane = 8;  // num_bins <= 256
constexpr float kLog2 = 0.6931471805599453f;
constexpr float kHalfLog2Pi = 0.9189385332046727f;

template <int NB>
struct RowSoftmaxT {
  float p[NB];
  float lg[NB];
  float lse, value;
};
using RowSoftmax = RowSoftmaxT<kMaxBinsPerLane>;

// logits = head + bias_scale * zero_dist; softmax; value = p . support      (NB = bins per lane the caller unrolls for)
template <int NB>
REPPO_DEVINL void row_softmax(const float* __restrict__ head_row, const float* __restrict__ zero_dist,
                              const float* __restrict__ support, float bias_scale, int B, int lane, RowSoftmaxT<NB>& o) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int c = lane + 32 * j;
    o.lg[j] = (c < B) ? head_row[c] + bias_scale * zero_dist[c] : -INFINITY;
    m = fmaxf(m, o.lg[j]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int c = lane + 32 * j;
    o.p[j] = (c < B) ? expf(o.lg[j] - m) : 0.f;
    s += o.p[j];
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int c = lane + 32 * j;
    o.p[j] *= inv;
    if (c < B) v += o.p[j] * support[c];
  }
  o.value = warp_sum(v);
  o.lse = m + logf(s);
}

// the same with the row-invariant per-lane constants (bias_scale * zero_dist, support) already in registers: the
// kernels that walk many rows per warp load them once
template <int NB>
REPPO_DEVINL void row_softmax_pre(const float* __restrict__ head_row, const float (&zb)[NB], const float (&sup)[NB], int B,
                                  int lane, RowSoftmaxT<NB>& o) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int c = lane + 32 * j;
    o.lg[j] = (c < B) ? head_row[c] + zb[j] : -INFINITY;
    m = fmaxf(m, o.lg[j]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    o.p[j] = expf(o.lg[j] - m);   // exp(-inf) = 0 for the bins past B
    s += o.p[j];
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    o.p[j] *= inv;
    v += o.p[j] * sup[j];
  }
  o.value = warp_sum(v);
  o.lse = m + logf(s);
}

REPPO_DEVINL float fldj_(float x) { return 2.f * (kLog2 - x - softplusf_(-2.f * x)); }

}  // namespace reppo
