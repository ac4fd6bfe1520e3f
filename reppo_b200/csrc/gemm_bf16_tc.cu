// tcgen05 / TMEM / TMA GEMM for the Linear layers of the REPPO MLPs (REPPO_GEMM_BF16 mode).
//
//   D[M,N] (fp32) = A · B  with bf16 operands staged in shared memory by TMA (128-byte swizzle),
//   fp32 accumulators in tensor memory, issued by one elected thread with tcgen05.mma
//   (cta_group::1, UMMA 128 x BLOCK_N x 16), read back with tcgen05.ld for the epilogue.
//
// Three operand arrangements cover forward, dgrad and wgrad of y = x W^T + b
// (src/networks/jax_models.py:35-49, torch_models.py:71-79):
//   A K-major,  B K-major  : forward  y  = x  · W^T    A = x  [rows,in],  B = W [out,in]  (reduction dim contiguous)
//   A K-major,  B MN-major : dgrad    dx = dy · W      A = dy [rows,out], B = W [out,in]  (same bf16 shadow, no W^T copy)
//   A MN-major, B MN-major : wgrad    dW[out,in] = dy^T · x    A = dy [rows,out], B = x [rows,in]  (reduction dim = rows)
// Warp roles (320 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer,
// warps 2..9 = epilogue (warp w owns TMEM lanes 32*(w%4)..+32; the two warps of a lane quarter split the columns).
// Pipeline: NUM_STAGES-deep smem ring with full/empty mbarriers (TMA -> MMA), one
// tmem_full mbarrier (MMA -> epilogue).  One output tile per CTA; split-K over gridDim.z for
// wgrad (fp32 atomics into the zeroed gradient arena).  Epilogue: tcgen05.ld gives every thread
// one accumulator ROW, which is transposed through the (by then idle) pipeline shared memory so
// that global stores are row-contiguous 128 B / 512 B segments.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "knobs.h"
#include "launch.h"
#include "tc_ptx.cuh"
#include "tmap.h"

namespace reppo {

using namespace tcx;

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

// Operand kind.  KIND_TF32 is the reference-precision mode on tensor cores (REPPO_GEMM_TF32): the fp32 activations and an
// aligned fp32 copy of the weights are read as they are by TMA and multiplied by tcgen05.mma.kind::tf32 (10-bit mantissa
// products, fp32 accumulation).  Everything scales with the element size: a 128-byte swizzle row holds 64 bf16 or 32 fp32
// values (= one k-block of a K-major operand, one MN atom of an MN-major one), one MMA consumes 32 bytes of K (16 / 8).
enum { KIND_BF16 = 0, KIND_TF32 = 1 };
template <int KIND> struct KindTraits {
  static constexpr int kElem = KIND == KIND_TF32 ? 4 : 2;
  static constexpr int kBlockK = 128 / kElem;
  static constexpr int kUmmaK = 32 / kElem;
  static constexpr int kAtomMN = 128 / kElem;
  static constexpr uint32_t kFormat = KIND == KIND_TF32 ? 2u : 1u;   // F16F32Format: 1 = BF16, 2 = TF32
};

// instruction descriptor (cute::UMMA::InstrDescriptor): {bf16 | tf32} x same -> f32, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major, uint32_t fmt = 1u) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(TC_BLOCK_M >> 4) << 24);
}

template <int BLOCK_N, int STAGES>
struct TcSmem {
  static constexpr int kABytes = TC_BLOCK_M * TC_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kStatOffset = kBarOffset + (2 * STAGES + 1) * 8 + 16;      // fused-norm row statistics
  // part[2][2][128], cta_stat[2][128] (unused), rs/mu[8][32], colc[3][BLOCK_N], allstat[8][2][128]
  static constexpr int kColcOffsetF = 4 * 128 + 2 * 128 + 2 * 8 * 32;
  static constexpr int kAllStatOffsetF = kColcOffsetF + 3 * BLOCK_N;
  static constexpr int kStatBytes = (kAllStatOffsetF + 8 * 2 * 128) * 4;
  static constexpr int kTotal = kStatOffset + kStatBytes + 1024;                   // + alignment slack
};

// mode: GEMM_STORE / GEMM_ACCUM / GEMM_ATOMIC.  A_MN / B_MN: operand is MN-major (reduction dim is the outer dim).
// NORM != NORM_NONE: forward Linear with the whole FCNN block fused into the epilogue
//   z = x W^T + b (stored fp32 for the backward), act_out = bf16( swish( norm(z) ) ), rstd / mean per row.
// The N = H columns of a 128-row slab are spread over the CTAs of one thread-block CLUSTER (gridDim.x CTAs of
// BLOCK_N columns); every CTA reduces its own columns and the per-row partial (sum, sum of squares) are exchanged
// through distributed shared memory (mapa + ld.shared::cluster) between two cluster barriers.
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int NORM, int KIND = KIND_BF16>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    float* __restrict__ C, int ldc, const float* __restrict__ bias, int M, int N, int K, int k_blocks_per_split,
                    int mode, __nv_bfloat16* __restrict__ act_out, int act_ld, NormArgs na) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  using S = TcSmem<BLOCK_N, STAGES>;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BLOCK_M, n0 = blockIdx.x * BLOCK_N;
  using KT = KindTraits<KIND>;
  static_assert(KIND == KIND_BF16 || NORM == NORM_NONE, "the fused-norm epilogue writes bf16 twins: bf16 kind only");
  const int total_k_blocks = (K + KT::kBlockK - 1) / KT::kBlockK;
  const int kb_begin = blockIdx.z * k_blocks_per_split;
  const int kb_end = min(total_k_blocks, kb_begin + k_blocks_per_split);
  const int num_kb = kb_end - kb_begin;

  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
    // TMEM: BLOCK_N fp32 columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail
  REPPO_PDL_PROLOGUE();

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * S::kStageBytes;
      uint8_t* sb = sa + S::kABytes;
      mbar_expect_tx(&full_bar[s], S::kStageBytes);
      const int k0 = (kb_begin + i) * KT::kBlockK;
      // K-major tile [rows][128 B of k]: coordinates {k, row}.  MN-major: atoms of [kBlockK k-rows][128 B of mn], coordinates {mn, k-row}.
      if (!A_MN) {
        tma_load_2d(&tmap_a, &full_bar[s], sa, k0, m0);
      } else {
#pragma unroll
        for (int a = 0; a < TC_BLOCK_M / KT::kAtomMN; ++a)
          tma_load_2d(&tmap_a, &full_bar[s], sa + a * (KT::kBlockK * 128), m0 + a * KT::kAtomMN, k0);
      }
      if (!B_MN) {
        tma_load_2d(&tmap_b, &full_bar[s], sb, k0, n0);
      } else {
#pragma unroll
        for (int b = 0; b < BLOCK_N / KT::kAtomMN; ++b)
          tma_load_2d(&tmap_b, &full_bar[s], sb + b * (KT::kBlockK * 128), n0 + b * KT::kAtomMN, k0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(BLOCK_N, A_MN, B_MN, KT::kFormat);
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
      const uint32_t sb = sa + S::kABytes;
#pragma unroll
      for (int k = 0; k < KT::kBlockK / KT::kUmmaK; ++k) {
        // K-major SW128: 8-row groups 1024 B apart; advance 32 B per UMMA_K inside the swizzle row.
        // MN-major SW128: LBO = stride between 128-byte-wide MN atoms, SBO = stride between 8-row K groups;
        // advance kUmmaK k-rows (2048 / 1024 B) per UMMA_K.
        // MN-major tf32: the 32-byte-atom flavour of the 128-byte swizzle (the only one tcgen05 takes for MN-major 32-bit
        // operands): swizzle groups of 4 k-rows, SBO = 512 B
        constexpr uint32_t kMnSbo = KIND == KIND_TF32 ? 512 : 1024, kMnType = KIND == KIND_TF32 ? 1 : 2;
        const uint64_t da = A_MN ? make_smem_desc(sa + k * (KT::kUmmaK * 128), KT::kBlockK * 128, kMnSbo, kMnType)
                                 : make_smem_desc(sa + k * 32, 16, 1024);
        const uint64_t db = B_MN ? make_smem_desc(sb + k * (KT::kUmmaK * 128), KT::kBlockK * 128, kMnSbo, kMnType)
                                 : make_smem_desc(sb + k * 32, 16, 1024);
        if constexpr (KIND == KIND_TF32) tc_mma_tf32(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        else tc_mma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(&empty_bar[s]);          // frees the smem slot when these MMAs retire
    }
    tc_commit(tmem_full_bar);            // accumulator complete
  } else if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====
    // 8 warps: warp w may only touch TMEM lanes 32*(w%4)..+32, so two warps share each lane quarter and split the
    // tile's columns between them (the epilogue, not the MMA, bounds these short-K GEMMs).
    float4 bias_pre = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (NORM == NORM_NONE) {
      // this thread's 4 bias values of the coalesced store layout, requested while the MMAs run
      if (bias != nullptr && blockIdx.z == 0) {
        const int colp = n0 + ((warp - 2) >> 2) * (BLOCK_N / 2) + (lane % (BLOCK_N / 8)) * 4;
        if (colp + 0 < N) bias_pre.x = bias[colp + 0];
        if (colp + 1 < N) bias_pre.y = bias[colp + 1];
        if (colp + 2 < N) bias_pre.z = bias[colp + 2];
        if (colp + 3 < N) bias_pre.w = bias[colp + 3];
      }
    }
    if constexpr (NORM != NORM_NONE) {
      // per-column constants are fetched while the MMAs run: after the cluster barrier below L1 is cold and every
      // global load would be an L2 round trip on the critical path
      float* colc = reinterpret_cast<float*>(smem + S::kStatOffset) + S::kColcOffsetF;
      for (int t = (warp - 2) * 32 + lane; t < BLOCK_N; t += 256) {
        colc[t] = bias[n0 + t];
        colc[BLOCK_N + t] = na.gamma[n0 + t];
        colc[2 * BLOCK_N + t] = (NORM == NORM_LAYER) ? na.beta[n0 + t] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // all TMA loads were consumed by MMAs that have retired: the pipeline stages are free to reuse
    constexpr int kCols = BLOCK_N / 2;           // columns per epilogue warp
    constexpr int kStgLd = kCols + 4;
    static_assert(8 * 32 * kStgLd * 4 <= STAGES * S::kStageBytes, "staging does not fit in the pipeline smem");
    const int ew = warp - 2;
    float* stg = reinterpret_cast<float*>(smem) + ew * 32 * kStgLd;
    const int lane_base = (warp & 3) * 32;
    const int cbase = (ew >> 2) * kCols;
    float s1 = 0.f, s2 = 0.f;   // fused norm: this thread's row, this warp's columns
#pragma unroll 1
    for (int c0 = 0; c0 < kCols; c0 += 32) {
      float v[32];
      if (num_kb > 0) {
        tc_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + cbase + c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if constexpr (NORM != NORM_NONE) {
        // bias first: the statistics are those of z = x W^T + b
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += (reinterpret_cast<const float*>(smem + S::kStatOffset) + S::kColcOffsetF)[cbase + c0 + j];
          s1 += v[j];
          s2 += v[j] * v[j];
        }
      }
      float4* dst = reinterpret_cast<float4*>(stg + lane * kStgLd + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if constexpr (NORM != NORM_NONE) {
      float* part = reinterpret_cast<float*>(smem + S::kStatOffset);        // [2 stats][2 halves][128 rows]
      const int rrow = lane_base + lane, half = ew >> 2;
      part[(0 * 2 + half) * 128 + rrow] = s1;
      part[(1 * 2 + half) * 128 + rrow] = s2;
      asm volatile("bar.sync 1, 256;" ::: "memory");                       // the 8 epilogue warps
      if (half == 0) {
        // push this CTA's per-row partial statistics into every peer of the cluster (remote stores; published by the
        // cluster barrier) instead of pulling 2 x cluster-size remote values per thread after it
        const float u1 = part[0 * 128 + rrow] + part[1 * 128 + rrow];
        const float u2 = part[2 * 128 + rrow] + part[3 * 128 + rrow];
        float* allstat = part + S::kAllStatOffsetF;
        const uint32_t ncl = gridDim.x, me = blockIdx.x;
        for (uint32_t r = 0; r < ncl; ++r) {
          st_dsmem_f32(allstat + (me * 2 + 0) * 128 + rrow, r, u1);
          st_dsmem_f32(allstat + (me * 2 + 1) * 128 + rrow, r, u2);
        }
      }
    }
    __syncwarp();
    const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
    if constexpr (NORM == NORM_NONE) {
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (mode != GEMM_ATOMIC) &&
                        (act_out == nullptr || (act_ld & 3) == 0);
    if (vec_ok) {
      constexpr int kLanesPerRow = kCols / 4;            // 16 (BLOCK_N = 128) or 8 (BLOCK_N = 64)
      constexpr int kRowsPerIter = 32 / kLanesPerRow;    // 2 or 4
      const int cl = (lane % kLanesPerRow) * 4;
      const int col = n0 + cbase + cl;
      const float4 bv = bias_pre;   // zero when there is no bias / not the first K split
#pragma unroll 4
      for (int r = lane / kLanesPerRow; r < 32; r += kRowsPerIter) {
        const int row = m0 + lane_base + r;
        if (row >= M || col >= N) continue;
        float4 x = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
        x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
        float* cp = C + static_cast<size_t>(row) * ldc + col;
        if (col + 3 < N) {
          if (mode == GEMM_ACCUM) {
            const float4 o = *reinterpret_cast<const float4*>(cp);
            x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
          }
          *reinterpret_cast<float4*>(cp) = x;
          if (act_out != nullptr) {   // fused epilogue: swish(y) as the bf16 operand of the next GEMM
            const __nv_bfloat162 lo = __floats2bfloat162_rn(fswish(x.x), fswish(x.y)), hi = __floats2bfloat162_rn(fswish(x.z), fswish(x.w));
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&lo);
            u.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(act_out + static_cast<size_t>(row) * act_ld + col) = u;
          }
        } else {
          const float xs[4] = {x.x, x.y, x.z, x.w};
          for (int j = 0; j < 4 && col + j < N; ++j) {
            const float o = (mode == GEMM_ACCUM) ? cp[j] + xs[j] : xs[j];
            cp[j] = o;
            if (act_out != nullptr) act_out[static_cast<size_t>(row) * act_ld + col + j] = __float2bfloat16_rn(swishf_(o));
          }
        }
      }
    } else {
      // ragged leading dimension or atomics: scalar, lanes sweep consecutive columns (128 B per row per access)
#pragma unroll 1
      for (int r = 0; r < 32; ++r) {
        const int row = m0 + lane_base + r;
        if (row >= M) break;
        float* crow = C + static_cast<size_t>(row) * ldc;
#pragma unroll
        for (int cc = 0; cc < kCols; cc += 32) {
          const int col = n0 + cbase + cc + lane;
          if (col < N) {
            float x = stg[r * kStgLd + cc + lane];
            if (add_bias) x += bias[col];
            if (mode == GEMM_STORE) crow[col] = x;
            else if (mode == GEMM_ACCUM) { x += crow[col]; crow[col] = x; }
            else atomicAdd(crow + col, x);
            if (act_out != nullptr && mode != GEMM_ATOMIC) act_out[static_cast<size_t>(row) * act_ld + col] = __float2bfloat16_rn(swishf_(x));
          }
        }
      }
    }
    }  // NORM == NORM_NONE
    tc_fence_before();
  }
  if constexpr (NORM != NORM_NONE) {
    // ---- every thread of every CTA of the cluster: row statistics are published ----
    cluster_sync_all();
    if (warp >= 2) {
      constexpr int kCols = BLOCK_N / 2;
      constexpr int kStgLd = kCols + 4;
      const int ew = warp - 2;
      const float* stg = reinterpret_cast<const float*>(smem) + ew * 32 * kStgLd;
      const int lane_base = (warp & 3) * 32, cbase = (ew >> 2) * kCols, rrow = lane_base + lane;
      float* part = reinterpret_cast<float*>(smem + S::kStatOffset);
      float* rs = part + 6 * 128 + ew * 32;          // per-warp copies of the row statistics
      float* mu = part + 6 * 128 + 8 * 32 + ew * 32;
      float t1 = 0.f, t2 = 0.f;
      const uint32_t ncta = gridDim.x;                // cluster spans the whole row
      const float* allstat = part + S::kAllStatOffsetF;
#pragma unroll
      for (uint32_t r = 0; r < 8; ++r)
        if (r < ncta) { t1 += allstat[(r * 2 + 0) * 128 + rrow]; t2 += allstat[(r * 2 + 1) * 128 + rrow]; }
      const float inv_n = 1.f / static_cast<float>(N);
      float mean = 0.f, rstd;
      if (NORM == NORM_RMS) {
        rstd = rsqrtf(t2 * inv_n + na.eps);
      } else {
        mean = t1 * inv_n;
        rstd = rsqrtf(t2 * inv_n - mean * mean + na.eps);   // flax use_fast_variance
      }
      rs[lane] = rstd; mu[lane] = mean;
      if (blockIdx.x == 0 && (ew >> 2) == 0 && m0 + rrow < M) {
        na.rstd_out[m0 + rrow] = rstd;
        if (NORM == NORM_LAYER) na.mean_out[m0 + rrow] = mean;
      }
      __syncwarp();
      constexpr int kLanesPerRow = kCols / 4, kRowsPerIter = 32 / kLanesPerRow;
      const int cl = (lane % kLanesPerRow) * 4;
      const int col = n0 + cbase + cl;
      float g[4], be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* colc = part + S::kColcOffsetF;
        g[j] = colc[BLOCK_N + cbase + cl + j];
        be[j] = colc[2 * BLOCK_N + cbase + cl + j];
      }
#pragma unroll 4
      for (int r = lane / kLanesPerRow; r < 32; r += kRowsPerIter) {
        const int row = m0 + lane_base + r;
        if (row >= M) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);   // z (bias included)
        *reinterpret_cast<float4*>(C + static_cast<size_t>(row) * ldc + col) = x;
        const float rr = rs[r], mm = mu[r];
        // bf16-operand mode: SFU sigmoid (ex2.approx / rcp.approx); the exact expf + IEEE division made this loop issue-bound
        const float y0 = fswish((x.x - mm) * rr * g[0] + be[0]), y1 = fswish((x.y - mm) * rr * g[1] + be[1]);
        const float y2 = fswish((x.z - mm) * rr * g[2] + be[2]), y3 = fswish((x.w - mm) * rr * g[3] + be[3]);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1), hi = __floats2bfloat162_rn(y2, y3);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&lo);
        u.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(act_out + static_cast<size_t>(row) * act_ld + col) = u;
      }
    }
    // (no second cluster barrier: peers only WRITE into this CTA's shared memory, and only before the barrier above)
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BLOCK_N));
  }
}

// ---- host side -------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int KIND = KIND_BF16>
static cudaError_t launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int ldc, const float* bias, int M, int N,
                             int K, int split_k, int mode, __nv_bfloat16* act_out, int act_ld, cudaStream_t st) {
  using S = TcSmem<BLOCK_N, STAGES>;
  constexpr int TC_BLOCK_K = KindTraits<KIND>::kBlockK;   // shadows the bf16 constant below
  static unsigned long long configured = 0ull;   // one bit per device: the attribute is per device
  auto kern = gemm_bf16_tc_kernel<BLOCK_N, STAGES, A_MN, B_MN, NORM_NONE, KIND>;
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!((configured >> (dev__ & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    configured |= 1ull << (dev__ & 63);
  }
  const int total_kb = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;
  if (split_k < 1) split_k = 1;
  if (split_k > total_kb) split_k = total_kb;
  const int per = (total_kb + split_k - 1) / split_k;
  split_k = (total_kb + per - 1) / per;
  if (split_k > 1) mode = GEMM_ATOMIC;
  dim3 grid((N + BLOCK_N - 1) / BLOCK_N, (M + TC_BLOCK_M - 1) / TC_BLOCK_M, split_k);
  launch_k((kern), dim3(grid), dim3(TC_THREADS), S::kTotal, st, ta, tb, C, ldc, bias, M, N, K, per, mode, act_out, act_ld, NormArgs{});
  return cudaGetLastError();
}

// forward Linear + bias + RMS/LayerNorm + swish in one kernel; one cluster of N / BLOCK_N CTAs per 128-row slab
template <int BLOCK_N, int NORM>
static cudaError_t launch_tc_norm(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int ldc, const float* bias, int M,
                                  int N, int K, __nv_bfloat16* act_out, int act_ld, const NormArgs& na, cudaStream_t st) {
  using S = TcSmem<BLOCK_N, 4>;
  static unsigned long long configured = 0ull;   // one bit per device: the attribute is per device
  auto kern = gemm_bf16_tc_kernel<BLOCK_N, 4, false, false, NORM>;
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!((configured >> (dev__ & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    configured |= 1ull << (dev__ & 63);
  }
  const int total_kb = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(N / BLOCK_N, (M + TC_BLOCK_M - 1) / TC_BLOCK_M, 1);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = N / BLOCK_N; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, C, ldc, bias, M, N, K, total_kb, static_cast<int>(GEMM_STORE), act_out, act_ld, na);
}

// y = x W^T + b with norm + swish fused (see the kernel header).  Returns cudaErrorNotSupported when the shape does
// not fit the cluster scheme (the caller then runs the unfused GEMM + row kernel).
cudaError_t launch_gemm_bf16_tc_norm(int norm, int M, int N, int K, const void* A, int lda, const void* B, int ldb, float* C,
                                     int ldc, const float* bias, void* act_out_v, int act_ld, const NormArgs& na,
                                     cudaStream_t st) {
  if (norm == NORM_NONE || bias == nullptr || act_out_v == nullptr) return cudaErrorNotSupported;
  // large shapes: the persistent CTA-pair GEMM + the row kernel beats this one-tile-per-CTA fused kernel
  if (!knobs().fuse_norm_large && gemm_bf16_tc2_eligible(false, false, M, N, K, GEMM_STORE, 1)) return cudaErrorNotSupported;
  if ((lda & 7) || (ldb & 7) || (ldc & 3) || (act_ld & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15))
    return cudaErrorNotSupported;
  int block_n = 0;
  if (knobs().norm_bn128 && N % 128 == 0 && N / 128 <= 8) block_n = 128;
  else if (N % 64 == 0 && N / 64 <= 8) block_n = 64;       // portable cluster size <= 8
  else if (N % 128 == 0 && N / 128 <= 8) block_n = 128;
  if (block_n == 0) return cudaErrorNotSupported;
  CUtensorMap ta, tb;
  if (!tmap_encode_2d(&ta, A, M, K, lda, TC_BLOCK_K, TC_BLOCK_M) || !tmap_encode_2d(&tb, B, N, K, ldb, TC_BLOCK_K, block_n))
    return cudaErrorInvalidValue;
  __nv_bfloat16* act_out = static_cast<__nv_bfloat16*>(act_out_v);
  if (block_n == 64) {
    return norm == NORM_RMS ? launch_tc_norm<64, NORM_RMS>(ta, tb, C, ldc, bias, M, N, K, act_out, act_ld, na, st)
                            : launch_tc_norm<64, NORM_LAYER>(ta, tb, C, ldc, bias, M, N, K, act_out, act_ld, na, st);
  }
  return norm == NORM_RMS ? launch_tc_norm<128, NORM_RMS>(ta, tb, C, ldc, bias, M, N, K, act_out, act_ld, na, st)
                          : launch_tc_norm<128, NORM_LAYER>(ta, tb, C, ldc, bias, M, N, K, act_out, act_ld, na, st);
}

// A, B: bf16 device pointers, row-major with leading dimensions lda / ldb (elements, multiples of 8).
//  a_mn == false: A stored [M, K]   | a_mn == true: A stored [K, M]
//  b_mn == false: B stored [N, K]   | b_mn == true: B stored [K, N]
cudaError_t launch_gemm_bf16_tc(bool a_mn, bool b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                float* C, int ldc, const float* bias, int mode, int split_k, void* act_out_v, int act_ld,
                                cudaStream_t st) {
  __nv_bfloat16* act_out = static_cast<__nv_bfloat16*>(act_out_v);
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if ((lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return cudaErrorInvalidValue;
  if (a_mn && !b_mn) return cudaErrorInvalidValue;  // arrangement not used by the MLP
  {
    // shapes that fill the machine with 256-row pair tiles: persistent cta_group::2 kernel
    const cudaError_t e2 = launch_gemm_bf16_tc2(a_mn, b_mn, M, N, K, A, lda, B, ldb, C, ldc, bias, mode, split_k, act_out_v, act_ld, st);
    if (e2 != cudaErrorNotSupported) return e2;
  }
  CUtensorMap ta, tb;
  // narrow tiles when 128-wide ones would leave most of the 148 SMs idle
  const int tiles128 = ((M + TC_BLOCK_M - 1) / TC_BLOCK_M) * ((N + 127) / 128) * (split_k > 0 ? split_k : 1);
  const bool wide = (N > 64) && (tiles128 >= knobs().tc_wide_min);
  const int block_n = wide ? 128 : 64;
  const bool ok_a = a_mn ? tmap_encode_2d(&ta, A, K, M, lda, 64, TC_BLOCK_K) : tmap_encode_2d(&ta, A, M, K, lda, TC_BLOCK_K, TC_BLOCK_M);
  const bool ok_b = b_mn ? tmap_encode_2d(&tb, B, K, N, ldb, 64, TC_BLOCK_K) : tmap_encode_2d(&tb, B, N, K, ldb, TC_BLOCK_K, block_n);
  if (!ok_a || !ok_b) return cudaErrorInvalidValue;
#define REPPO_TC_DISPATCH(AMN, BMN)                                                                           \
  return wide ? launch_tc<128, 4, AMN, BMN>(ta, tb, C, ldc, bias, M, N, K, split_k, mode, act_out, act_ld, st) \
              : launch_tc<64, 4, AMN, BMN>(ta, tb, C, ldc, bias, M, N, K, split_k, mode, act_out, act_ld, st)
  if (!a_mn && !b_mn) { REPPO_TC_DISPATCH(false, false); }
  if (!a_mn && b_mn) { REPPO_TC_DISPATCH(false, true); }
  REPPO_TC_DISPATCH(true, true);
#undef REPPO_TC_DISPATCH
}

// The same three arrangements on fp32 operands read as tf32 (REPPO_GEMM_TF32).  A, B: fp32 device pointers, leading
// dimensions in elements.  cudaErrorNotSupported when TMA cannot address an operand (base not 16-byte aligned, leading
// dimension not a multiple of 4 floats): the caller falls back to the FFMA kernel for that GEMM.
cudaError_t launch_gemm_tf32_tc(bool a_mn, bool b_mn, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                                float* C, int ldc, const float* bias, int mode, int split_k, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return cudaErrorNotSupported;
  if (a_mn && !b_mn) return cudaErrorNotSupported;
  using KT = KindTraits<KIND_TF32>;
  CUtensorMap ta, tb;
  const int tiles128 = ((M + TC_BLOCK_M - 1) / TC_BLOCK_M) * ((N + 127) / 128) * (split_k > 0 ? split_k : 1);
  const bool wide = (N > 64) && (tiles128 >= knobs().tc_wide_min);
  const int block_n = wide ? 128 : 64;
  const bool ok_a = a_mn ? tmap_encode_2d(&ta, A, K, M, lda, KT::kAtomMN, KT::kBlockK, 4, true) : tmap_encode_2d(&ta, A, M, K, lda, KT::kBlockK, TC_BLOCK_M, 4);
  const bool ok_b = b_mn ? tmap_encode_2d(&tb, B, K, N, ldb, KT::kAtomMN, KT::kBlockK, 4, true) : tmap_encode_2d(&tb, B, N, K, ldb, KT::kBlockK, block_n, 4);
  if (!ok_a || !ok_b) return cudaErrorNotSupported;
#define REPPO_TF32_DISPATCH(AMN, BMN)                                                                                        \
  return wide ? launch_tc<128, 4, AMN, BMN, KIND_TF32>(ta, tb, C, ldc, bias, M, N, K, split_k, mode, nullptr, 0, st)         \
              : launch_tc<64, 4, AMN, BMN, KIND_TF32>(ta, tb, C, ldc, bias, M, N, K, split_k, mode, nullptr, 0, st)
  if (!a_mn && !b_mn) { REPPO_TF32_DISPATCH(false, false); }
  if (!a_mn && b_mn) { REPPO_TF32_DISPATCH(false, true); }
  REPPO_TF32_DISPATCH(true, true);
#undef REPPO_TF32_DISPATCH
}

}  // namespace reppo
