// Input-gradient GEMM with the FCNN block's backward fused into the tcgen05 epilogue (REPPO_GEMM_BF16 mode).
//
//   dx[rows, in] = dy[rows, out] . W[out, in]                     (A = dy K-major, B = W read MN-major)
//   NORM_NONE        : dz = (dx [+ addend]) * swish'(z)
//   NORM_RMS / LAYER : x = swish(norm(z)):  dy' = dx * swish'(y),  dzh = dy' * gamma,
//                      dz = rstd * (dzh - mean(dzh) - zh * mean(dzh * zh))           (mean(dzh) only for LayerNorm)
//   outputs: dz as the bf16 operand of the next dgrad / wgrad GEMM; column sums d gamma, d beta, d bias (atomics).
// It replaces a dgrad GEMM launch + a norm_act_bwd launch on the critical chain of the step
// (reference: autodiff of FCNN, src/networks/jax_models.py:109-124 / torch_models.py:71-79).
// The in = H columns of a 128-row slab are spread over one thread-block cluster (H / 64 CTAs); the per-row partial
// sums are pushed into every peer's shared memory (st.shared::cluster) and published by one cluster barrier.
// The saved z tile, rstd / mean and the optional addend are requested from global memory BEFORE the accumulator is
// ready, so their latency hides behind the TMA / MMA main loop.
#include <stdlib.h>

#include "kernels.h"
#include "launch.h"
#include "slab.h"
#include "tc_ptx.cuh"

namespace reppo {
namespace {

using namespace tcx;

constexpr int kStages = 4;
constexpr int kThreads = 320;
constexpr int kABytes = SB_M * SB_K * 2, kBBytes = SB_N * SB_K * 2, kStageBytes = kABytes + kBBytes;
constexpr int kRingBytes = kStages * kStageBytes;
constexpr int kBarOffset = kRingBytes;
constexpr int kStatOffset = kBarOffset + 128;
// part[2][2][128] | rowa[8][32] | rowb[8][32] | colpart[3][8][32] | colc[2][64] | allstat[8][2][128]
constexpr int kOffRowA = 512, kOffRowB = 768, kOffColPart = 1024, kOffColC = 1792, kOffAllStat = 1920, kStatFloats = 1920 + 2048;
constexpr int kSmem = kStatOffset + kStatFloats * 4 + 1024;
constexpr int kStgLd = 36;
static_assert(2 * 8 * 32 * kStgLd * 4 <= kRingBytes, "staging must fit in the ring");

struct BwdArgs {
  const float* z; int ldz;          // saved forward pre-activation [rows, N]
  const float* rstd; const float* mean;
  const float* gamma; const float* beta;
  const float* addend; int ld_add;  // optional fp32 [rows, N] added to the accumulator (second branch of a gradient sum)
  __nv_bfloat16* twin; int ld_twin; // dz, bf16
  float* g_gamma; float* g_beta; float* g_bias;
};

template <int NORM>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, BwdArgs args,
                        int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* stats = reinterpret_cast<float*>(smem + kStatOffset);
  constexpr bool kIsNorm = NORM != NORM_NONE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x, ncta = gridDim.x;
  const int m0 = blockIdx.y * SB_M, n0 = rank * SB_N;
  const int nkb = (K + SB_K - 1) / SB_K;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(SB_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  REPPO_PDL_PROLOGUE();

  const int ew = (warp - 2) & 7, half = (ew >> 2) & 1, lane_base = (warp & 3) * 32, cbase = half * 32;
  const int rrow = lane_base + lane;
  const int cl = (lane & 7) * 4, rg = lane >> 3;
  const int col = n0 + cbase + cl;
  float* stg = reinterpret_cast<float*>(smem) + ew * 32 * kStgLd;
  float* stg2 = stg + 8 * 32 * kStgLd;
  float* part = stats;
  float* rowa = stats + kOffRowA + ew * 32;
  float* rowb = stats + kOffRowB + ew * 32;
  float* colpart = stats + kOffColPart;
  float* colc = stats + kOffColC;       // gamma | beta of this CTA's columns
  float* allstat = stats + kOffAllStat;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int sg = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[sg], ph ^ 1);
        uint8_t* sa = smem + sg * kStageBytes;
        mbar_expect_tx(&full_bar[sg], kStageBytes);
        tma_load_2d(&tmap_a, &full_bar[sg], sa, kb * SB_K, m0);
        tma_load_2d(&tmap_b, &full_bar[sg], sa + kABytes, n0, kb * SB_K);   // [k][n] tile, coordinates {n, k}
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(true);
      for (int kb = 0; kb < nkb; ++kb) {
        const int sg = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[sg], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + sg * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < SB_K / SB_UMMA_K; ++k) {
          const uint64_t da = make_smem_desc(sa + k * (SB_UMMA_K * 2), 16, 1024);
          const uint64_t db = make_smem_desc(sb + k * (SB_UMMA_K * 128), SB_K * 128, 1024);
          tc_mma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[sg]);
      }
      tc_commit(tmem_full_bar);
    }
    __syncwarp();
  }

  float4 pre[8];                 // saved z (coalesced layout: rows rg + 4 i, columns col .. col + 3)
  float4 add[8];                 // optional addend
  float prow[8], prow2[8];       // rstd, mean of those rows
  if (warp >= 2) {
    // ---- part 0: global-memory operands of the epilogue, requested while the MMAs run ----
    const int te = ew * 32 + lane;
    if (kIsNorm && te < SB_N) {
      colc[te] = args.gamma[n0 + te];
      colc[64 + te] = (NORM == NORM_LAYER) ? args.beta[n0 + te] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = m0 + lane_base + rg + 4 * i;
      pre[i] = make_float4(0.f, 0.f, 0.f, 0.f); add[i] = make_float4(0.f, 0.f, 0.f, 0.f); prow[i] = 0.f; prow2[i] = 0.f;
      if (row < M) {
        pre[i] = *reinterpret_cast<const float4*>(args.z + static_cast<size_t>(row) * args.ldz + col);
        if (!kIsNorm && args.addend != nullptr) add[i] = *reinterpret_cast<const float4*>(args.addend + static_cast<size_t>(row) * args.ld_add + col);
        if (kIsNorm) { prow[i] = args.rstd[row]; if (NORM == NORM_LAYER) prow2[i] = args.mean[row]; }
      }
    }
    if (kIsNorm) epi_bar();   // colc complete
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // ---- part 1: accumulators -> staging ----
    {
      float v[32];
      tc_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + cbase, v);
      float4* dst = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    if (kIsNorm) {
      // pass A (coalesced): dy = d * swish'(y) and zh replace d in the staging; partial sums of dzh and dzh * zh
      float g[4], be[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { g[j] = colc[cbase + cl + j]; be[j] = colc[64 + cbase + cl + j]; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg + 4 * i;
        const float4 d4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, zz[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
        float dy[4], zh[4], a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zh[j] = (zz[j] - prow2[i]) * prow[i];
          dy[j] = dd[j] * fswish_grad(zh[j] * g[j] + be[j]);   // rows past the end: d = 0 (TMA zero fill)
          const float dzh = dy[j] * g[j];
          a1 += dzh; a2 += dzh * zh[j];
        }
        *reinterpret_cast<float4*>(stg + r * kStgLd + cl) = make_float4(dy[0], dy[1], dy[2], dy[3]);
        *reinterpret_cast<float4*>(stg2 + r * kStgLd + cl) = make_float4(zh[0], zh[1], zh[2], zh[3]);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
        if ((lane & 7) == 0) {
          part[(0 * 2 + half) * 128 + lane_base + r] = a1;
          part[(1 * 2 + half) * 128 + lane_base + r] = a2;
        }
      }
      epi_bar();
      if (half == 0) {
        const float u1 = part[0 * 128 + rrow] + part[1 * 128 + rrow];
        const float u2 = part[2 * 128 + rrow] + part[3 * 128 + rrow];
        for (int r = 0; r < ncta; ++r) {
          st_dsmem_f32(allstat + (rank * 2 + 0) * 128 + rrow, r, u1);
          st_dsmem_f32(allstat + (rank * 2 + 1) * 128 + rrow, r, u2);
        }
      }
    }
    tc_fence_before();
  }
  if (kIsNorm) { __syncwarp(); cluster_sync_all(); }   // every CTA's row statistics have landed

  if (warp >= 2) {
    float pg[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f}, pd[4] = {0.f, 0.f, 0.f, 0.f};
    if (kIsNorm) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r < ncta) { t1 += allstat[(r * 2 + 0) * 128 + rrow]; t2 += allstat[(r * 2 + 1) * 128 + rrow]; }
      }
      const float inv_n = 1.f / static_cast<float>(N);
      rowa[lane] = (NORM == NORM_LAYER) ? t1 * inv_n : 0.f;
      rowb[lane] = t2 * inv_n;
      __syncwarp();
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = colc[cbase + cl + j];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg + 4 * i, row = m0 + lane_base + r;
        const float4 y4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);    // dy (0 for rows past the end)
        const float4 h4 = *reinterpret_cast<const float4*>(stg2 + r * kStgLd + cl);   // zh
        const float dy[4] = {y4.x, y4.y, y4.z, y4.w}, zh[4] = {h4.x, h4.y, h4.z, h4.w};
        const float rstd = prow[i], S1 = rowa[r], S2 = rowb[r];
        float dz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dz[j] = rstd * (dy[j] * g[j] - S1 - zh[j] * S2);
          pg[j] += dy[j] * zh[j];
          pb[j] += dy[j];
          pd[j] += dz[j];
        }
        if (row < M) store_bf16x4(args.twin + static_cast<size_t>(row) * args.ld_twin + col, dz[0], dz[1], dz[2], dz[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg + 4 * i, row = m0 + lane_base + r;
        const float4 x4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
        const float dz0 = (x4.x + add[i].x) * fswish_grad(pre[i].x), dz1 = (x4.y + add[i].y) * fswish_grad(pre[i].y);
        const float dz2 = (x4.z + add[i].z) * fswish_grad(pre[i].z), dz3 = (x4.w + add[i].w) * fswish_grad(pre[i].w);
        if (row < M) {
          pd[0] += dz0; pd[1] += dz1; pd[2] += dz2; pd[3] += dz3;
          store_bf16x4(args.twin + static_cast<size_t>(row) * args.ld_twin + col, dz0, dz1, dz2, dz3);
        }
      }
    }
    const bool need_g = kIsNorm && args.g_gamma != nullptr;
    const bool need_b = NORM == NORM_LAYER && args.g_beta != nullptr;
    const bool need_d = args.g_bias != nullptr;
    if (need_g || need_b || need_d) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pg[j] += __shfl_xor_sync(0xffffffffu, pg[j], 8);  pg[j] += __shfl_xor_sync(0xffffffffu, pg[j], 16);
        pb[j] += __shfl_xor_sync(0xffffffffu, pb[j], 8);  pb[j] += __shfl_xor_sync(0xffffffffu, pb[j], 16);
        pd[j] += __shfl_xor_sync(0xffffffffu, pd[j], 8);  pd[j] += __shfl_xor_sync(0xffffffffu, pd[j], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          colpart[(0 * 8 + ew) * 32 + cl + j] = pg[j];
          colpart[(1 * 8 + ew) * 32 + cl + j] = pb[j];
          colpart[(2 * 8 + ew) * 32 + cl + j] = pd[j];
        }
      }
      epi_bar();
      const int te = ew * 32 + lane;
      if (te < SB_N) {
        const int c = n0 + te, h = te >> 5, cc = te & 31;
        float sg = 0.f, sb = 0.f, sd = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          sg += colpart[(0 * 8 + h * 4 + w) * 32 + cc];
          sb += colpart[(1 * 8 + h * 4 + w) * 32 + cc];
          sd += colpart[(2 * 8 + h * 4 + w) * 32 + cc];
        }
        if (need_g) atomicAdd(args.g_gamma + c, sg);
        if (need_b) atomicAdd(args.g_beta + c, sb);
        if (need_d) atomicAdd(args.g_bias + c, sd);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SB_N));
  }
}

template <int NORM>
cudaError_t launch_bwd(const CUtensorMap& ta, const CUtensorMap& tb, const BwdArgs& args, int M, int N, int K, cudaStream_t st) {
  static bool configured = false;
  auto kern = gemm_bf16_tc_bwd_kernel<NORM>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(N / SB_N, (M + SB_M - 1) / SB_M, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (NORM != NORM_NONE) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = N / SB_N; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, args, M, N, K);
}

}  // namespace

// dz = backward of [Linear -> norm -> swish] input gradient, see the file header.  dy: bf16 [M, K] (ld lda), W: bf16 [K, N] (ld ldw).
// cudaErrorNotSupported: shape outside the cluster scheme (caller runs the separate dgrad + row kernels).
cudaError_t launch_gemm_bf16_tc_bwd(int norm, int M, int N, int K, const void* dy, int lda, const void* W, int ldw, const float* z,
                                    const float* rstd, const float* mean, const float* gamma, const float* beta,
                                    const float* addend, void* dz_twin, int ld_twin, float* g_gamma, float* g_beta, float* g_bias,
                                    cudaStream_t st) {
  // measured at C2 (r01): 116.0 ms per update fused vs 111.6 ms with the separate dgrad + row kernels -- the 8-warp epilogue of a
  // one-CTA-per-SM tcgen05 kernel is slower at this elementwise work than a full-occupancy row kernel.  Opt-in.
  // REPPO_FUSE_BWD=1: every eligible layer; =2: only passes through a frozen network (no column-sum atomics in the epilogue)
  const char* env = getenv("REPPO_FUSE_BWD");   // read per launch (launches are captured into the graph once)
  const int fmode = env ? atoi(env) : 0;
  const bool off = fmode == 0 || (fmode == 2 && (g_gamma != nullptr || g_beta != nullptr || g_bias != nullptr));
  if (off || M <= 0 || N % SB_N != 0 || (norm != NORM_NONE && N / SB_N > 8) || (ld_twin & 3) || z == nullptr || dz_twin == nullptr)
    return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(z) & 15) || (addend != nullptr && (reinterpret_cast<uintptr_t>(addend) & 15)))
    return cudaErrorNotSupported;
  CUtensorMap ta, tb;
  if (!slab_encode_2d(&ta, dy, M, K, lda, SB_K, SB_M) || !slab_encode_2d(&tb, W, K, N, ldw, SB_N, SB_K)) return cudaErrorNotSupported;
  BwdArgs a{};
  a.z = z; a.ldz = N; a.rstd = rstd; a.mean = mean; a.gamma = gamma; a.beta = beta; a.addend = addend; a.ld_add = N;
  a.twin = static_cast<__nv_bfloat16*>(dz_twin); a.ld_twin = ld_twin; a.g_gamma = g_gamma; a.g_beta = g_beta; a.g_bias = g_bias;
  if (norm == NORM_NONE) return launch_bwd<NORM_NONE>(ta, tb, a, M, N, K, st);
  if (norm == NORM_RMS) return launch_bwd<NORM_RMS>(ta, tb, a, M, N, K, st);
  return launch_bwd<NORM_LAYER>(ta, tb, a, M, N, K, st);
}

}  // namespace reppo
