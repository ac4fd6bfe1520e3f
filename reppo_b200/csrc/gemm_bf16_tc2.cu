// Persistent CTA-pair tcgen05 GEMM for the LARGE Linear layers (wide-net / 256k-sample sweep).
//
//   D[M,N] (fp32) = A · B, bf16 operands, same three operand arrangements as gemm_bf16_tc.cu (forward, dgrad, wgrad of
//   y = x W^T + b; src/networks/jax_models.py:35-49, torch_models.py:71-79).
//
// Why a second kernel: at M = 16k-32k rows and H = 1024-2048 the one-tile-per-CTA kernel is bound by operand delivery
// from L2 (a 128x128 tile pulls 64 FLOP per L2 byte; the L2 -> SM fabric sustains ~42 B/clk/SM) and its epilogue is
// serialised behind its main loop.  Here:
//   * tcgen05.mma.cta_group::2 — the two CTAs of a cluster (one TPC) compute ONE 256 x BN tile: each CTA stages its own
//     128 rows of A and HALF of the B tile, the leader CTA issues UMMA 256 x BN x 16 for both, each CTA's 128 x BN fp32
//     accumulator lands in its own TMEM.  A 256x256 tile pulls 131 FLOP per L2 byte.
//   * persistent: one pair per TPC walks the tile list (N fastest, so the pairs running together share A rows in L2);
//   * TWO accumulator stages in TMEM (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1
//     (tmem_full / tmem_empty mbarriers; the peer CTA's epilogue arrives on the leader's tmem_empty remotely);
//   * TMA (cp.async.bulk.tensor.2d.cta_group::2) from both CTAs signals the LEADER's full barrier; tcgen05.commit
//     multicast releases the smem stage in both CTAs.
// Warp roles (320 threads per CTA): warp 0 = TMA producer (+ TMEM alloc), warp 1 = MMA issuer (leader CTA only),
// warps 2..9 = epilogue (TMEM -> registers -> smem transpose -> coalesced global stores, 32 x 32 chunks).
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

namespace {

constexpr int T2_M = 128;         // rows per CTA (256 per pair)
constexpr int T2_K = 64;          // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int T2_UMMA_K = 16;
constexpr int T2_THREADS = 320;
constexpr int T2_STG_LD = 33;     // staging row stride (floats): scalar transpose, conflict-free both ways

template <int BN, int STAGES>
struct T2Smem {
  static constexpr int kABytes = T2_M * T2_K * 2;            // 16 KB
  static constexpr int kBBytes = (BN / 2) * T2_K * 2;        // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = STAGES;
  static constexpr int kStgOffset = kStages * kStageBytes;
  static constexpr int kStgBytes = 8 * 32 * T2_STG_LD * 4;
  static constexpr int kBarOffset = kStgOffset + kStgBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
};

REPPO_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Both CTAs of the pair signal the barrier of the EVEN CTA (bit 24 of a shared::cluster address selects the peer)
REPPO_DEVINL void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
REPPO_DEVINL void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (when every MMA issued so far has retired) on the barrier at this smem offset in BOTH CTAs of the pair
REPPO_DEVINL void tc_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
REPPO_DEVINL void mbar_arrive_cluster(uint64_t* local_bar, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_bar)), "r"(cta_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// bf16 x bf16 -> f32, UMMA M = 256 (two CTAs), N = n
__host__ __device__ constexpr uint32_t make_idesc_pair(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

struct T2Tile { int m0, n0, kb_begin, kb_end; };

REPPO_DEVINL T2Tile t2_decode(int t, int n_tiles, int m_tiles, int BN, int per, int total_kb) {
  T2Tile r;
  const int nt = t % n_tiles;
  const int rest = t / n_tiles;
  const int mt = rest % m_tiles, z = rest / m_tiles;
  r.m0 = mt * (2 * T2_M); r.n0 = nt * BN;
  r.kb_begin = z * per; r.kb_end = min(total_kb, r.kb_begin + per);
  return r;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     float* __restrict__ C, int ldc, const float* __restrict__ bias, int M, int N, int K, int k_blocks_per_split,
                     int splits, int mode, __nv_bfloat16* __restrict__ act_out, int act_ld) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  using S = T2Smem<BN, STAGES>;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2], used in the leader CTA only
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (K + T2_K - 1) / T2_K;
  const int m_tiles = (M + 2 * T2_M - 1) / (2 * T2_M), n_tiles = (N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles * splits;

  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
    // TMEM: two accumulator stages of BN fp32 columns, allocated in both CTAs of the pair by the same-numbered warp
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 16); }   // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of BOTH CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  REPPO_PDL_PROLOGUE();

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own half of the B tile =====
    uint32_t kbi = 0;
    for (int t = pair; t < total_tiles; t += npairs) {
      const T2Tile tl = t2_decode(t, n_tiles, m_tiles, BN, k_blocks_per_split, total_kb);
      const int am0 = tl.m0 + static_cast<int>(rank) * T2_M;
      const int bn0 = tl.n0 + static_cast<int>(rank) * (BN / 2);
      for (int kb = tl.kb_begin; kb < tl.kb_end; ++kb, ++kbi) {
        const int s = kbi % STAGES;
        const uint32_t ph = (kbi / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);   // bytes of both CTAs land on the leader's barrier
        const int k0 = kb * T2_K;
        if (!A_MN) {
          tma_load_2d_pair(&tmap_a, &full_bar[s], sa, k0, am0);
        } else {
#pragma unroll
          for (int a = 0; a < T2_M / 64; ++a) tma_load_2d_pair(&tmap_a, &full_bar[s], sa + a * (T2_K * 128), am0 + a * 64, k0);
        }
        if (!B_MN) {
          tma_load_2d_pair(&tmap_b, &full_bar[s], sb, k0, bn0);
        } else {
#pragma unroll
          for (int b = 0; b < (BN / 2) / 64; ++b) tma_load_2d_pair(&tmap_b, &full_bar[s], sb + b * (T2_K * 128), bn0 + b * 64, k0);
        }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (leader CTA): UMMA 256 x BN x 16 over both CTAs' shared memory =====
    constexpr uint32_t idesc = make_idesc_pair(BN, A_MN, B_MN);
    uint32_t kbi = 0, it = 0;
    for (int t = pair; t < total_tiles; t += npairs, ++it) {
      const T2Tile tl = t2_decode(t, n_tiles, m_tiles, BN, k_blocks_per_split, total_kb);
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[as], aph ^ 1);      // both CTAs' epilogues have drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = tl.kb_begin; kb < tl.kb_end; ++kb, ++kbi) {
        const int s = kbi % STAGES;
        const uint32_t ph = (kbi / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < T2_K / T2_UMMA_K; ++k) {
          const uint64_t da = A_MN ? make_smem_desc(sa + k * (T2_UMMA_K * 128), T2_K * 128, 1024)
                                   : make_smem_desc(sa + k * (T2_UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb + k * (T2_UMMA_K * 128), T2_K * 128, 1024)
                                   : make_smem_desc(sb + k * (T2_UMMA_K * 2), 16, 1024);
          tc_mma_bf16_pair(tmem_d, da, db, idesc, (kb > tl.kb_begin || k > 0) ? 1u : 0u);
        }
        tc_commit_pair(&empty_bar[s], 0x3);          // frees the stage in both CTAs when these MMAs retire
      }
      tc_commit_pair(&tmem_full_bar[as], 0x3);       // accumulator complete: wake both CTAs' epilogues
    }
  } else if (warp >= 2) {
    // ===== epilogue (both CTAs): this CTA's 128 rows x BN columns, 32 x 32 chunks =====
    const int ew = warp - 2;
    const int lane_base = (warp & 3) * 32;          // TMEM lane quarter this warp may access
    constexpr int kColsW = BN / 2;                   // the two warps of a lane quarter split the columns
    const int cbase = (ew >> 2) * kColsW;
    float* stg = reinterpret_cast<float*>(smem + S::kStgOffset) + ew * 32 * T2_STG_LD;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (mode != GEMM_ATOMIC) &&
                        (act_out == nullptr || (act_ld & 3) == 0);
    const int cl = (lane & 7) * 4;
    uint32_t it = 0;
    for (int t = pair; t < total_tiles; t += npairs, ++it) {
      const T2Tile tl = t2_decode(t, n_tiles, m_tiles, BN, k_blocks_per_split, total_kb);
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const bool add_bias = (bias != nullptr) && (tl.kb_begin == 0);
      const int row_base = tl.m0 + static_cast<int>(rank) * T2_M + lane_base;
      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < kColsW; c0 += 32) {
        const int colv = tl.n0 + cbase + c0 + cl;     // this lane's 4 columns in the coalesced store layout
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add_bias && vec_ok) {
          if (colv + 0 < N) bv.x = __ldg(bias + colv + 0);
          if (colv + 1 < N) bv.y = __ldg(bias + colv + 1);
          if (colv + 2 < N) bv.z = __ldg(bias + colv + 2);
          if (colv + 3 < N) bv.w = __ldg(bias + colv + 3);
        }
        float v[32];
        tc_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + as * BN + cbase + c0, v);
        if (c0 + 32 >= kColsW) {
          // last read of this accumulator stage: hand it back to the MMA issuer (leader CTA's barrier)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * T2_STG_LD + j] = v[j];
        __syncwarp();
        if (vec_ok) {
#pragma unroll 4
          for (int r = lane >> 3; r < 32; r += 4) {
            const int row = row_base + r;
            if (row >= M || colv >= N) continue;
            const float* sp = stg + r * T2_STG_LD + cl;
            float4 x = make_float4(sp[0], sp[1], sp[2], sp[3]);
            x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
            float* cp = C + static_cast<size_t>(row) * ldc + colv;
            if (colv + 3 < N) {
              if (mode == GEMM_ACCUM) {
                const float4 o = *reinterpret_cast<const float4*>(cp);
                x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
              }
              *reinterpret_cast<float4*>(cp) = x;
              if (act_out != nullptr)   // swish(y) as the bf16 operand of the next GEMM
                store_bf16x4(act_out + static_cast<size_t>(row) * act_ld + colv, fswish(x.x), fswish(x.y), fswish(x.z), fswish(x.w));
            } else {
              const float xs[4] = {x.x, x.y, x.z, x.w};
              for (int j = 0; j < 4 && colv + j < N; ++j) {
                const float o = (mode == GEMM_ACCUM) ? cp[j] + xs[j] : xs[j];
                cp[j] = o;
                if (act_out != nullptr) act_out[static_cast<size_t>(row) * act_ld + colv + j] = __float2bfloat16_rn(fswish(o));
              }
            }
          }
        } else {
          // ragged leading dimension or split-K atomics: lanes sweep 32 consecutive columns of one row
          const int col = tl.n0 + cbase + c0 + lane;
          const float bs = (add_bias && col < N) ? __ldg(bias + col) : 0.f;
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const int row = row_base + r;
            if (row >= M) break;
            if (col < N) {
              float x = stg[r * T2_STG_LD + lane] + bs;
              float* cp = C + static_cast<size_t>(row) * ldc + col;
              if (mode == GEMM_STORE) *cp = x;
              else if (mode == GEMM_ACCUM) { x += *cp; *cp = x; }
              else atomicAdd(cp, x);
              if (act_out != nullptr && mode != GEMM_ATOMIC) act_out[static_cast<size_t>(row) * act_ld + col] = __float2bfloat16_rn(fswish(x));
            }
          }
        }
        __syncwarp();
      }
    }
  }
  // both CTAs stay until every MMA / remote arrive of the pair has landed, then each frees its TMEM
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
cudaError_t launch_t2(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int ldc, const float* bias, int M, int N, int K,
                      int splits, int mode, __nv_bfloat16* act_out, int act_ld, int max_pairs, cudaStream_t st) {
  using S = T2Smem<BN, STAGES>;
  static_assert(S::kTotal <= 227 * 1024, "shared memory");
  static unsigned long long configured = 0ull;   // one bit per device: the attribute is per device
  auto kern = gemm_bf16_tc2_kernel<BN, STAGES, A_MN, B_MN>;
  int dev__ = 0;
  cudaGetDevice(&dev__);
  if (!((configured >> (dev__ & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    configured |= 1ull << (dev__ & 63);
  }
  const int total_kb = (K + T2_K - 1) / T2_K;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  const int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;
  if (splits > 1) mode = GEMM_ATOMIC;
  const int tiles = ((M + 2 * T2_M - 1) / (2 * T2_M)) * ((N + BN - 1) / BN) * splits;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, C, ldc, bias, M, N, K, per, splits, mode, act_out, act_ld);
}

int t2_max_pairs() {
  static int pairs = 0;
  if (pairs == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2)
      sms = 148;
    pairs = sms / 2;
  }
  return pairs;
}

}  // namespace

// Large-shape path of launch_gemm_bf16_tc.  cudaErrorNotSupported: the shape does not fill the machine with pair tiles
// (the caller then runs the one-tile-per-CTA kernel).  REPPO_TC2=0 disables, REPPO_TC2_MIN=<pair tiles> moves the threshold,
// REPPO_TC2_BN=128|256 forces the tile width.
// tile width (256 / 128) and split-K factor of the pair kernel for this shape; bn == 0: not eligible
struct T2Plan { int bn, splits; };
static T2Plan t2_plan(bool a_mn, bool b_mn, int M, int N, int K, int mode, int split_k) {
  const bool enabled = knobs().tc2;
  const int min_tiles = knobs().tc2_min, force_bn = knobs().tc2_bn;
  T2Plan p{0, 1};
  if (!enabled || (a_mn && !b_mn)) return p;
  if (K < 8 || N < 128 || M < 256) return p;   // thin-K first layers too: at these row counts they are epilogue (store) bound
  const int max_pairs = t2_max_pairs();
  const int total_kb = (K + T2_K - 1) / T2_K;
  const int mt = (M + 2 * T2_M - 1) / (2 * T2_M);
  const int t256 = mt * ((N + 255) / 256), t128 = mt * ((N + 127) / 128);
  // split-K (wgrad, atomics into the zeroed gradient arena) is re-chosen for pair tiles: fill the pairs once, >= 8 k-blocks each
  auto pick_split = [&](int tiles) {
    if (mode != GEMM_ATOMIC && split_k <= 1) return 1;
    int s = max_pairs / tiles;
    if (s > total_kb / 8) s = total_kb / 8;
    return s < 1 ? 1 : s;
  };
  // padded columns are computed and thrown away: no more than a quarter of the tile row
  auto waste_ok = [&](int w) { const int padded = ((N + w - 1) / w) * w; return 4 * (padded - N) <= padded; };
  if (force_bn == 128 || force_bn == 256) p.bn = force_bn;
  else if (t256 * pick_split(t256) >= min_tiles && waste_ok(256)) p.bn = 256;
  else if (t128 * pick_split(t128) >= min_tiles && waste_ok(128)) p.bn = 128;
  if (p.bn != 0) p.splits = pick_split(p.bn == 256 ? t256 : t128);
  return p;
}

bool gemm_bf16_tc2_eligible(bool a_mn, bool b_mn, int M, int N, int K, int mode, int split_k) {
  return t2_plan(a_mn, b_mn, M, N, K, mode, split_k).bn != 0;
}

cudaError_t launch_gemm_bf16_tc2(bool a_mn, bool b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                 float* C, int ldc, const float* bias, int mode, int split_k, void* act_out_v, int act_ld,
                                 cudaStream_t st) {
  const T2Plan plan = t2_plan(a_mn, b_mn, M, N, K, mode, split_k);
  if (plan.bn == 0) return cudaErrorNotSupported;
  const int bn = plan.bn, splits = plan.splits, max_pairs = t2_max_pairs();
  CUtensorMap ta, tb;
  const bool ok_a = a_mn ? tmap_encode_2d(&ta, A, K, M, lda, 64, T2_K) : tmap_encode_2d(&ta, A, M, K, lda, T2_K, T2_M);
  const bool ok_b = b_mn ? tmap_encode_2d(&tb, B, K, N, ldb, 64, T2_K) : tmap_encode_2d(&tb, B, N, K, ldb, T2_K, bn / 2);
  if (!ok_a || !ok_b) return cudaErrorNotSupported;
  __nv_bfloat16* act_out = static_cast<__nv_bfloat16*>(act_out_v);
  const int stages_env = knobs().tc2_stages;   // experiments: 3 / 4 / 5
#define REPPO_T2_ARGS ta, tb, C, ldc, bias, M, N, K, splits, mode, act_out, act_ld, max_pairs, st
#define REPPO_T2_DISPATCH(AMN, BMN)                                                    \
  if (bn == 256) {                                                                     \
    if (stages_env == 3) return launch_t2<256, 3, AMN, BMN>(REPPO_T2_ARGS);            \
    if (stages_env == 4) return launch_t2<256, 4, AMN, BMN>(REPPO_T2_ARGS);            \
    if (stages_env == 5) return launch_t2<256, 5, AMN, BMN>(REPPO_T2_ARGS);            \
    return launch_t2<256, 6, AMN, BMN>(REPPO_T2_ARGS);                                 \
  }                                                                                    \
  if (stages_env == 4) return launch_t2<128, 4, AMN, BMN>(REPPO_T2_ARGS);              \
  return launch_t2<128, 8, AMN, BMN>(REPPO_T2_ARGS)
  if (!a_mn && !b_mn) { REPPO_T2_DISPATCH(false, false); }
  if (!a_mn && b_mn) { REPPO_T2_DISPATCH(false, true); }
  REPPO_T2_DISPATCH(true, true);
#undef REPPO_T2_DISPATCH
#undef REPPO_T2_ARGS
}

}  // namespace reppo
