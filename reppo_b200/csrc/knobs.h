// Experiment knobs of libreppo_b200.so, read from the environment ONCE per process.
//
// None of these selects a different algorithm: they switch between kernel variants / schedules that are all
// parity-tested, so that two variants can be A/B-timed on the same GPU box (tools/run_knobs.sh).  Defaults are the
// measured winners; DESIGN.md section 7 records the numbers behind each default.
#pragma once
#include <stdlib.h>

namespace reppo {

struct Knobs {
  // orchestration (api.cu)
  bool serial;           // REPPO_SERIAL=1        : no side streams, every kernel of a step on one stream
  bool no_pipeline;      // REPPO_NO_PIPELINE=1   : no two-chain software pipeline over steps
  bool no_fuse_norm;     // REPPO_NO_FUSE_NORM=1  : Linear and norm+swish as two kernels
  bool no_fuse_act;      // REPPO_NO_FUSE_ACT=1   : swish(features) by a row kernel instead of the GEMM epilogue
  bool no_sibling;       // REPPO_NO_SIBLING=1    : critic head / pred head first layers as two GEMMs instead of one
  bool no_pregather;     // REPPO_NO_PREGATHER=1  : gather the minibatch rows inside every step instead of once per epoch
  int prio;              // REPPO_PRIO=0|1        : graph node priorities for the actor chain (-1 = auto: small steps, 1 GPU)
  int wgrad_split;       // REPPO_WGRAD_SPLIT=n   : split-K factor of the weight-gradient GEMMs (0 = auto)
  int nccl_own;          // REPPO_ALLREDUCE=nccl|mc : gradient all-reduce through NCCL or the NVSwitch-multicast kernel (-1 auto)
  // GEMM selection (gemm_bf16_tc.cu / gemm_bf16_tc2.cu)
  int tc_wide_min;       // REPPO_TC_WIDE_MIN=n   : 128-wide tiles in the one-CTA kernel when >= n of them exist
  bool norm_bn128;       // REPPO_NORM_BN128=1    : 128-wide tiles in the one-CTA fused-norm kernel
  bool fuse_norm_large;  // REPPO_FUSE_NORM_LARGE=1 : keep the fused-norm kernel at the wide-net shapes
  bool tc2;              // REPPO_TC2=0           : disable the persistent CTA-pair kernel
  int tc2_min, tc2_bn, tc2_stages;
  // row kernels / optimizer
  bool no_vec, no_wide, no_pdl;
  int bwd_rpb, wide_bps, adam_blocks;
  bool adam_fused;       // REPPO_ADAM_FUSED=1 : global norm + clip + Adam in ONE launch with a grid barrier (default: two launches)
  int adam_threads;      // REPPO_ADAM_THREADS=256|512 : block size of clip_adam_kernel (default 512: 96.2 vs 101.7 ms per C2 update)
  const char* nccl_lib;  // REPPO_NCCL_LIB
};

inline int knob_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v != nullptr && *v != '\0') ? atoi(v) : dflt;
}
inline bool knob_set(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && *v != '\0' && !(v[0] == '0' && v[1] == '\0');
}

inline const Knobs& knobs() {
  static const Knobs k = [] {
    Knobs x{};
    x.serial = knob_set("REPPO_SERIAL");
    x.no_pipeline = knob_set("REPPO_NO_PIPELINE");
    x.no_fuse_norm = knob_set("REPPO_NO_FUSE_NORM");
    x.no_fuse_act = knob_set("REPPO_NO_FUSE_ACT");
    x.no_sibling = knob_set("REPPO_NO_SIBLING");
    x.no_pregather = knob_set("REPPO_NO_PREGATHER");
    x.prio = knob_int("REPPO_PRIO", -1);
    x.wgrad_split = knob_int("REPPO_WGRAD_SPLIT", 0);
    {
      const char* v = getenv("REPPO_ALLREDUCE");
      x.nccl_own = (v == nullptr) ? -1 : ((v[0] == 'm' || v[0] == 'M') ? 1 : 0);
    }
    x.tc_wide_min = knob_int("REPPO_TC_WIDE_MIN", 120);
    x.norm_bn128 = knob_set("REPPO_NORM_BN128");
    x.fuse_norm_large = knob_set("REPPO_FUSE_NORM_LARGE");
    x.tc2 = knob_int("REPPO_TC2", 1) != 0;
    x.tc2_min = knob_int("REPPO_TC2_MIN", 56);
    x.tc2_bn = knob_int("REPPO_TC2_BN", 0);
    x.tc2_stages = knob_int("REPPO_TC2_STAGES", 0);
    x.no_vec = knob_set("REPPO_NO_VEC");
    x.no_wide = knob_set("REPPO_NO_WIDE");
    x.no_pdl = knob_set("REPPO_NO_PDL");
    x.bwd_rpb = knob_int("REPPO_BWD_RPB", 0);
    x.wide_bps = knob_int("REPPO_WIDE_BPS", 2);
    x.adam_blocks = knob_int("REPPO_ADAM_BLOCKS", 0);
    x.adam_fused = knob_set("REPPO_ADAM_FUSED");
    x.adam_threads = knob_int("REPPO_ADAM_THREADS", 512);
    x.nccl_lib = getenv("REPPO_NCCL_LIB");
    return x;
  }();
  return k;
}

}  // namespace reppo
