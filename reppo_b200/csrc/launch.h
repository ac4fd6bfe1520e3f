// Kernel launch helper + programmatic dependent launch (PDL).
//
// Every kernel of the update step is launched with cudaLaunchAttributeProgrammaticStreamSerialization, and every
// kernel begins with `griddepcontrol.wait` (before its first global-memory access) followed by
// `griddepcontrol.launch_dependents`.  Inside the captured CUDA graph consecutive kernels of one stream are then
// connected by programmatic edges: the next kernel's CTAs are scheduled, read their parameters and run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch) while the current kernel is still running, and only
// block at the wait.  The update is a chain of ~45 dependent, few-microsecond kernels per minibatch step, so the
// launch latency hidden this way is a first-order term.  `griddepcontrol.wait` returns only when the prerequisite
// grids have COMPLETED and their memory is visible, so correctness never depends on the early launch; because every
// kernel waits unconditionally, completion is transitive along the chain.  REPPO_NO_PDL=1 disables the attribute.
#pragma once
#include <cuda_runtime.h>

#include "knobs.h"

#define REPPO_PDL_PROLOGUE()                                   \
  do {                                                         \
    asm volatile("griddepcontrol.wait;" ::: "memory");         \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

namespace reppo {

inline bool pdl_enabled() { return !knobs().no_pdl; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace reppo
