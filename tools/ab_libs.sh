#!/bin/bash
# A/B two builds of libreppo_b200.so on the SAME GPU box (box-to-box variance is ~1 %, code-generation effects ~3 %):
#   1. on the build host: build variant X, `cp reppo_b200/libreppo_b200.so reppo_b200/lib_X.so`; same for Y (the *.so files
#      are git-ignored but travel with the gpurun snapshot);
#   2. gpurun -- 'bash tools/ab_libs.sh X Y X Y'
# Prints samples/s, ms per update and launches per update of the C2 bench for every variant in turn.
for v in "$@"; do
  cp "reppo_b200/lib_$v.so" reppo_b200/libreppo_b200.so || exit 1
  echo "== $v"
  bash tools/run_knobs.sh A=1 | tail -1
done
