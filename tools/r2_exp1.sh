#!/bin/bash
# round-2 experiment 1: where is the C2 update today, and how does it react to the existing tile-shape knobs
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
bash tools/run_knobs.sh A=1 \
  REPPO_TC_WIDE_MIN=0 \
  REPPO_TC_WIDE_MIN=0,REPPO_NORM_BN128=1 \
  REPPO_NORM_BN128=1 \
  REPPO_TC2_MIN=1,REPPO_TC2_BN=128,REPPO_TC2_STAGES=4 \
  REPPO_NO_FUSE_NORM=1 \
  REPPO_PRIO=0 \
  REPPO_NO_PIPELINE=1 \
  REPPO_SERIAL=1 \
  A=1 2>&1 | tee gpurun_out/r2_exp1.txt
