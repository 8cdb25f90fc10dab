#!/usr/bin/env bash
# sweep of the resident-CTAs-per-SM launch bound (registers per thread), per kernel family, on the round-2 kernels
set -u
mkdir -p gpurun_out
: > gpurun_out/tune.log
for mb in 2 3 4 5; do
  B200PT_MIN_BLOCKS_CORNELL=$mb B200PT_MIN_BLOCKS_V4=$mb python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
  for prof in v2 v4_equirect v4_cubemap simt v3redo; do
    echo "MIN_BLOCKS=$mb $(python scripts/prof_any.py $prof 256 3 2>&1 | tail -1)" | tee -a gpurun_out/tune.log
  done
done
python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
