#!/usr/bin/env bash
# sweep of CTA size x launch bound per kernel family (pt_render_kernel)
set -u
mkdir -p gpurun_out
: > gpurun_out/tune3.log
run() { # T C V4 V3
  B200PT_THREADS_CORNELL=$1 B200PT_THREADS_V4=$1 B200PT_THREADS_V3REDO=$1 B200PT_MIN_BLOCKS_CORNELL=$2 B200PT_MIN_BLOCKS_V4=$3 B200PT_MIN_BLOCKS_V3REDO=$4 \
    python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1 || { echo "build failed $*" | tee -a gpurun_out/tune3.log; return; }
  for prof in v2 simt v4_equirect v4_cubemap v3redo; do
    echo "T=$1 MB(cornell,v4,v3)=$2,$3,$4 $(python scripts/prof_any.py $prof 256 3 2>&1 | tail -1)" | tee -a gpurun_out/tune3.log
  done
}
run 128 6 5 7
run 128 7 6 8
run 128 8 7 8
run 192 4 3 5
run 192 5 4 6
run 256 3 2 4
python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
