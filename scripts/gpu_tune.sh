#!/usr/bin/env bash
for mb in 1 2 3 4 5; do
  B200PT_MIN_BLOCKS=$mb python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
  echo "== MIN_BLOCKS=$mb"; python scripts/gpu_quick.py 2>&1 | grep "1080p" | grep "nframes=256"
done
