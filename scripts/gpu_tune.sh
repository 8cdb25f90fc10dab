#!/usr/bin/env bash
# sweep of the resident-CTAs-per-SM launch bound, per kernel family
for mb in 2 3 4 5; do
  B200PT_MIN_BLOCKS_CORNELL=$mb B200PT_MIN_BLOCKS_V4=$mb python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
  echo "== MIN_BLOCKS=$mb"; python scripts/gpu_profiles_quick.py 2>&1 | grep "1080p"
done
python -m cpuperformanceraytracer_b200.build --force > /dev/null 2>&1
