#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for prof in v2 simt v4_equirect v4_cubemap v3redo; do python scripts/prof_any.py $prof 1024 3 2>&1 | tail -1; done | tee gpurun_out/sqrt_ab.log
