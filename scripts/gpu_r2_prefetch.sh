#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sorted.py -m gpu -q -x 2>&1 | tail -4
REPS=4 SPPS=1,4,16,128,1024 PROFILES=v2,v4_equirect,v4_cubemap timeout 900 python scripts/gpu_order_ab.py 2>&1 | grep scene-first | tee gpurun_out/prefetch_ab.jsonl
