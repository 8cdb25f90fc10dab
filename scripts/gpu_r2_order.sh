#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
REPS=4 SPPS=1,16,128,1024 PROFILES=v2,v4_equirect timeout 900 python scripts/gpu_order_ab.py 2>&1 | tee gpurun_out/order_ab.jsonl
