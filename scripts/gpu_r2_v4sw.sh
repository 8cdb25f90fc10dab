#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_switch_cost.py 2>&1 | tee gpurun_out/switch_cost.jsonl
