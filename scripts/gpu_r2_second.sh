#!/usr/bin/env bash
# round 2, second visit (1 GPU): whole -m gpu suite (no -x), scheduler A/B
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 900 python scripts/gpu_sched_ab.py 2>&1 | tee gpurun_out/sched_ab.jsonl
