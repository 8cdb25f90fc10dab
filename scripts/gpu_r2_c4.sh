#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "present" 2>&1 | tail -5
timeout 600 python scripts/bench_configs.py --config 4 2>gpurun_out/c4.err | tee gpurun_out/config4.jsonl
tail -3 gpurun_out/c4.err
