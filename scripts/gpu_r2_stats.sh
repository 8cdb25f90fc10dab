#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python scripts/compare_asis.py 2>gpurun_out/asis.err | tee gpurun_out/compare_asis.jsonl
tail -3 gpurun_out/asis.err
timeout 1200 python scripts/fast_math_rmse.py 2>gpurun_out/fast.err | tee gpurun_out/fast_math_rmse.jsonl
tail -3 gpurun_out/fast.err
