#!/usr/bin/env bash
# the headline job through the C group API on N GPUs (tile sharding must reproduce the reference build's SHA-256), then the bench line
set -u
mkdir -p gpurun_out
N=${N:-4}
nvidia-smi -L | wc -l
timeout 900 python scripts/full_job_multi_gpu.py --gpus $N 2>gpurun_out/full_job_multi_gpu.err | tee gpurun_out/full_job_multi_gpu_n$N.json
tail -3 gpurun_out/full_job_multi_gpu.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400
