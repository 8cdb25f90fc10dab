#!/usr/bin/env bash
# BASELINE configs[4] geometry (8192x8192, 16 bounces) at a bounded sample count against the reference's own code and the oracle
set -u
mkdir -p gpurun_out
timeout 1800 python scripts/full_job_parity.py --width 8192 --height 8192 --bounces 16 --spp ${SPP:-4} 2>gpurun_out/full_job_parity_c5.err | tee gpurun_out/full_job_parity_c5.json
tail -3 gpurun_out/full_job_parity_c5.err
