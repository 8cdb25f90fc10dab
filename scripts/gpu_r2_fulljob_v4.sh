#!/usr/bin/env bash
# the reference's default renderer (P_v4, equirect env, random-jitter sampler) at the headline size and sample count against the
# reference's own code on all host cores (scripts/full_job_parity.py --profile v4)
set -u
mkdir -p gpurun_out
timeout 600 python scripts/full_job_parity.py --profile v4 --spp 4 2>gpurun_out/full_job_parity_v4.err | tee gpurun_out/full_job_parity_v4_spp4.json || exit 1
grep -q '"gpu_equals_reference_bit_for_bit": true' gpurun_out/full_job_parity_v4_spp4.json || { tail -5 gpurun_out/full_job_parity_v4.err; exit 1; }
timeout 2400 python scripts/full_job_parity.py --profile v4 --spp ${SPP:-1024} --skip-oracle 2>>gpurun_out/full_job_parity_v4.err | tee gpurun_out/full_job_parity_v4.json
tail -3 gpurun_out/full_job_parity_v4.err
