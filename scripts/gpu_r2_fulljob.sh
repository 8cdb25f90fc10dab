#!/usr/bin/env bash
# the headline job at full size against the reference's own code and the oracle (scripts/full_job_parity.py): a quick 4-spp pass
# first (proves the script), then the 1024-spp job
set -u
mkdir -p gpurun_out
nproc
timeout 600 python scripts/full_job_parity.py --spp 4 2>gpurun_out/full_job_parity.err | tee gpurun_out/full_job_parity_spp4.json || exit 1
grep -q '"gpu_equals_reference_bit_for_bit": true' gpurun_out/full_job_parity_spp4.json || { tail -5 gpurun_out/full_job_parity.err; exit 1; }
timeout 2400 python scripts/full_job_parity.py --spp ${SPP:-1024} 2>>gpurun_out/full_job_parity.err | tee gpurun_out/full_job_parity.json
tail -3 gpurun_out/full_job_parity.err
