#!/usr/bin/env bash
# scripts/full_job_parity.py for the remaining renderers / samplers at bench-size images: whole buffers against the reference build
set -u
mkdir -p gpurun_out
out=gpurun_out/full_job_parity_all.jsonl
rm -f $out
run() { timeout 1500 python scripts/full_job_parity.py "$@" 2>>gpurun_out/full_job_parity_all.err | tee -a $out; }
run --profile simt --width 3840 --height 2160 --bounces 4 --spp 32
run --profile v3redo --spp 128
run --profile v3redo0 --spp 128
run --profile v4_cubemap --spp 1024 --skip-oracle
run --profile v4_bilinear --spp 1024 --skip-oracle
run --profile v4_cubemap_bilinear --spp 1024 --skip-oracle
tail -3 gpurun_out/full_job_parity_all.err
