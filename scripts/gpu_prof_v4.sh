#!/usr/bin/env bash
# ncu full capture of the P_v4 megakernel (equirect random sampler), after a plain run exited 0
set -u
mkdir -p gpurun_out
python scripts/prof_v4.py ${V4ENV:-equirect} 128 > gpurun_out/v4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -o gpurun_out/prof_v4 -f \
    python scripts/prof_v4.py ${V4ENV:-equirect} 128 > gpurun_out/v4_ncu_full.log 2>&1
tail -3 gpurun_out/v4_plain.log; tail -3 gpurun_out/v4_ncu_full.log
