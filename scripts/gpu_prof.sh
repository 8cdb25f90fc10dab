#!/usr/bin/env bash
# ncu launch list + full capture of the megakernel on the bench workload (short spp)
set -u
mkdir -p gpurun_out
ARGS="--spp 128 --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -o gpurun_out/prof -f \
    python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log; tail -3 gpurun_out/ncu_full.log
