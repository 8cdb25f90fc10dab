#!/usr/bin/env bash
# round 2 measurement visit (1 GPU): primitive sweep, bench (ours / fast / reference), ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 600 python scripts/primitive_sweep.py 2>gpurun_out/sweep.err | tee gpurun_out/primitive_sweep.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
timeout 600 python bench.py --math fast --no-cpu-baseline --steps 5 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_fast.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
timeout 600 python scripts/bench_configs.py --config 1 2>>gpurun_out/bench.err | tee gpurun_out/configs_n1.jsonl
timeout 600 python scripts/bench_configs.py --config 3 2>>gpurun_out/bench.err | tee -a gpurun_out/configs_n1.jsonl
timeout 600 python scripts/bench_configs.py --config 4 2>>gpurun_out/bench.err | tee -a gpurun_out/configs_n1.jsonl
timeout 600 python scripts/bench_configs.py --config 5 2>>gpurun_out/bench.err | tee -a gpurun_out/configs_n1.jsonl
# ncu: launch list of the bench command, then full captures (each after its plain run exited 0)
ARGS="--spp 128 --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
PROFS="v2 v4_equirect v4_cubemap simt v3redo" SCHEDS="lane" bash scripts/gpu_prof_sched.sh
