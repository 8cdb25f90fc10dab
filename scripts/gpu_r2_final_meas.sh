#!/usr/bin/env bash
# end-of-round measurement visit (1 GPU)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python scripts/gpu_sched_ab.py 2>&1 | tee gpurun_out/sched_ab.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
timeout 600 python bench.py --math fast --no-cpu-baseline --steps 5 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_fast.json
rm -f gpurun_out/configs_n1.jsonl
for cfg in 1 3 4 5; do timeout 600 python scripts/bench_configs.py --config $cfg 2>>gpurun_out/bench.err | tee -a gpurun_out/configs_n1.jsonl; done
timeout 600 python scripts/primitive_sweep.py 2>gpurun_out/sweep.err | tee gpurun_out/primitive_sweep.jsonl
ARGS="--spp 128 --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
PROFS="v2 v4_equirect v4_cubemap" SCHEDS="lane" bash scripts/gpu_prof_sched.sh
