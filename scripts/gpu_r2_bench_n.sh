#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -4 gpurun_out/bench_n$N.err
