#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in ${NS:-4 8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-260
  tail -2 gpurun_out/bench_n$N.err | cut -c1-300
done
