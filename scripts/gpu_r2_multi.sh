#!/usr/bin/env bash
# multi-GPU visit: N = number of GPUs of the box
set -u
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py tests/test_gpu_host_mirror.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
# the C++ CLI on N GPUs: timing of the group paths on the headline geometry (1080p, 256 frames)
for mode in "--shard spp --combine nccl" "--shard spp --combine peer" "--shard tiles"; do
  timeout 300 ./cpuperformanceraytracer_b200/render_offline --variant v2 --bounces 8 --width 1920 --height 1080 --tiles-x 10 --tiles-y 15 --frames 256 --gpus $N $mode --out gpurun_out/cli.bmp 2>&1 | tail -2
done
timeout 300 ./cpuperformanceraytracer_b200/render_offline --variant v2 --bounces 8 --width 1920 --height 1080 --tiles-x 10 --tiles-y 15 --frames 256 --out gpurun_out/cli.bmp 2>&1 | tail -1
