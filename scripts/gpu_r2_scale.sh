#!/usr/bin/env bash
# N-GPU visit: multi-GPU tests, bench at N (strong scaling + parity_check), CLI group modes, configs 3 / 5 (torchrun + C group)
set -u
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py tests/test_gpu_host_mirror.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -2 gpurun_out/bench_n$N.err
for mode in "--shard spp --combine nccl" "--shard spp --combine peer" "--shard tiles"; do
  timeout 300 ./cpuperformanceraytracer_b200/render_offline --variant v2 --bounces 8 --width 1920 --height 1080 --tiles-x 10 --tiles-y 15 --frames 1024 --gpus $N $mode --out gpurun_out/cli.bmp 2>&1 | tail -2 | tee -a gpurun_out/cli_n$N.log
done
rm -f gpurun_out/configs_n$N.jsonl
for cfg in 3 5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --config $cfg --no-cpu-baseline 2>gpurun_out/cfg${cfg}_n$N.err | grep '^{' | tee -a gpurun_out/configs_n$N.jsonl
  timeout 900 python scripts/bench_configs.py --config $cfg --group $N --no-cpu-baseline 2>gpurun_out/cfg${cfg}_group_n$N.err | grep '^{' | tee -a gpurun_out/configs_n$N.jsonl
done
