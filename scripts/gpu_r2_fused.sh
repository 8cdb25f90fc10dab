#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_host_mirror.py tests/test_gpu_sorted.py -m gpu -q 2>&1 | tail -8
if [ "$N" -gt 1 ]; then
rm -f gpurun_out/configs_group_fused_n$N.jsonl
for cfg in 3 5; do
  timeout 900 python scripts/bench_configs.py --config $cfg --group $N --no-cpu-baseline 2>gpurun_out/cfg${cfg}_group_n$N.err | grep '^{' | tee -a gpurun_out/configs_group_fused_n$N.jsonl
  tail -2 gpurun_out/cfg${cfg}_group_n$N.err
done
for mode in "--combine nccl" "--combine peer" "--combine fused"; do
  timeout 300 ./cpuperformanceraytracer_b200/render_offline --variant v2 --bounces 8 --width 1920 --height 1080 --tiles-x 10 --tiles-y 15 --frames 1024 --gpus $N --shard spp $mode --out gpurun_out/cli.bmp 2>&1 | tail -2
done
fi
