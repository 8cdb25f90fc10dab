#!/usr/bin/env bash
# multi-GPU visit 2: banded exchange tests + configs 3 / 5 (torchrun and C group)
set -u
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_multi2_n$N.log
for cfg in 3 5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --config $cfg 2>gpurun_out/cfg${cfg}_n$N.err | grep '^{' | tee -a gpurun_out/configs_n$N.jsonl
  tail -2 gpurun_out/cfg${cfg}_n$N.err
  timeout 900 python scripts/bench_configs.py --config $cfg --group $N --no-cpu-baseline 2>gpurun_out/cfg${cfg}_group_n$N.err | grep '^{' | tee -a gpurun_out/configs_n$N.jsonl
  tail -2 gpurun_out/cfg${cfg}_group_n$N.err
done
