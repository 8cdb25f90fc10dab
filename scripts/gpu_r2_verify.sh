#!/usr/bin/env bash
# last visit of the round: the GPU suite, smoke(), both bench arms as the driver runs them
set -u
mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = 1 ]; then
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/verify_pytest_n1.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/verify_smoke.log
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/verify_ref.err | tee gpurun_out/verify_bench_reference.json
  timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/verify_bench.err | tee gpurun_out/verify_bench_n1.json
else
  nvidia-smi -L | head -8
  timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py tests/test_gpu_host_mirror.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/verify_pytest_n$N.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/verify_bench_n$N.err | tee gpurun_out/verify_bench_n$N.json
  tail -2 gpurun_out/verify_bench_n$N.err
fi
