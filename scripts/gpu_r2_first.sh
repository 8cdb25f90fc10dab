#!/usr/bin/env bash
# round 2, first visit (1 GPU): the whole -m gpu suite, smoke, bench (ours + reference arm)
set -u
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
