#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the megakernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --math fast --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_fast.json
python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
if [ "${NCU:-1}" = 1 ]; then
  python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -o gpurun_out/prof -f \
      python bench.py --spp 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
