#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
# offline CLI (C++ host mirror) vs oracle
./cpuperformanceraytracer_b200/render_offline --variant v2 --width 256 --height 128 --tiles-x 2 --tiles-y 4 --frames 6 --bounces 8 --dump-f32 gpurun_out/cli_v2.f32 --out gpurun_out/cli_v2.bmp
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from oracle import pyoracle as po
o,_ = po.render(po.PROFILE_V2, 256, 128, 2, 4, 8, 8)   # 2 warm-up + 6 frames
g = np.fromfile('gpurun_out/cli_v2.f32', np.float32)
print('render_offline v2 dump identical to oracle:', np.array_equal(g, o))
PY
