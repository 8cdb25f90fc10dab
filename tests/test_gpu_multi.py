"""Multi-GPU spp-shard render (NCCL) vs the sequential single-GPU render.  Needs >= 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as tdist
sys.path.insert(0, %r)
from cpuperformanceraytracer_b200 import api, dist as ptdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, ntx, nty, total = 192, 96, 2, 4, 24
factory = lambda accum_mode=api.ACCUM_RUNNING_AVERAGE, device=local: api.Renderer(profile=api.PROFILE_V2, num_bounces=8, device=device, accum_mode=accum_mode)
sr = ptdist.SppShardedRenderer(factory, W, H, ntx, nty, rank, world, local)
buf = sr.render(total)
sr.stream.synchronize()
out = buf.cpu().numpy()
buf = sr.render(total, bands=3)   # all-reduce per band of tile rows, overlapped with the next band's render
sr.stream.synchronize()
banded = buf.cpu().numpy()
# continued job: 10 more frames on top of the reduced image (rank 0's buffer holds the average after `total` calls)
buf = sr.render(10, first_frame=total + 1, resume=True)
sr.stream.synchronize()
cont = buf.cpu().numpy()
tr = ptdist.TileShardedRenderer(factory, W, H, ntx, nty, rank, world, local)
tbuf = tr.render(total)
tr.stream.synchronize()
tout = tbuf.cpu().numpy()
if rank == 0:
    with factory() as r:
        r.resize(W, H, ntx, nty); r.render_frames(total); seq = r.download_target()
        r.render_frames(10); seq2 = r.download_target()
    np.save(sys.argv[1], np.stack([out, seq, tout, banded, cont, seq2]))
tdist.barrier(); tdist.destroy_process_group()
'''


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_spp_shard_matches_sequential(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    out = tmp_path / "res.npy"
    n = min(_ngpus(), 4)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                    "127.0.0.1", "--master-port", "29541", str(script), str(out)], check=True, timeout=600)
    sharded, seq, tiled, banded, cont, seq2 = np.load(out)
    # same samples, different summation order (sum then scale vs running average): ~1e-6 relative
    assert np.allclose(sharded, seq, rtol=3e-6, atol=3e-6)
    # tile-shard: bit-identical to the single-GPU render
    assert np.array_equal(tiled, seq)
    # band-pipelined exchange: the same sums (NCCL may pick another reduction order per message size: tolerance, not bits)
    assert np.allclose(banded, seq, rtol=3e-6, atol=3e-6)
    # resume=True continues the job from the reduced image
    assert np.allclose(cont, seq2, rtol=3e-6, atol=3e-6)
