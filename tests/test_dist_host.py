"""Host logic of the multi-GPU sharding (cpuperformanceraytracer_b200/dist.py) on CPU: shard
arithmetic, and a world_size-2 gloo run of the spp-shard reduce / tile-shard gather paths with a
deterministic stand-in for the per-frame radiance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from cpuperformanceraytracer_b200 import dist as ptdist


def test_shard_frames_partition():
    for total in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                sh = ptdist.shard_frames(total, world, r, first_frame=5)
                seen += list(range(sh.first_frame, sh.first_frame + sh.nframes))
                assert abs(sh.nframes - total / world) < 1
            assert seen == list(range(5, 5 + total))  # contiguous, ordered, each frame exactly once
    with pytest.raises(ValueError):
        ptdist.shard_frames(8, 2, 2)


def test_shard_tile_rows_contiguous_spans():
    W, H, nty = 1920, 1080, 15
    for world in (1, 2, 4, 8):
        off = 0
        for r in range(world):
            sh = ptdist.shard_tile_rows(W, H, nty, world, r)
            assert sh.float_offset == off and sh.float_count == sh.num_tile_rows * (H // nty) * W * 3
            off += sh.float_count
        assert off == W * H * 3


def test_finalize_scale_is_biased_like_the_reference():
    assert ptdist.finalize_scale(1) == 0.5 and ptdist.finalize_scale(1023) == 1.0 / 1024.0


def _fake_radiance(npix, frame):
    # deterministic per-(pixel, frame) "colour", independent of which rank evaluates it
    i = torch.arange(npix * 3, dtype=torch.float64)
    return (torch.sin(i * 0.37 + frame * 1.13) * 0.5 + 0.5).to(torch.float32)


def _worker(rank, world, port, total_frames, npix, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    # spp-shard: SUM buffer of this rank's frame block, all-reduce, scale
    sh = ptdist.shard_frames(total_frames, world, rank)
    buf = torch.zeros(npix * 3, dtype=torch.float32)
    for f in range(sh.first_frame, sh.first_frame + sh.nframes):
        buf += _fake_radiance(npix, f)
    ptdist.reduce_sum_(buf)
    buf *= ptdist.finalize_scale(total_frames)
    # continued job (SppShardedRenderer.render(resume=True)): rank 0 holds the average after F frames, turns it back into a sum
    # (x (F + 1)), the other ranks start from zero, `more` further frames are sharded, reduced and scaled by 1/(F + more + 1)
    F, more = total_frames, 5
    cont = buf.clone() * float(F + 1) if rank == 0 else torch.zeros(npix * 3, dtype=torch.float32)
    sh2 = ptdist.shard_frames(more, world, rank, first_frame=F + 1)
    for f in range(sh2.first_frame, sh2.first_frame + sh2.nframes):
        cont += _fake_radiance(npix, f)
    ptdist.reduce_sum_(cont)
    cont *= ptdist.finalize_scale(F + more)
    np.save(os.path.join(out_dir, f"cont_{rank}.npy"), cont.numpy())
    # tile-shard: every rank fills its own contiguous span, all_gather reassembles
    W, H, nty = 16, 8, 4
    ts = ptdist.shard_tile_rows(W, H, nty, world, rank)
    full = torch.arange(W * H * 3, dtype=torch.float32)
    mine = full[ts.float_offset:ts.float_offset + ts.float_count].clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    tdist.all_gather(parts, mine)
    np.save(os.path.join(out_dir, f"spp_{rank}.npy"), buf.numpy())
    np.save(os.path.join(out_dir, f"tile_{rank}.npy"), torch.cat(parts).numpy())
    tdist.destroy_process_group()


def test_gloo_world2_spp_reduce_and_tile_gather(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total_frames, npix = 9, 40
    mp.spawn(_worker, args=(2, port, total_frames, npix, str(tmp_path)), nprocs=2, join=True)
    seq = torch.zeros(npix * 3, dtype=torch.float64)
    for f in range(1, total_frames + 1):
        seq += _fake_radiance(npix, f).double()
    seq = (seq / (total_frames + 1)).float().numpy()
    a, b = np.load(tmp_path / "spp_0.npy"), np.load(tmp_path / "spp_1.npy")
    assert np.array_equal(a, b)  # every rank ends with the same image
    assert np.allclose(a, seq, rtol=1e-6, atol=1e-7)
    # the continued job equals the reference's running average run over all 14 frames in order
    avg = np.zeros(npix * 3, dtype=np.float32)
    for f in range(1, total_frames + 5 + 1):
        c = _fake_radiance(npix, f).numpy()
        avg = avg + (c - avg) * np.float32(1.0 / (f + 1.0))
    c0 = np.load(tmp_path / "cont_0.npy")
    assert np.array_equal(c0, np.load(tmp_path / "cont_1.npy")) and np.allclose(c0, avg, rtol=2e-6, atol=2e-7)
    t0, t1 = np.load(tmp_path / "tile_0.npy"), np.load(tmp_path / "tile_1.npy")
    assert np.array_equal(t0, np.arange(16 * 8 * 3, dtype=np.float32)) and np.array_equal(t0, t1)


def test_cost_weighted_tile_rows_cover_the_image_once():
    """cost-balanced bands: contiguous, disjoint, complete; a cheap sky above an expensive scene moves the cut down"""
    costs = [1.0] * 6 + [10.0] * 6
    for world in (1, 2, 3, 5, 12, 16):
        rows, total = [], 0
        for r in range(world):
            sh = ptdist.shard_tile_rows(96, 120, 12, world, r, costs)
            rows += list(range(sh.first_tile_row, sh.first_tile_row + sh.num_tile_rows))
            assert sh.float_offset == sh.first_tile_row * 10 * 96 * 3 and sh.float_count == sh.num_tile_rows * 10 * 96 * 3
            total += sh.num_tile_rows
        assert rows == list(range(12)) and total == 12
    a, b = (ptdist.shard_tile_rows(96, 120, 12, 2, r, costs) for r in (0, 1))
    assert a.num_tile_rows > b.num_tile_rows  # 6 sky rows + some scene rows against the rest of the scene rows
    per_rank = [sum(costs[s.first_tile_row:s.first_tile_row + s.num_tile_rows]) for s in (a, b)]
    assert max(per_rank) <= 0.6 * sum(costs)
    import pytest
    with pytest.raises(ValueError):
        ptdist.shard_tile_rows(96, 120, 12, 2, 0, [1.0] * 5)


def test_tile_row_costs_from_cull_rects():
    rects = [(10.0, 20.0, 50.0, 60.0)]  # fragCoord space, y flipped: image rows 59 .. 99 of a 120-row image
    c = ptdist.tile_row_costs(rects, 96, 120, 12)
    assert len(c) == 12 and c[0] == c[1] == min(c) and max(c) == c[7] and c[11] == min(c)
    assert ptdist.tile_row_costs(None, 96, 120, 12) == [1.0] * 12
