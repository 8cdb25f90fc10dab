"""b200pt_group_*: several ranks behind one set of render entry points, in one process, without torch.

The sharding and combine logic is exercised on ONE GPU too: a group may place several ranks (contexts with their own
streams and SUM buffers) on the same device -- legal for tile sharding and for the library's own peer-memory combine
kernel.  With >= 2 GPUs the same tests run across devices, NCCL included.

Bars: tile sharding is BIT-IDENTICAL to the single-context render (hence to the oracle); spp sharding renders the same
samples in another summation order: <= 3e-6 relative (stated in include/b200pt.h)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cpuperformanceraytracer_b200 import api  # noqa: E402


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _device_sets():
    sets = [("3 ranks on gpu 0", [0, 0, 0])]
    n = _ngpus()
    if n >= 2:
        sets.append((f"{min(n, 4)} gpus", list(range(min(n, 4)))))
    return sets


W, H, NTX, NTY, BOUNCES = 192, 96, 2, 4, 8


def _sequential(frames, profile=api.PROFILE_V2, env=None, **kw):
    with api.Renderer(profile=profile, num_bounces=BOUNCES, **kw) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, NTX, NTY)
        out = []
        for f in frames:
            r.render_frames(f)
            out.append(r.download_target())
        return out


@pytest.mark.parametrize("label,devices", _device_sets(), ids=[s[0] for s in _device_sets()])
def test_tile_shard_group_is_bit_identical(oracle, label, devices):
    seq = _sequential([7, 5])
    o, _ = oracle.render(0, W, H, NTX, NTY, BOUNCES, 12)
    assert np.array_equal(seq[1], o)
    with api.Group(devices, sharding=api.SHARD_TILES, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(W, H, NTX, NTY)
        g.render_frames(7)
        assert np.array_equal(g.download_target(), seq[0])
        g.render_frames(5)  # continued: every rank keeps the running average of its own tiles
        assert np.array_equal(g.download_target(), seq[1])
        assert g.frame_counter == 12
        c = g.counters()
        assert c["paths"] == W * H * 12
        # the reference-facing call on a host buffer: spans travel over each rank's own link
        buf = np.zeros(W * H * 3, dtype=np.float32)
        g.frame_counter = 0
        g.render_host(buf, W, H, NTX, NTY, 7)
        assert np.array_equal(buf, seq[0])
        g.render_host(buf, W, H, NTX, NTY, 5)
        assert np.array_equal(buf, seq[1])
        # tone-mapped output of the gathered image == single context
        g.reset()
        g.render_frames(12)
        ldr = g.resolve_ldr()
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=BOUNCES) as r:
        r.resize(W, H, NTX, NTY)
        r.render_frames(12)
        assert np.array_equal(ldr, r.resolve_ldr())


@pytest.mark.parametrize("combine", [api.COMBINE_PEER, api.COMBINE_FUSED], ids=["peer", "fused"])
@pytest.mark.parametrize("label,devices", _device_sets(), ids=[s[0] for s in _device_sets()])
def test_spp_shard_group_peer_combine(label, devices, combine):
    """the library's own combines: a kernel that reads every rank's SUM buffer over peer memory (PEER), or render kernels that
    scatter their sums to the owners' staging slots while they run + a local finish (FUSED)"""
    seq = _sequential([24, 9])
    with api.Group(devices, sharding=api.SHARD_SPP, combine=combine, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(W, H, NTX, NTY)
        g.render_frames(24)
        a = g.download_target()
        assert np.allclose(a, seq[0], rtol=3e-6, atol=3e-6)
        g.render_frames(9)  # continued job: the average is turned back into a sum first
        b = g.download_target()
        assert np.allclose(b, seq[1], rtol=3e-6, atol=3e-6)
        # deterministic: the combine adds the partial sums in rank order
        g.reset()
        g.render_frames(24)
        assert np.array_equal(g.download_target(), a)
        # band-pipelined: the combine of a band of tile rows runs while the next band renders -- same sums, same order
        for bands in (2, 4, 9):  # (the fused combine has no bands: the exchange is already spread over the launch)
            g.set_bands(bands)
            g.reset()
            g.render_frames(24)
            assert np.array_equal(g.download_target(), a), bands
        g.set_bands(0)
        c = g.counters()
        assert c["paths"] == W * H * (24 + 9 + 24 + 3 * 24) and c["combine_ms"] > 0.0
        # fewer frames than ranks: some ranks render nothing
        g.reset()
        g.render_frames(2)
        with api.Renderer(profile=api.PROFILE_V2, num_bounces=BOUNCES) as r:
            r.resize(W, H, NTX, NTY)
            r.render_frames(2)
            assert np.allclose(g.download_target(), r.download_target(), rtol=3e-6, atol=3e-6)
        # host-buffer call, continued from a non-empty accumulation state
        buf = seq[0].copy()
        g.frame_counter = 24
        g.render_host(buf, W, H, NTX, NTY, 9)
        assert np.allclose(buf, seq[1], rtol=3e-6, atol=3e-6)


def test_tile_shard_group_other_tile_shapes(oracle):
    """tiles whose area is not a multiple of 32 pixels cannot interleave (work items would straddle tiles): the group falls
    back to contiguous, cost-balanced tile ranges -- still bit-identical"""
    w, h, ntx, nty = 72, 15, 3, 5  # tiles of 24 x 3 pixels: 9 SoA8 groups
    o, _ = oracle.render(0, w, h, ntx, nty, BOUNCES, 9)
    with api.Group([0, 0, 0], sharding=api.SHARD_TILES, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(w, h, ntx, nty)
        g.render_frames(4)
        g.render_frames(5)
        assert np.array_equal(g.download_target(), o)
        buf = np.zeros(w * h * 3, dtype=np.float32)
        g.frame_counter = 0
        g.render_host(buf, w, h, ntx, nty, 9)
        assert np.array_equal(buf, o)


def test_tile_stride_on_one_context(oracle):
    """b200pt_set_tile_stride: every modulus-th tile per launch; the residues assemble the full render, bit for bit"""
    w, h, ntx, nty, frames = 192, 120, 3, 5, 9
    o, oc = oracle.render(0, w, h, ntx, nty, BOUNCES, frames)
    for sched in (api.SCHED_LANE, api.SCHED_SORTED):
        with api.Renderer(profile=api.PROFILE_V2, num_bounces=BOUNCES, scheduler=sched) as r:
            r.resize(w, h, ntx, nty)
            for mod in (4, 1, 16):
                r.reset()
                for rem in range(mod):
                    r.frame_counter = 0
                    r.set_tile_stride(rem, mod)
                    r.render_frames(frames)
                assert np.array_equal(r.download_target(), o), (sched, mod)
            r.set_tile_stride(0, 0)
            with pytest.raises(api.B200PTError):
                r.set_tile_stride(3, 3)
        c = None
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=BOUNCES) as r:
        r.resize(72, 15, 3, 5)
        with pytest.raises(api.B200PTError, match="32 pixels"):
            r.set_tile_stride(0, 2)


def test_group_with_env_profile(oracle):
    env = oracle.synthetic_env(64, 384)
    seq = _sequential([10], profile=api.PROFILE_OPT_V4, env=env, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM)[0]
    o, _ = oracle.render(2, W, H, NTX, NTY, BOUNCES, 10, env=env, env_kind=2, env_sampler=2)
    assert np.array_equal(seq, o)
    for sharding, combine, sched in ((api.SHARD_TILES, api.COMBINE_PEER, api.SCHED_DEFAULT), (api.SHARD_SPP, api.COMBINE_PEER, api.SCHED_DEFAULT),
                                     (api.SHARD_SPP, api.COMBINE_FUSED, api.SCHED_DEFAULT), (api.SHARD_SPP, api.COMBINE_FUSED, api.SCHED_SORTED)):
        with api.Group([0, 0], sharding=sharding, combine=combine, profile=api.PROFILE_OPT_V4, num_bounces=BOUNCES,
                       env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM, scheduler=sched) as g:
            g.set_env(env)
            g.resize(W, H, NTX, NTY)
            g.render_frames(10)
            out = g.download_target()
        if sharding == api.SHARD_TILES:
            assert np.array_equal(out, o)
        else:
            assert np.allclose(out, o, rtol=3e-6, atol=3e-6)


@pytest.mark.skipif(_ngpus() < 2, reason="NCCL wants one device per rank: needs at least 2 GPUs")
def test_spp_shard_group_nccl():
    seq = _sequential([24, 9])
    n = min(_ngpus(), 4)
    with api.Group(list(range(n)), sharding=api.SHARD_SPP, combine=api.COMBINE_NCCL, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(W, H, NTX, NTY)
        g.render_frames(24)
        assert np.allclose(g.download_target(), seq[0], rtol=3e-6, atol=3e-6)
        g.render_frames(9)
        assert np.allclose(g.download_target(), seq[1], rtol=3e-6, atol=3e-6)
        g.set_bands(3)  # NCCL reduce per band on the second streams
        g.reset()
        g.render_frames(24)
        assert np.allclose(g.download_target(), seq[0], rtol=3e-6, atol=3e-6)


def test_group_at_baseline_image_sizes(oracle):
    """BASELINE geometry: 1920x1080 tiles 10x15 (config 2) against the oracle, 4096x4096 tiles 16x64 (config 5's tiling at a
    quarter of its pixels) against one context: interleaved tiles bit-identical, fused / peer spp combine within 3e-6"""
    devices = list(range(min(_ngpus(), 8))) if _ngpus() >= 2 else [0, 0, 0]
    w, h, ntx, nty, frames = 1920, 1080, 10, 15, 4
    o, _ = oracle.render(0, w, h, ntx, nty, BOUNCES, frames)
    with api.Group(devices, sharding=api.SHARD_TILES, profile=api.PROFILE_V2, num_bounces=BOUNCES) as g:
        g.resize(w, h, ntx, nty)
        g.render_frames(frames)
        assert np.array_equal(g.download_target(), o)
    w, h, ntx, nty, frames = 4096, 4096, 16, 64, 3
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=16) as r:
        r.resize(w, h, ntx, nty)
        r.render_frames(frames)
        seq = r.download_target()
    with api.Group(devices, sharding=api.SHARD_TILES, profile=api.PROFILE_V2, num_bounces=16) as g:
        g.resize(w, h, ntx, nty)
        g.render_frames(frames)
        assert np.array_equal(g.download_target(), seq)
    for combine in (api.COMBINE_FUSED, api.COMBINE_PEER):
        with api.Group(devices, sharding=api.SHARD_SPP, combine=combine, profile=api.PROFILE_V2, num_bounces=16) as g:
            g.resize(w, h, ntx, nty)
            g.render_frames(frames)
            assert np.allclose(g.download_target(), seq, rtol=3e-6, atol=3e-6)


def test_group_argument_checking():
    import ctypes
    lib = api.load_library()
    p = api.default_params(api.PROFILE_V2)
    g = ctypes.c_void_p()
    devs = (ctypes.c_int32 * 2)(0, 0)
    assert lib.b200pt_group_create(ctypes.byref(p), devs, 0, 0, 0, ctypes.byref(g)) == 1
    assert lib.b200pt_group_create(ctypes.byref(p), devs, 2, 7, 0, ctypes.byref(g)) == 1
    # NCCL cannot place two ranks on one device
    assert lib.b200pt_group_create(ctypes.byref(p), devs, 2, api.SHARD_SPP, api.COMBINE_NCCL, ctypes.byref(g)) == 1
    with api.Group([0], profile=api.PROFILE_V2) as one:
        with pytest.raises(api.B200PTError):
            one.render_frames(1)  # resize first
