"""B200PT_SCHED_SORTED (pt_render_sorted_kernel: every loop trip the CTA sorts its 256 paths by what they need next)
must reproduce the oracle BIT FOR BIT exactly like the per-lane kernel: f32 buffers, per-pixel wang_hash states after
the last path, segment / escape / culled counters, for every profile and sampler, at any tiling, frame chunking,
accumulation mode, tile range and bounce count -- only WHICH thread evaluates a path segment differs."""
import ctypes

import numpy as np
import pytest

from conftest import stats
from test_gpu_parity import CONFIGS, make_renderer

pytestmark = pytest.mark.gpu

from cpuperformanceraytracer_b200 import api  # noqa: E402

SORTED = dict(scheduler=api.SCHED_SORTED)


@pytest.mark.parametrize("name,profile,envshape,ek,es,bounces", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_sorted_scheduler_bit_exact_vs_oracle(oracle, name, profile, envshape, ek, es, bounces):
    W, H, ntx, nty, frames = 256, 192, 4, 6, 12
    env = oracle.synthetic_env(*envshape) if envshape else None
    o, oc = oracle.render(profile, W, H, ntx, nty, bounces, frames, env=env, env_kind=ek, env_sampler=es)
    with make_renderer(profile, bounces, ek, es, **SORTED) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
        assert np.array_equal(g, o), "max abs %g rmse %g" % stats(g, o)
        c = r.counters()
        assert (c["paths"], c["segments"], c["escapes"]) == (oc["paths"], oc["segments"], oc["escapes"])
        rs = r.rng_state()
        p, keep = oracle.make_params(profile, W, H, ntx, nty, bounces, env, ek, es)
        rng = np.random.default_rng(5)
        for _ in range(300):
            x, y = int(rng.integers(W)), int(rng.integers(H))
            assert int(rs[y, x]) == oracle.lib().oracle_final_rng_state(ctypes.byref(p), x, y, frames)


def test_sorted_equals_lane_scheduler_everywhere(oracle):
    """same buffers, RNG states and counters as the per-lane kernel: ragged images, tile ranges, frame chunks,
    SUM mode, culling off, generic scene tables, fused tone map"""
    env = oracle.synthetic_env(128, 64)
    cases = [
        dict(profile=api.PROFILE_V2, num_bounces=8), dict(profile=api.PROFILE_V2, num_bounces=0),
        dict(profile=api.PROFILE_V2, num_bounces=16, accum_mode=api.ACCUM_SUM),
        dict(profile=api.PROFILE_V2, num_bounces=4, disable_camera_culling=True, generic_scene_tables=True),
        dict(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM, output_to_screen=True),
        dict(profile=api.PROFILE_OPT_V4, num_bounces=3, env_kind=api.ENV_NONE, generic_scene_tables=True),
        dict(profile=api.PROFILE_SIMT_TEXTURED, num_bounces=4), dict(profile=api.PROFILE_V3_REDO, num_bounces=8),
        dict(profile=api.PROFILE_V3_REDO_SCENE0, num_bounces=5),
    ]
    sizes = [(8, 1, 1, 1), (24, 5, 3, 5), (40, 3, 1, 3), (320, 200, 4, 5), (264, 130, 3, 2)]
    for kw in cases:
        for (W, H, ntx, nty) in sizes:
            res = []
            for sched in (api.SCHED_LANE, api.SCHED_SORTED):
                with api.Renderer(scheduler=sched, **kw) as r:
                    if kw["profile"] != api.PROFILE_V2 and kw.get("env_kind", 1) != api.ENV_NONE:
                        r.set_env(env)
                    r.resize(W, H, ntx, nty)
                    r.render_frames(0)
                    r.render_frames(3)
                    r.render_frames(7)
                    if nty > 1:  # a band of tile rows rendered again on top
                        r.set_tile_row_range(1, nty - 1)
                        r.render_frames(2)
                    c = r.counters()
                    ldr = r.resolve_ldr() if kw.get("output_to_screen") else None
                    res.append((r.download_target(), r.rng_state(), (c["paths"], c["segments"], c["escapes"], c["culled_segments"]), ldr))
            assert np.array_equal(res[0][0], res[1][0]), (kw, W, H)
            assert np.array_equal(res[0][1], res[1][1]), (kw, W, H)
            assert res[0][2] == res[1][2], (kw, W, H)
            if res[0][3] is not None:
                assert np.array_equal(res[0][3], res[1][3])


def test_sorted_full_size(oracle):
    """BASELINE config 2 geometry (1920x1080, tiles 10x15, 8 bounces), bounded spp: whole buffer vs the oracle"""
    W, H, ntx, nty, frames = 1920, 1080, 10, 15, 6
    o, oc = oracle.render(oracle.PROFILE_V2, W, H, ntx, nty, 8, frames)
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, **SORTED) as r:
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        assert np.array_equal(r.download_target(), o)
        c = r.counters()
        assert (c["segments"], c["escapes"]) == (oc["segments"], oc["escapes"])


def test_sorted_present_ring(oracle):
    """the fused tone map of the sorted kernel feeds the present ring like the per-lane kernel's"""
    W, H, ntx, nty = 128, 72, 4, 6
    cube = oracle.synthetic_env(32, 192)
    frames = {}
    for sched in (api.SCHED_LANE, api.SCHED_SORTED):
        with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM,
                          output_to_screen=True, scheduler=sched) as r:
            r.set_env(cube)
            r.resize(W, H, ntx, nty)
            out = []
            for _ in range(4):
                r.present_submit(1)
                out.append(r.present_acquire()[0])
            frames[sched] = out
    for a, b in zip(frames[api.SCHED_LANE], frames[api.SCHED_SORTED]):
        assert np.array_equal(a, b)
