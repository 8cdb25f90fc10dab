"""Run-time scene description for the OPT_V4 profile: oracle-side pinning of the
API path against the built-in (reference-pinned) scene, culling bounds of arbitrary scenes, and GPU
parity on scenes with a different primitive count."""
import numpy as np
import pytest

from cpuperformanceraytracer_b200 import api
from scene_fixtures import default_v4_scene, random_v4_scene


def test_oracle_default_scene_through_the_api_equals_builtin(oracle):
    q, s, m = default_v4_scene()
    env = oracle.synthetic_env(128, 64)
    a, ca = oracle.render(oracle.PROFILE_V4, 128, 72, 4, 6, 8, 4, env=env, env_kind=1, env_sampler=2)
    b, cb = oracle.render(oracle.PROFILE_V4, 128, 72, 4, 6, 8, 4, env=env, env_kind=1, env_sampler=2,
                          scene_v4=oracle.make_scene_v4(q, s, m, (0.0, 0.0, 40.0), oracle.lib().oracle_camera_distance()))
    assert np.array_equal(a, b) and ca == cb


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_culling_bounds_of_runtime_scenes(oracle, seed):
    q, s, m = random_v4_scene(seed)
    cam, dist, W, H = (1.0, 2.0, 38.0), 1.2, 240, 136
    rects = api.cull_rects_scene_v4(q, s, cam, dist, W, H)
    assert rects is not None and len(rects) == len(q) + len(s)
    xs = np.arange(W, dtype=np.float32)[None, :]
    yf = (H - 1 - np.arange(H, dtype=np.float32))[:, None]
    hit = np.zeros((H, W), dtype=bool)
    for x0, y0, x1, y1 in rects:
        hit |= (xs + 0.5 >= x0) & (xs - 0.5 <= x1) & (yf + 0.5 >= y0) & (yf - 0.5 <= y1)
    seg = oracle.max_segments(oracle.PROFILE_V4, W, H, 8, 12, scene_v4=oracle.make_scene_v4(q, s, m, cam, dist))
    assert (seg[~hit] == 1).all()


def test_culling_is_disabled_when_geometry_reaches_the_camera_plane():
    q, s, m = random_v4_scene(1)
    q[1, :, 2] += 60.0  # behind the camera
    assert api.cull_rects_scene_v4(q, s, (0.0, 0.0, 40.0), 1.0, 128, 72) is None


@pytest.mark.gpu
def test_gpu_default_scene_through_the_api_is_bit_identical(oracle):
    q, s, m = default_v4_scene()
    env = oracle.synthetic_env(128, 64)
    W, H, ntx, nty, frames = 192, 108, 4, 6, 8
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8) as r:
        r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        builtin = r.download_target()
        r.set_scene_v4(q, s, m, (0.0, 0.0, 40.0), float(oracle.lib().oracle_camera_distance()))
        r.reset()
        r.render_frames(frames)
        via_api = r.download_target()
        r.set_scene_v4()  # back to InitializeScene's scene
        r.reset()
        r.render_frames(frames)
        again = r.download_target()
    assert np.array_equal(builtin, via_api) and np.array_equal(builtin, again)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nq,ns,ek,es", [(1, 5, 6, 1, 2), (2, 2, 10, 2, 1), (3, 8, 1, 0, 0)])
def test_gpu_runtime_scene_bit_exact_vs_oracle(oracle, seed, nq, ns, ek, es):
    q, s, m = random_v4_scene(seed, nq, ns)
    cam, dist = (0.5, 1.0, 39.0), 1.1
    env = oracle.synthetic_env(128, 64) if ek == 1 else (oracle.synthetic_env(32, 192) if ek == 2 else None)
    W, H, ntx, nty, frames = 192, 112, 4, 7, 8
    o, oc = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, 8, frames, env=env, env_kind=ek, env_sampler=es,
                          scene_v4=oracle.make_scene_v4(q, s, m, cam, dist))
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=ek, env_sampler=es if ek else api.SAMPLER_RANDOM) as r:
        if env is not None:
            r.set_env(env)
        r.set_scene_v4(q, s, m, cam, dist)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
        c = r.counters()
    assert np.array_equal(g, o)
    assert (c["segments"], c["escapes"]) == (oc["segments"], oc["escapes"])


@pytest.mark.gpu
def test_gpu_runtime_scene_errors():
    q, s, m = random_v4_scene(1, 7, 6)  # 13 objects > MAX_OBJECTS
    with api.Renderer(profile=api.PROFILE_OPT_V4) as r:
        with pytest.raises(api.B200PTError):
            r.set_scene_v4(q, s, m)
    with api.Renderer(profile=api.PROFILE_V2) as r:
        q, s, m = random_v4_scene(1)
        with pytest.raises(api.B200PTError, match="OPT_V4"):
            r.set_scene_v4(q, s, m)


# ---- Cornell-family profiles (V2, SIMT_TEXTURED): b200pt_set_scene_cornell ------------------------------------------
from scene_fixtures import default_cornell_scene, random_cornell_scene  # noqa: E402


@pytest.mark.parametrize("profile,simt", [(0, False), (1, True)], ids=["v2", "simt_textured"])
def test_oracle_default_cornell_scene_through_the_api_equals_builtin(oracle, profile, simt):
    q, s, m = default_cornell_scene(simt)
    env = oracle.synthetic_env(128, 64) if simt else None
    kw = dict(env=env, env_kind=1 if simt else 0)
    a, ca = oracle.render(profile, 128, 72, 4, 6, 8, 4, **kw)
    b, cb = oracle.render(profile, 128, 72, 4, 6, 8, 4, scene_cornell=oracle.make_scene_cornell(q, s, m), **kw)
    assert np.array_equal(a, b) and ca == cb


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_culling_bounds_of_runtime_cornell_scenes(oracle, seed):
    q, s, m = random_cornell_scene(seed)
    W, H = 240, 136
    rects = api.cull_rects_scene_cornell(q, s, W, H)
    assert rects is not None and len(rects) == 9
    xs = np.arange(W, dtype=np.float32)[None, :]
    yf = (H - 1 - np.arange(H, dtype=np.float32))[:, None]
    hit = np.zeros((H, W), dtype=bool)
    for x0, y0, x1, y1 in rects:
        hit |= (xs + 0.5 >= x0) & (xs - 0.5 <= x1) & (yf + 0.5 >= y0) & (yf - 0.5 <= y1)
    seg = oracle.max_segments(oracle.PROFILE_V2, W, H, 8, 12, scene_cornell=oracle.make_scene_cornell(q, s, m))
    assert (~hit).any() and (seg[~hit] == 1).all()
    # built-in scene through the data path: its rectangles contain the one the specialised path uses
    q0, s0, _ = default_cornell_scene()
    r0 = api.cull_rects_scene_cornell(q0, s0, W, H)
    b = api.cull_rects(api.PROFILE_V2, W, H)[0]
    assert r0[:, 0].min() >= b[0] - 1 and r0[:, 2].max() <= b[2] + 1 and r0[:, 1].min() >= b[1] - 1 and r0[:, 3].max() <= b[3] + 1


@pytest.mark.gpu
@pytest.mark.parametrize("profile,oprofile,simt", [(api.PROFILE_V2, 0, False), (api.PROFILE_SIMT_TEXTURED, 1, True)], ids=["v2", "simt_textured"])
def test_gpu_default_cornell_scene_through_the_api_is_bit_identical(oracle, profile, oprofile, simt):
    q, s, m = default_cornell_scene(simt)
    env = oracle.synthetic_env(128, 64) if simt else None
    W, H, ntx, nty, frames = 192, 108, 4, 6, 8
    with api.Renderer(profile=profile, num_bounces=8) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        builtin = r.download_target()
        r.set_scene_cornell(q, s, m)
        r.reset()
        r.render_frames(frames)
        via_api = r.download_target()
        r.set_scene_cornell()  # back to the literals of v2.cpp:320-454
        r.reset()
        r.render_frames(frames)
        again = r.download_target()
    assert np.array_equal(builtin, via_api) and np.array_equal(builtin, again)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_NONE) as r:
        with pytest.raises(api.B200PTError):
            r.set_scene_cornell(q, s, m)
    with api.Renderer(profile=profile, num_bounces=8) as r:
        bad = s.copy()
        bad[1, 3] = 0.0
        with pytest.raises(api.B200PTError, match="radii"):
            r.set_scene_cornell(q, bad, m)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,profile,oprofile,bounces", [(1, api.PROFILE_V2, 0, 8), (2, api.PROFILE_V2, 0, 4), (3, api.PROFILE_SIMT_TEXTURED, 1, 4),
                                                           (4, api.PROFILE_V2, 0, 16)])
def test_gpu_runtime_cornell_scene_bit_exact_vs_oracle(oracle, seed, profile, oprofile, bounces):
    q, s, m = random_cornell_scene(seed)
    env = oracle.synthetic_env(128, 64) if oprofile == 1 else None
    W, H, ntx, nty, frames = 192, 112, 4, 7, 8
    o, oc = oracle.render(oprofile, W, H, ntx, nty, bounces, frames, env=env, env_kind=1 if oprofile == 1 else 0,
                          scene_cornell=oracle.make_scene_cornell(q, s, m))
    for sched in (api.SCHED_LANE, api.SCHED_SORTED):
        with api.Renderer(profile=profile, num_bounces=bounces, scheduler=sched) as r:
            if env is not None:
                r.set_env(env)
            r.set_scene_cornell(q, s, m)
            r.resize(W, H, ntx, nty)
            r.render_frames(frames)
            g = r.download_target()
            c = r.counters()
        assert np.array_equal(g, o), (sched, float(np.abs(g - o).max()))
        assert (c["segments"], c["escapes"]) == (oc["segments"], oc["escapes"])
