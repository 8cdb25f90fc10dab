"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars:
  * B200PT_MATH_PARITY: BIT-EXACT f32 accumulation buffers, RNG states and segment counters.
  * B200PT_MATH_FAST (FMA contraction + MUFU approximations, same RNG streams): tolerances stated
    per test (RMSE over the f32 buffer at matched seeds/spp).
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, stats

pytestmark = pytest.mark.gpu

from cpuperformanceraytracer_b200 import api  # noqa: E402

GPU_PROFILE = {0: api.PROFILE_V2, 1: api.PROFILE_SIMT_TEXTURED, 2: api.PROFILE_OPT_V4, 3: api.PROFILE_V3_REDO, 4: api.PROFILE_V3_REDO_SCENE0}


def make_renderer(case_profile, bounces, env_kind, env_sampler, math_mode=api.MATH_PARITY, **kw):
    gp = GPU_PROFILE[case_profile]
    if gp == api.PROFILE_OPT_V4:
        return api.Renderer(profile=gp, math_mode=math_mode, num_bounces=bounces, env_kind=env_kind,
                            env_sampler=env_sampler, **kw)
    return api.Renderer(profile=gp, math_mode=math_mode, num_bounces=bounces, **kw)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_parity_mode_reproduces_reference_golden(oracle, case):
    """GPU vs buffers produced by the reference's own code (tests/golden)."""
    g = load_golden(case["name"])
    env = oracle.synthetic_env(*case["env_shape"]) if case["env_shape"] else None
    flags = case.get("v4_flags", 0)  # the non-default sides of global_preprocessor_flags.h:64-65
    kw = dict(exact_exp=bool(flags & oracle.V4_EXACT_EXP), sincos_unit_vectors=bool(flags & oracle.V4_SINCOS_UNIT_VECTORS)) if flags else {}
    with make_renderer(case["profile"], case["bounces"], case["env_kind"], case["env_sampler"], **kw) as r:
        if env is not None:
            r.set_env(env)
        r.resize(case["width"], case["height"], case["ntx"], case["nty"])
        r.render_frames(case["frames"])
        assert np.array_equal(r.download_target(), g["buffer"])
        r.render_frames(case["continued_frames"])
        assert np.array_equal(r.download_target(), g["continued"])
        assert r.frame_counter == case["frames"] + case["continued_frames"]


CONFIGS = [
    ("v2", 0, None, 0, 0, 8),
    ("simt_textured", 1, (256, 128), 1, 0, 4),
    ("v4_equirect_random", 2, (256, 128), 1, 2, 8),
    ("v4_equirect_bilinear", 2, (256, 128), 1, 1, 8),
    ("v4_cubemap_random", 2, (64, 384), 2, 2, 8),
    ("v4_cubemap_bilinear", 2, (64, 384), 2, 1, 8),
    ("v4_no_env", 2, None, 0, 0, 8),
    ("v3_redo", 3, (256, 128), 1, 1, 8),
    ("v3_redo_scene0", 4, (256, 128), 1, 1, 8),
]


@pytest.mark.parametrize("name,profile,envshape,ek,es,bounces", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_parity_mode_bit_exact_vs_oracle(oracle, name, profile, envshape, ek, es, bounces):
    """Seeded 256x192, 12 frames: buffer, per-pixel RNG state after the last path, counters."""
    W, H, ntx, nty, frames = 256, 192, 4, 6, 12
    env = oracle.synthetic_env(*envshape) if envshape else None
    o, oc = oracle.render(profile, W, H, ntx, nty, bounces, frames, env=env, env_kind=ek, env_sampler=es)
    with make_renderer(profile, bounces, ek, es) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
        assert np.array_equal(g, o), "max abs %g rmse %g" % stats(g, o)
        c = r.counters()
        assert (c["paths"], c["segments"], c["escapes"]) == (oc["paths"], oc["segments"], oc["escapes"])
        # wang_hash stream parity: state of a sample of pixels after frame `frames`
        rs = r.rng_state()
        p, keep = oracle.make_params(profile, W, H, ntx, nty, bounces, env, ek, es)
        import ctypes
        rng = np.random.default_rng(3)
        for _ in range(300):
            x, y = int(rng.integers(W)), int(rng.integers(H))
            assert int(rs[y, x]) == oracle.lib().oracle_final_rng_state(ctypes.byref(p), x, y, frames)


def _smooth_env(w, h):
    """what an HDR photograph looks like at texel scale: neighbouring texels nearly equal, plus a small sun"""
    y, x = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    e = np.stack([0.6 + 0.5 * np.sin(6.283 * x) * y, 0.5 + 0.4 * np.cos(6.283 * 2 * x), 0.3 + 1.5 * y * y], axis=2)
    e[(x - 0.25) ** 2 + (y - 0.75) ** 2 < 0.02 ** 2] = 50.0
    return e.astype(np.float32)


# B200PT_MATH_FAST = FMA contraction + MUFU rcp / rsqrt / sqrt / sincos + CUDA atan2f / asinf / __expf: ULP-level
# perturbations of every ray, identical RNG streams.  A perturbed path occasionally takes another discrete decision
# (which object at a silhouette, which texel), which changes that SAMPLE by O(1): the image error is the Monte-Carlo
# error of those rare flips and falls like 1/sqrt(spp).  Measured on B200 (scripts/fast_math_rmse.py,
# profiles/r02_b_fast_math_rmse.jsonl), RMSE over the f32 buffer at 256x192, 64 -> 1024 spp:
#   v2 1.5e-3 -> 4.0e-4 | v4 equirect random 5.9e-4 -> 1.8e-4 | v4 equirect bilinear 1.9e-4 -> 1.2e-4 |
#   v4 cubemap random 2.0e-3 -> 3.6e-4 (per-texel-noise env; 9.6e-5 on a smooth env) | v3_redo 3.5e-3 -> 8.4e-4
# i.e. <= 1e-3 at the headline sample count for every profile with a jittered camera.
# simt_textured is the exception and NOT a sampling error: its camera rays carry no jitter (simt_textured.cpp:433-474),
# so a pixel whose ray passes exactly through a seam between two quads (u == 0 or w == 0 in exact arithmetic) takes
# the same side in every frame, and FMA contraction can move it to the other side: a fixed set of seam pixels
# (0.2 % of the image) differs at O(1) whatever the spp (RMSE 0.077 on the noise env, 0.03 on a smooth one).
# The tolerances below are 2x the measured values; the bit-exact mode (B200PT_MATH_PARITY) is the parity claim.
FAST_CASES = [
    ("v2", 0, None, 0, 0, 8, "noise", 64, 3e-3),
    ("v2", 0, None, 0, 0, 8, "noise", 1024, 1e-3),
    ("v4_equirect_random", 2, (256, 128), 1, 2, 8, "noise", 64, 1.5e-3),
    ("v4_equirect_random", 2, (256, 128), 1, 2, 8, "noise", 1024, 1e-3),
    ("v4_equirect_bilinear", 2, (256, 128), 1, 1, 8, "noise", 64, 1e-3),
    ("v4_cubemap_random", 2, (64, 384), 2, 2, 8, "noise", 64, 4e-3),
    ("v4_cubemap_random", 2, (64, 384), 2, 2, 8, "noise", 1024, 1e-3),
    ("v4_cubemap_random", 2, (64, 384), 2, 2, 8, "smooth", 64, 1e-3),
    ("v3_redo", 3, (256, 128), 1, 1, 8, "noise", 64, 7e-3),
    ("v3_redo", 3, (256, 128), 1, 1, 8, "smooth", 1024, 1.7e-3),
    ("simt_textured", 1, (256, 128), 1, 0, 4, "noise", 64, 0.16),
    ("simt_textured", 1, (256, 128), 1, 0, 4, "smooth", 64, 0.07),
]


@pytest.mark.parametrize("name,profile,envshape,ek,es,bounces,envname,frames,tol", FAST_CASES,
                         ids=["%s-%s-%dspp" % (c[0], c[6], c[7]) for c in FAST_CASES])
def test_fast_mode_within_tolerance(oracle, name, profile, envshape, ek, es, bounces, envname, frames, tol):
    import os
    W, H, ntx, nty = 256, 192, 4, 6
    env = None if not envshape else (oracle.synthetic_env(*envshape) if envname == "noise" else _smooth_env(*envshape))
    o, oc = oracle.render(profile, W, H, ntx, nty, bounces, frames, env=env, env_kind=ek, env_sampler=es, nthreads=os.cpu_count() or 4)
    with make_renderer(profile, bounces, ek, es, math_mode=api.MATH_FAST) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
    mx, rmse = stats(g, o)
    assert np.isfinite(g).all()
    assert rmse <= tol, (mx, rmse)
    assert abs(float(g.mean()) - float(o.mean())) <= 1e-3 * float(o.mean())


def test_frame_chunking_and_tiling_invariance(oracle):
    """16 frames in one launch == 5 + 11 frames in two launches == any tile grid (bit-exact)."""
    W, H = 192, 96
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(W, H, 2, 4)
        r.render_frames(16)
        a = api.detile(r.download_target(), W, H, 2, 4)
        r.resize(W, H, 6, 3)
        r.render_frames(5)
        r.render_frames(11)
        b = api.detile(r.download_target(), W, H, 6, 3)
    assert np.array_equal(a, b)


def test_render_host_is_the_reference_call(oracle):
    """b200pt_render_host == DemofoxRenderOptV4 called nframes times on the caller's host buffer."""
    W, H, ntx, nty = 128, 72, 4, 6
    env = oracle.synthetic_env(128, 64)
    o, _ = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, 8, 5, env=env, env_kind=1, env_sampler=2)
    buf = np.zeros(W * H * 3, dtype=np.float32)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8) as r:
        r.render_host(buf, W, H, ntx, nty, 2, env=env)
        r.render_host(buf, W, H, ntx, nty, 3, env=env)
    assert np.array_equal(buf, o)


def test_ldr_resolve_bit_exact(oracle):
    g = load_golden("v4_ldr")
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, output_to_screen=True) as r:
        r.set_env(oracle.synthetic_env(128, 64))
        r.resize(128, 72, 4, 6)
        r.upload_target(g["buffer"])
        assert np.array_equal(r.resolve_ldr(api.LDR_FILE_RGBA), g["ldr"])
        assert np.array_equal(r.resolve_ldr(api.LDR_SCREEN_BGRA), oracle.resolve_ldr(g["buffer"], 128, 72, 4, 6, mode=1))
        f0 = r.frame_counter
        r.resolve_ldr(api.LDR_FILE_RGBA, bump_frame_counter=True)  # CopyOutputToFile bumps iFrame, v4.cpp:1741
        assert r.frame_counter == f0 + 1
        # OUTPUT_TO_SCREEN path of render_host
        buf = np.zeros(128 * 72 * 3, dtype=np.float32)
        scr = np.zeros((72, 128), dtype=np.uint32)
        r.frame_counter = 0
        r.render_host(buf, 128, 72, 4, 6, 6, screen=scr)
        assert np.array_equal(buf, g["buffer"])
        assert np.array_equal(scr, oracle.resolve_ldr(buf, 128, 72, 4, 6, mode=1))


SWITCH_CASES = [(1, 0, api.ENV_EQUIRECT, api.SAMPLER_RANDOM), (0, 1, api.ENV_CUBEMAP, api.SAMPLER_BILINEAR),
                (1, 1, api.ENV_CUBEMAP, api.SAMPLER_RANDOM), (1, 1, api.ENV_NONE, api.SAMPLER_RANDOM)]


@pytest.mark.parametrize("exact_exp,sincos,ek,es", SWITCH_CASES)
@pytest.mark.parametrize("sched,generic", [(api.SCHED_LANE, False), (api.SCHED_LANE, True), (api.SCHED_SORTED, False)],
                         ids=["static-kernel", "generic-kernel", "sorted"])
def test_non_default_switches_bit_exact_vs_oracle(oracle, exact_exp, sincos, ek, es, sched, generic):
    """USE_FAST_APPROXIMATE_EXP 0 / USE_UNIT_VECTOR_REJECTION_SAMPLING 0 (global_preprocessor_flags.h:64-65): buffer, RNG
    states (2 + 2 draws per bounce instead of 3 + 3) and counters against the oracle, which is pinned on reference builds
    with those switches flipped (tests/golden v4_exact_exp, v4_sincos_unit_vectors, v4_all_exact).  Three kernels carry them:
    the scene-specialised ones with the switches compiled in (pt_kernels_*_v4sw.cu), the generic per-lane kernel and the
    generic CTA-sorted kernel (switches read at run time)"""
    import ctypes
    W, H, ntx, nty, frames, bounces = 256, 192, 4, 6, 10, 8
    env = None if ek == api.ENV_NONE else oracle.synthetic_env(*((64, 384) if ek == api.ENV_CUBEMAP else (256, 128)))
    flags = (oracle.V4_EXACT_EXP if exact_exp else 0) | (oracle.V4_SINCOS_UNIT_VECTORS if sincos else 0)
    o, oc = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, bounces, frames, env=env, env_kind=ek, env_sampler=es, v4_flags=flags)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=bounces, env_kind=ek, env_sampler=es, scheduler=sched,
                      exact_exp=exact_exp, sincos_unit_vectors=sincos, generic_scene_tables=generic) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(4)
        r.render_frames(frames - 4)
        g = r.download_target()
        assert np.array_equal(g, o), "max abs %g rmse %g" % stats(g, o)
        c = r.counters()
        assert (c["paths"], c["segments"], c["escapes"]) == (oc["paths"], oc["segments"], oc["escapes"])
        rs = r.rng_state()
        p, keep = oracle.make_params(oracle.PROFILE_V4, W, H, ntx, nty, bounces, env, ek, es, v4_flags=flags)
        rng = np.random.default_rng(11)
        for _ in range(200):
            x, y = int(rng.integers(W)), int(rng.integers(H))
            assert int(rs[y, x]) == oracle.lib().oracle_final_rng_state(ctypes.byref(p), x, y, frames)


def test_non_default_switches_are_v4_only_and_fast_math_stays_close(oracle):
    for prof in (api.PROFILE_V2, api.PROFILE_SIMT_TEXTURED, api.PROFILE_V3_REDO):
        with pytest.raises(api.B200PTError):
            api.Renderer(profile=prof, exact_exp=True)
        with pytest.raises(api.B200PTError):
            api.Renderer(profile=prof, sincos_unit_vectors=True)
    # fast math with the switches on: same image within Monte-Carlo noise of the parity render (both 64 spp)
    W, H, ntx, nty, frames = 256, 192, 4, 6, 64
    env = oracle.synthetic_env(256, 128)
    imgs = []
    for mm in (api.MATH_PARITY, api.MATH_FAST):
        with api.Renderer(profile=api.PROFILE_OPT_V4, math_mode=mm, exact_exp=True, sincos_unit_vectors=True) as r:
            r.set_env(env)
            r.resize(W, H, ntx, nty)
            r.render_frames(frames)
            imgs.append(r.download_target().astype(np.float64))
    assert np.isfinite(imgs[1]).all()
    assert abs(imgs[0].mean() - imgs[1].mean()) < 0.01 * imgs[0].mean()


def test_exact_aces_tonemap_bit_exact(oracle):
    """USE_FAST_APPROXIMATE_ACES_TONEMAP 0 (global_preprocessor_flags.h:63): resolve, fused screen output and present"""
    g = load_golden("v4_ldr_exact_aces")
    W, H, ntx, nty = 128, 72, 4, 6
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, output_to_screen=True, exact_exp=True, sincos_unit_vectors=True,
                      exact_aces_tonemap=True) as r:
        r.set_env(oracle.synthetic_env(128, 64))
        r.resize(W, H, ntx, nty)
        r.upload_target(g["buffer"])
        assert np.array_equal(r.resolve_ldr(api.LDR_FILE_RGBA), g["ldr"])  # the reference build's CopyOutputToFile
        assert np.array_equal(r.resolve_ldr(api.LDR_SCREEN_BGRA), oracle.resolve_ldr(g["buffer"], W, H, ntx, nty, mode=3))
        buf = np.zeros(W * H * 3, dtype=np.float32)
        scr = np.zeros((H, W), dtype=np.uint32)
        r.frame_counter = 0
        r.render_host(buf, W, H, ntx, nty, 6, screen=scr)  # fused tone map in the render kernel's tail
        assert np.array_equal(buf, g["buffer"])
        assert np.array_equal(scr, oracle.resolve_ldr(buf, W, H, ntx, nty, mode=3))
    # per call (mode | LDR_EXACT_ACES) on a context with the default curve, on values where the two curves differ at 8 bits
    rng = np.random.default_rng(1)
    W2, H2 = 4096, 64
    vals = (rng.random(W2 * H2 * 3, dtype=np.float32) ** 3 * 4).astype(np.float32)
    with api.Renderer(profile=api.PROFILE_V2) as r:
        r.resize(W2, H2, 1, 1)
        r.upload_target(vals)
        fast, exact = r.resolve_ldr(api.LDR_FILE_RGBA), r.resolve_ldr(api.LDR_FILE_RGBA | api.LDR_EXACT_ACES)
        assert np.array_equal(fast, oracle.resolve_ldr(vals, W2, H2, 1, 1, mode=0))
        assert np.array_equal(exact, oracle.resolve_ldr(vals, W2, H2, 1, 1, mode=2))
        assert not np.array_equal(fast, exact)


@pytest.mark.parametrize("aces,gamma", [(False, True), (True, True)])
def test_exact_gamma_tonemap_bit_exact(oracle, aces, gamma):
    """USE_FAST_APPROXIMATE_GAMMA 0 (global_preprocessor_flags.h:62), alone and with the exact ACES curve: resolve, OUTPUT_TO_SCREEN
    (the resolve kernel runs behind the render kernel), the present ring, and per call on a default context"""
    g = load_golden("v4_ldr_exact_aces_gamma" if aces else "v4_ldr_exact_gamma")
    W, H, ntx, nty = 128, 72, 4, 6
    mode = (2 if aces else 0) | 4
    env = oracle.synthetic_env(128, 64)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, output_to_screen=True, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM,
                      exact_aces_tonemap=aces, exact_gamma=gamma) as r:
        r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.upload_target(g["buffer"])
        assert np.array_equal(r.resolve_ldr(api.LDR_FILE_RGBA), g["ldr"])  # the reference build's CopyOutputToFile
        assert np.array_equal(r.resolve_ldr(api.LDR_SCREEN_BGRA), oracle.resolve_ldr(g["buffer"], W, H, ntx, nty, mode=mode | 1))
        buf = np.zeros(W * H * 3, dtype=np.float32)
        scr = np.zeros((H, W), dtype=np.uint32)
        r.frame_counter = 0
        r.render_host(buf, W, H, ntx, nty, 6, screen=scr)
        assert np.array_equal(buf, g["buffer"])
        assert np.array_equal(scr, oracle.resolve_ldr(buf, W, H, ntx, nty, mode=mode | 1))
        # progressive present: every presented frame is the tone map of the buffer after that frame
        r.reset()
        for k in range(1, 4):
            r.present_submit(1)
            frame = r.present_acquire()[0].copy()
            o, _ = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, 8, k, env=env, env_kind=1, env_sampler=2)
            assert np.array_equal(frame.reshape(H, W), oracle.resolve_ldr(o, W, H, ntx, nty, mode=mode | 1).reshape(H, W))
    rng = np.random.default_rng(2)
    W2, H2 = 2048, 64
    vals = (rng.random(W2 * H2 * 3, dtype=np.float32) ** 3 * 4).astype(np.float32)
    with api.Renderer(profile=api.PROFILE_V2) as r:
        r.resize(W2, H2, 1, 1)
        r.upload_target(vals)
        for m in (api.LDR_EXACT_GAMMA, api.LDR_EXACT_GAMMA | api.LDR_EXACT_ACES, api.LDR_SCREEN_BGRA | api.LDR_EXACT_GAMMA):
            assert np.array_equal(r.resolve_ldr(m), oracle.resolve_ldr(vals, W2, H2, 1, 1, mode=m))
        with pytest.raises(api.B200PTError):
            r.resolve_ldr(8)


def test_sum_mode_matches_running_average(oracle):
    """ACCUM_SUM + finalize (the spp-shard epilogue) vs the sequential running average: different
    rounding order only; tolerance 2e-6 relative to the image scale (documented in DESIGN.md)."""
    W, H, frames = 128, 96, 32
    o, _ = oracle.render(oracle.PROFILE_V2, W, H, 2, 4, 8, frames)
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, accum_mode=api.ACCUM_SUM) as r:
        r.resize(W, H, 2, 4)
        r.render_frames(10)
        r.render_frames(22)
        r.finalize_sum(frames)
        g = r.download_target()
    assert np.allclose(g, o, rtol=2e-6, atol=2e-6)


def test_edge_cases(oracle):
    # smallest legal image (one SoA8 group), ragged warp (groups not a multiple of 4), bounces = 0
    for (W, H, ntx, nty, b) in [(8, 1, 1, 1, 4), (24, 5, 3, 5, 4), (40, 3, 1, 3, 0), (8, 8, 1, 8, 16)]:
        o, _ = oracle.render(oracle.PROFILE_V2, W, H, ntx, nty, b, 3)
        with api.Renderer(profile=api.PROFILE_V2, num_bounces=b) as r:
            r.resize(W, H, ntx, nty)
            r.render_frames(0)
            r.render_frames(3)
            assert np.array_equal(r.download_target(), o)


def test_error_behaviour():
    with api.Renderer(profile=api.PROFILE_SIMT_TEXTURED) as r:
        with pytest.raises(api.B200PTError, match="invalid tiling"):
            r.resize(60, 32, 2, 2)  # tile width 30: CheckValidSettings would __debugbreak
        with pytest.raises(api.B200PTError):
            r.render_frames(1)  # before resize
        r.resize(64, 32, 2, 2)
        with pytest.raises(api.B200PTError, match="env"):
            r.render_frames(1)  # env-sampling profile without env


def test_full_size_properties():
    """BASELINE config 2 size (1920x1080, 8 bounces): determinism and chunk invariance, bit-exact;
    path accounting."""
    W, H = 1920, 1080
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(W, H, 10, 15)
        r.render_frames(8)
        a = r.download_target()
        c = r.counters()
        assert c["paths"] == W * H * 8 and c["paths"] <= c["segments"] <= 9 * c["paths"]
        r.reset()
        r.render_frames(3)
        r.render_frames(5)
        b = r.download_target()
    assert np.array_equal(a, b) and np.isfinite(a).all() and a.min() >= 0


@pytest.mark.parametrize("profile,envshape,ek,es", [(api.PROFILE_V2, None, None, None), (api.PROFILE_OPT_V4, (128, 64), 1, 2),
                                                    (api.PROFILE_SIMT_TEXTURED, (128, 64), None, None),
                                                    (api.PROFILE_V3_REDO, (128, 64), None, None),
                                                    (api.PROFILE_V3_REDO_SCENE0, (128, 64), None, None)],
                         ids=["v2", "v4", "simt_textured", "v3_redo", "v3_redo_scene0"])
def test_camera_culling_changes_nothing(oracle, profile, envshape, ek, es):
    """Skipping the scene trace for pixels whose jitter footprint misses every primitive's screen
    bounds must not change a single bit (buffer, RNG states, counters)."""
    W, H, ntx, nty, frames = 384, 216, 4, 6, 10
    env = oracle.synthetic_env(*envshape) if envshape else None
    res = []
    for off in (False, True):
        kw = dict(env_kind=ek, env_sampler=es) if profile == api.PROFILE_OPT_V4 else {}
        with api.Renderer(profile=profile, num_bounces=8, disable_camera_culling=off, **kw) as r:
            if env is not None:
                r.set_env(env)
            r.resize(W, H, ntx, nty)
            r.render_frames(frames)
            c = r.counters()
            res.append((r.download_target(), r.rng_state(), c["segments"], c["escapes"]))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert res[0][2:] == res[1][2:]


@pytest.mark.parametrize("sched", [api.SCHED_LANE, api.SCHED_SORTED], ids=["lane", "sorted"])
def test_item_pull_order_changes_nothing(oracle, sched):
    """scene-first / sky-last pull order of the work items (a table built on the device per launch geometry): the same
    bits as buffer order -- image, RNG states, counters -- for whole images, tile ranges and strip-shaped items"""
    for (W, H, ntx, nty) in ((640, 360, 4, 5), (648, 363, 3, 3)):  # tile height 72 (8x4 block items) / 121 (32x1 strips)
        res = []
        for off in (False, True):
            with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, disable_item_order=off, scheduler=sched) as r:
                r.resize(W, H, ntx, nty)
                for n in (2, 9, 4):          # built by the first launch of >= 8 frames on a geometry, reused afterwards
                    r.render_frames(n)
                r.set_tile_row_range(1, nty - 1)
                for n in (8, 2, 2):
                    r.render_frames(n)
                c = r.counters()
                res.append((r.download_target(), r.rng_state(), (c["paths"], c["segments"], c["escapes"], c["culled_segments"])))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    o, _ = oracle.render(oracle.PROFILE_V2, 640, 360, 4, 5, 8, 6)
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, scheduler=sched) as r:
        r.resize(640, 360, 4, 5)
        r.render_frames(1); r.render_frames(2); r.render_frames(3)
        assert np.array_equal(r.download_target(), o)
        o12, _ = oracle.render(oracle.PROFILE_V2, 640, 360, 4, 5, 8, 12)
        r.reset()
        r.render_frames(12)   # one launch, ordered pull
        assert np.array_equal(r.download_target(), o12)


def test_tile_row_bands_assemble_to_the_full_render(oracle):
    """tile-shard building block: rendering tile rows band by band == the full render, bit for bit."""
    W, H, ntx, nty, frames = 192, 120, 3, 5, 9
    o, _ = oracle.render(oracle.PROFILE_V2, W, H, ntx, nty, 8, frames)
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(W, H, ntx, nty)
        for first, count in ((3, 2), (0, 1), (1, 2)):
            r.frame_counter = 0
            r.set_tile_row_range(first, count)
            r.render_frames(frames)
        assert np.array_equal(r.download_target(), o)
        with pytest.raises(api.B200PTError):
            r.set_tile_row_range(4, 2)
        # per-tile scheduling, like the reference's work-queue entries (one FlatTileIndex per entry)
        r.reset()
        for flat in np.random.default_rng(0).permutation(ntx * nty):
            r.frame_counter = 0
            r.set_tile_range(int(flat), 1)
            r.render_frames(frames)
        assert np.array_equal(r.download_target(), o)


@pytest.mark.parametrize("profile", [api.PROFILE_V2, api.PROFILE_SIMT_TEXTURED, api.PROFILE_V3_REDO, api.PROFILE_OPT_V4, api.PROFILE_V3_REDO_SCENE0],
                         ids=["v2", "simt_textured", "v3_redo", "opt_v4", "v3_redo_scene0"])
def test_static_scene_specialisation_changes_nothing(oracle, profile):
    """Built-in scene as compile-time knowledge (default: quad vertices / sphere centres as immediates, zero
    components of the v4 quad tables dropped, unchecked reciprocals for the built-in materials) vs the generic
    kernel that reads everything from the scene table: identical bits."""
    W, H, ntx, nty, frames = 256, 160, 4, 5, 10
    env = oracle.synthetic_env(128, 64) if profile != api.PROFILE_V2 else None
    kw = dict(env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM) if profile == api.PROFILE_OPT_V4 else {}
    res = []
    for generic in (False, True):
        with api.Renderer(profile=profile, num_bounces=8, generic_scene_tables=generic, **kw) as r:
            if env is not None:
                r.set_env(env)
            r.resize(W, H, ntx, nty)
            r.render_frames(frames)
            res.append((r.download_target(), r.rng_state()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_present_ring_frames_are_the_reference_screen_frames(oracle):
    """Progressive present path: frame k of the ring == OutputToScreen of the buffer after k render
    calls (fused tone map in the render kernel, asynchronous double-buffered copy)."""
    W, H, ntx, nty = 128, 72, 4, 6
    cube = oracle.synthetic_env(32, 192)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM,
                      output_to_screen=True) as r:
        r.set_env(cube)
        r.resize(W, H, ntx, nty)
        with pytest.raises(api.B200PTError):
            r.present_acquire()  # nothing in flight
        frames = []
        r.present_submit(1)
        for k in range(2, 7):
            r.present_submit(1)                 # frame k renders while frame k-1 is copied
            frames.append(r.present_acquire())
        with pytest.raises(api.B200PTError):
            r.present_submit(1); r.present_submit(1); r.present_submit(1)  # ring holds two frames
        frames.append(r.present_acquire())
    buf = np.zeros(W * H * 3, dtype=np.float32)
    for k, (img, iframe) in enumerate(frames[:6], start=1):
        buf, _ = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, 8, 1, first_frame=k, env=cube, env_kind=2, env_sampler=2,
                               target=buf)
        assert iframe == k
        assert np.array_equal(img, oracle.resolve_ldr(buf, W, H, ntx, nty, mode=1)), k


def test_present_blocking_is_the_reference_screen_frame(oracle):
    """b200pt_present_blocking: render + fused tone map + band-pipelined copy == OutputToScreen of the oracle's buffer,
    for any band count, and the accumulation buffer / frame counter advance exactly once"""
    W, H, ntx, nty = 192, 120, 3, 5
    cube = oracle.synthetic_env(32, 192)
    buf = np.zeros(W * H * 3, dtype=np.float32)
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_CUBEMAP, env_sampler=api.SAMPLER_RANDOM) as r:
        r.set_env(cube)
        r.resize(W, H, ntx, nty)
        import torch
        frame = torch.zeros((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)  # page-locked + device-mapped
        k = 0
        for bands, n in ((0, 1), (1, 2), (3, 1), (5, 3), (64, 1), (-1, 1), (-1, 2)):
            r.present_blocking(frame, nframes=n, bands=bands)
            buf, _ = oracle.render(oracle.PROFILE_V4, W, H, ntx, nty, 8, n, first_frame=k + 1, env=cube, env_kind=2, env_sampler=2,
                                   target=buf)
            k += n
            assert r.frame_counter == k
            assert np.array_equal(frame, oracle.resolve_ldr(buf, W, H, ntx, nty, mode=1)), (bands, n)
            assert np.array_equal(r.download_target(), buf)
        assert r.counters()["paths"] == W * H * k


def test_present_ring_zero_copy_view_survives_the_next_submits():
    """The pointer b200pt_present_acquire hands out stays valid until the NEXT acquire, whatever is submitted
    meanwhile (three host slots): a zero-copy consumer holds frame k while frames k+1 and k+2 render and copy."""
    import time
    W, H, ntx, nty = 256, 128, 4, 4
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8, output_to_screen=True) as r:
        r.resize(W, H, ntx, nty)
        r.present_submit(1)
        for k in range(1, 8):                                  # every phase of the 3-slot rotation
            view, iframe = r.present_acquire(copy=False)       # frame k, a view of pinned memory
            assert iframe == k
            snapshot = view.copy()
            r.present_submit(1)                                # frames k+1 and k+2 go through the ring
            r.present_submit(1)
            r.synchronize()
            time.sleep(0.02)                                   # let both asynchronous copies land
            assert np.array_equal(view, snapshot), k           # ... and frame k is still frame k
            nxt, fnext = r.present_acquire()                   # k+1 (the view of k is dead from here on)
            assert fnext == k + 1 and not np.array_equal(nxt, snapshot)
            # leave exactly one frame (k+2) in flight for the next round: skip it and submit the following one
            skipped, fs = r.present_acquire()
            assert fs == k + 2
            r.frame_counter = k
            r.present_submit(1)


def test_full_size_bit_exact_vs_oracle(oracle):
    """BASELINE config 2 geometry (1920x1080, tiles 10x15, 8 bounces) at a bounded spp: the whole f32
    buffer, the RNG states and the counters equal the oracle's, bit for bit (the oracle takes ~10 s)."""
    W, H, ntx, nty, frames = 1920, 1080, 10, 15, 6
    o, oc = oracle.render(oracle.PROFILE_V2, W, H, ntx, nty, 8, frames)
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
        c = r.counters()
    assert np.array_equal(g, o)
    assert (c["paths"], c["segments"], c["escapes"]) == (oc["paths"], oc["segments"], oc["escapes"])


@pytest.mark.parametrize("profile,W,H,ntx,nty,bounces,frames,envshape,ek,es", [
    (0, 4096, 4096, 16, 32, 16, 2, None, 0, 0),          # BASELINE config 5 geometry class: 16 bounces, large square image
    (2, 3840, 2160, 10, 15, 8, 2, (2048, 1024), 1, 2),   # config 3 geometry: 4K, v4, 2048x1024 equirect, random-jitter sampler
    (1, 3840, 2160, 10, 15, 8, 1, (2048, 1024), 1, 0),   # config 3: simt_textured, point sampler
], ids=["v2_4096sq_16b", "v4_4k_equirect", "simt_4k"])
def test_large_images_bit_exact_vs_oracle(oracle, profile, W, H, ntx, nty, bounces, frames, envshape, ek, es):
    """the BASELINE configurations' image sizes at a bounded frame count: whole buffers and counters equal the oracle's"""
    env = oracle.synthetic_env(*envshape) if envshape else None
    o, oc = oracle.render(profile, W, H, ntx, nty, bounces, frames, env=env, env_kind=ek, env_sampler=es)
    with make_renderer(profile, bounces, ek, es) as r:
        if env is not None:
            r.set_env(env)
        r.resize(W, H, ntx, nty)
        r.render_frames(frames)
        g = r.download_target()
        c = r.counters()
    assert np.array_equal(g, o)
    assert (c["paths"], c["segments"], c["escapes"]) == (oc["paths"], oc["segments"], oc["escapes"])


def test_randomised_configurations_bit_exact(oracle):
    """Seeded sweep over resolution, tiling, bounce budget, frame offset, profile and a random
    pre-existing accumulation state: GPU == oracle, bit for bit."""
    rng = np.random.default_rng(20260)
    envs = {1: oracle.synthetic_env(96, 48), 2: oracle.synthetic_env(16, 96)}
    for trial in range(24):
        profile = int(rng.choice([0, 1, 2, 3]))
        ntx, nty = int(rng.integers(1, 5)), int(rng.integers(1, 6))
        W, H = 8 * ntx * int(rng.integers(1, 7)), nty * int(rng.integers(1, 25))
        bounces, frames, start = int(rng.integers(0, 13)), int(rng.integers(1, 6)), int(rng.integers(0, 2000))
        ek, es = 0, 0
        if profile == 1:
            ek, es = 1, 0
        elif profile == 3:
            ek, es = 1, 1
        elif profile == 2:
            ek = int(rng.choice([0, 1, 2]))
            es = int(rng.choice([1, 2])) if ek else 0
        env = envs.get(ek)
        state = (rng.random(W * H * 3) * 2).astype(np.float32)
        o, oc = oracle.render(profile, W, H, ntx, nty, bounces, frames, first_frame=start + 1, env=env, env_kind=ek,
                              env_sampler=es, target=state)
        with make_renderer(profile, bounces, ek, es if ek else api.SAMPLER_RANDOM) as r:
            if env is not None:
                r.set_env(env)
            r.resize(W, H, ntx, nty)
            r.upload_target(state)
            r.frame_counter = start
            r.render_frames(frames)
            g = r.download_target()
            c = r.counters()
        assert np.array_equal(g, o), (trial, profile, W, H, ntx, nty, bounces, frames, start, ek, es)
        assert (c["segments"], c["escapes"]) == (oc["segments"], oc["escapes"])
