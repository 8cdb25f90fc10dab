"""The CUDA path (parity mode) against the UNTOUCHED reference build (`asis`: hardware rcpps / rsqrtps as
mathlib.h:417,444 use them, glibc libm), statistically -- SURVEY.md section 7 hard part 1(a).

The bit-exact anchor is the `exact` reference build (tests/test_oracle_vs_ref.py, tests/golden).  The `asis` build
differs from it at approximation level (~2^-12 relative in every reciprocal): where the renderer uses none of them
in its path loop (P_v2: exact divisions, sqrt, libm sin/cos) the images agree to RMSE ~2e-4 at 1024 spp; where it
does (P_v4: rcp / rsqrt everywhere) individual paths decorrelate and only image statistics agree.  Measured on B200
next to a 16-core host (scripts/compare_asis.py, profiles/r02_b_compare_asis.jsonl, 512x512, 1024 spp):
    P_v2   RMSE 1.9e-4, max-abs 0.023, relative mean-brightness shift -1.4e-6   (Monte-Carlo noise floor 0.11)
    P_v4   RMSE 0.071,  max-abs 2.7,   relative mean-brightness shift -1.2e-4   (noise floor 0.14)
Tolerances below: ~3x those values at the test's smaller size."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cpuperformanceraytracer_b200 import api  # noqa: E402


def _rel_shift(g, ref):
    return (float(ref.astype(np.float64).mean()) - float(g.astype(np.float64).mean())) / float(g.astype(np.float64).mean())


def test_v2_matches_the_untouched_reference_build(oracle):
    if not oracle.ref_binary("ref_v2_asis"):
        pytest.skip("oracle/_ref/ref_v2_asis not built (needs /root/reference at build time)")
    W = H = 256
    spp = 512
    ref = oracle.run_ref("ref_v2_asis", W, H, 2, 4, spp, bounces=8, threads=8)["buffer"]
    with api.Renderer(profile=api.PROFILE_V2, num_bounces=8) as r:
        r.resize(W, H, 2, 4)
        r.render_frames(spp)
        g = r.download_target()
    d = g.astype(np.float64) - ref
    rmse, mx = float(np.sqrt((d * d).mean())), float(np.abs(d).max())
    assert rmse <= 1e-3, (rmse, mx)            # north_star: image RMSE vs reference <= 1e-3 at matched seeds / spp
    assert mx <= 0.1, (rmse, mx)               # per-channel max abs error
    assert abs(_rel_shift(g, ref)) <= 2e-5


def test_v4_statistics_match_the_untouched_reference_build(oracle):
    if not oracle.ref_binary("ref_v4_equirect_random_asis"):
        pytest.skip("oracle/_ref/ref_v4_equirect_random_asis not built (needs /root/reference at build time)")
    W = H = 256
    spp = 512
    env = oracle.synthetic_env(512, 256)
    ref = oracle.run_ref("ref_v4_equirect_random_asis", W, H, 4, 4, spp, bounces=8, env=env, threads=os.cpu_count() or 8)["buffer"]
    with api.Renderer(profile=api.PROFILE_OPT_V4, num_bounces=8, env_kind=api.ENV_EQUIRECT, env_sampler=api.SAMPLER_RANDOM) as r:
        r.set_env(env)
        r.resize(W, H, 4, 4)
        r.render_frames(spp)
        g = r.download_target()
        r.reset()
        r.render_frames(spp // 4)
        quarter = r.download_target()
    d = g.astype(np.float64) - ref
    rmse = float(np.sqrt((d * d).mean()))
    noise = float(np.sqrt(((quarter.astype(np.float64) - g) ** 2).mean()))  # what 4x fewer samples of the SAME renderer differ by
    assert np.isfinite(ref).all() and np.isfinite(g).all()
    assert rmse <= noise, (rmse, noise)        # decorrelated paths, same estimator: below the Monte-Carlo noise of spp / 4
    assert abs(_rel_shift(g, ref)) <= 1e-3     # hardware rcpps / rsqrtps bias the image by ~1e-4 relative (-8e-4 at most)
