"""Camera-ray culling (b200pt_compute_cull_rects): every pixel the engine treats as 'cannot hit the
scene' must, in the reference's own arithmetic (oracle), escape on its camera ray for every jitter
-- checked exhaustively per pixel over many frames and several resolutions -- and the rectangles must
not be uselessly loose."""
import numpy as np
import pytest

from cpuperformanceraytracer_b200 import api


def sure_miss_mask(profile, W, H):
    rects = api.cull_rects(profile, W, H)
    assert rects is not None
    xs = np.arange(W, dtype=np.float32)[None, :]
    yf = (H - 1 - np.arange(H, dtype=np.float32))[:, None]  # flipped row index = fragCoord.y
    hit = np.zeros((H, W), dtype=bool)
    for x0, y0, x1, y1 in rects:
        hit |= (xs + 0.5 >= x0) & (xs - 0.5 <= x1) & (yf + 0.5 >= y0) & (yf - 0.5 <= y1)
    return ~hit


@pytest.mark.parametrize("profile,oprofile,W,H,frames", [
    (api.PROFILE_V2, 0, 320, 180, 24), (api.PROFILE_V2, 0, 256, 256, 16), (api.PROFILE_V2, 0, 200, 64, 16),
    (api.PROFILE_OPT_V4, 2, 320, 180, 24), (api.PROFILE_OPT_V4, 2, 128, 256, 16), (api.PROFILE_V3_REDO, 3, 320, 180, 16),
    (api.PROFILE_V3_REDO_SCENE0, 4, 320, 180, 16), (api.PROFILE_V3_REDO_SCENE0, 4, 96, 160, 12),
])
def test_culled_pixels_always_escape_in_the_oracle(oracle, profile, oprofile, W, H, frames):
    mask = sure_miss_mask(profile, W, H)
    env = oracle.synthetic_env(64, 32) if oprofile >= 3 else None
    seg = oracle.max_segments(oprofile, W, H, 8, frames, env=env, env_kind=1 if oprofile >= 3 else 0, env_sampler=1 if oprofile >= 3 else 0)
    assert (seg[mask] == 1).all(), "a culled pixel hit geometry in the reference arithmetic"
    # usefulness: most pixels that always escape are culled (the bounds are not absurdly loose)
    always_escape = seg == 1
    assert mask.sum() >= 0.5 * always_escape.sum()
    # safety margin: no culled pixel within 1 pixel of a pixel that ever hits
    hits = ~always_escape
    grown = hits.copy()
    grown[1:, :] |= hits[:-1, :]; grown[:-1, :] |= hits[1:, :]; grown[:, 1:] |= hits[:, :-1]; grown[:, :-1] |= hits[:, 1:]
    assert not (grown & mask).any()


def test_simt_textured_uses_the_cornell_bounds():
    a, b = api.cull_rects(api.PROFILE_V2, 640, 360), api.cull_rects(api.PROFILE_SIMT_TEXTURED, 640, 360)
    assert np.array_equal(a, b) and len(a) == 1
