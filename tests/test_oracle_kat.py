"""Known-answer tests for the oracle's RNG, seeding, buffer layout and constants.
KAT values: SURVEY.md section 8c (derived from the reference code via the shim, cross-checked
with an independent model): mathutils.h:8-26, demofox_path_tracing_optimization_v4.cpp:1096-1101."""
import ctypes

import numpy as np


def test_wang_hash_kat(oracle):
    s = ctypes.c_uint32(12345)
    assert oracle.lib().oracle_wang_hash(ctypes.byref(s)) == 232713235
    s = ctypes.c_uint32(12345)
    f = oracle.lib().oracle_random01(ctypes.byref(s))
    assert abs(f - 0.108366) < 1e-6
    assert f == np.float32(np.int32(232713235 & 0x7FFFFFFF)) / np.float32(2147483648.0)


def test_seed_kat(oracle):
    # row y' = 5, frame 1, x = 0..7
    seeds = [oracle.lib().oracle_seed(x, 5, 1) for x in range(8)]
    assert seeds == [73085, 75057, 77031, 79003, 80977, 82949, 84923, 86895]


def test_stream_kat(oracle):
    L = oracle.lib()
    for x, states, floats in [
        (0, [2464760169, 3189164220, 3159020168, 1430376506], [0.147743389, 0.485070318, 0.471033394, 0.666070938]),
        (3, [4013349876, 1777298695, 589394112, 2823360539], [0.868861675, 0.827619195, 0.274458021, 0.314729691]),
    ]:
        s = ctypes.c_uint32(L.oracle_seed(x, 5, 1))
        for st, fl in zip(states, floats):
            f = L.oracle_random01(ctypes.byref(s))
            assert s.value == st
            assert abs(f - fl) < 1e-8


def test_random01_can_reach_one(oracle):
    # (hash & 0x7FFFFFFF) close to 2^31 rounds up to 2^31 in binary32: 1.0 is reachable
    assert np.float32(np.int32(0x7FFFFFFF)) / np.float32(2147483648.0) == np.float32(1.0)


def test_camera_distance_is_one(oracle):
    assert oracle.lib().oracle_camera_distance() == 1.0


def test_buffer_index_matches_detile(oracle):
    W, H, ntx, nty = 64, 24, 4, 3
    buf = np.arange(W * H * 3, dtype=np.float32)
    img = oracle.detile(buf, W, H, ntx, nty)
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, y, c = int(rng.integers(W)), int(rng.integers(H)), int(rng.integers(3))
        assert img[y, x, c] == buf[L.oracle_buffer_index(W, H, ntx, nty, x, y, c)]
    assert np.array_equal(oracle.tile(img, ntx, nty), buf)


def test_buffer_index_formula(oracle):
    # SURVEY.md section 8a a11
    W, H, ntx, nty = 1920, 1080, 10, 15
    TW, TH = W // ntx, H // nty
    L = oracle.lib()
    for (x, y, c) in [(0, 0, 0), (7, 0, 2), (8, 0, 0), (191, 71, 1), (192, 0, 0), (0, 72, 0), (1919, 1079, 2), (1003, 517, 1)]:
        tx, ty, lx, ly = x // TW, y // TH, x % TW, y % TH
        want = ty * TH * W * 3 + tx * TW * TH * 3 + (ly * TW + (lx & ~7)) * 3 + c * 8 + (lx & 7)
        assert L.oracle_buffer_index(W, H, ntx, nty, x, y, c) == want


def test_invalid_tiling_rejected(oracle):
    import pytest
    with pytest.raises(ValueError):
        oracle.render(oracle.PROFILE_V2, 60, 32, 2, 2, 4, 1)  # tile width 30 not a multiple of 8
    with pytest.raises(ValueError):
        oracle.render(oracle.PROFILE_SIMT_TEXTURED, 64, 32, 2, 2, 4, 1)  # env missing
