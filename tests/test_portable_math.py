"""oracle/portable_math.h defines sin/cos/atan2/asin for the oracle (MSVC SVML is closed source:
'parity unpinned' at that boundary).  Check the definitions are faithful: <= 1 ulp from the
platform libm and equal to the correctly rounded float64-evaluated value on sampled inputs."""
import ctypes

import numpy as np


def _ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def test_sincos(oracle):
    rng = np.random.default_rng(1)
    # the hot path's domain: a = u * (2*pi) with u = k / 2^31 (v2.cpp:79-83)
    k = rng.integers(0, 2 ** 31, size=2_000_000, dtype=np.int64)
    u = k.astype(np.int32).astype(np.float32) / np.float32(2147483648.0)
    a = (u * np.float32(2.0 * np.float32(3.14159265359))).astype(np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    oracle.lib().oracle_pm_sincosf_array(_fp(a), _fp(s), _fp(c), a.size)
    s64, c64 = np.sin(a.astype(np.float64)).astype(np.float32), np.cos(a.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(s, s64).max() <= 1 and (s == s64).mean() > 0.9999
    assert _ulp_diff(c, c64).max() <= 1 and (c == c64).mean() > 0.9999
    assert _ulp_diff(s, np.sin(a)).max() <= 1 and _ulp_diff(c, np.cos(a)).max() <= 1
    # exact landmarks
    for ang, es, ec in [(0.0, 0.0, 1.0)]:
        ss, cc = ctypes.c_float(), ctypes.c_float()
        oracle.lib().oracle_pm_sincosf(ctypes.c_float(ang), ctypes.byref(ss), ctypes.byref(cc))
        assert ss.value == es and cc.value == ec


def test_atan2_asin(oracle):
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, 1_000_000).astype(np.float32)
    y = rng.uniform(-1, 1, 1_000_000).astype(np.float32)
    out = np.empty_like(x)
    oracle.lib().oracle_pm_atan2f_array(_fp(y), _fp(x), _fp(out), x.size)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(out, ref).max() <= 1 and (out == ref).mean() > 0.9999
    oracle.lib().oracle_pm_asinf_array(_fp(x), _fp(out), x.size)
    ref = np.arcsin(x.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(out, ref).max() <= 1 and (out == ref).mean() > 0.9999


def test_edge_cases(oracle):
    L = oracle.lib()
    L.oracle_pm_atan2f.restype = ctypes.c_float
    L.oracle_pm_atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    L.oracle_pm_asinf.restype = ctypes.c_float
    L.oracle_pm_asinf.argtypes = [ctypes.c_float]
    pi = np.float32(np.pi)
    assert L.oracle_pm_atan2f(0.0, -1.0) == pi
    assert L.oracle_pm_atan2f(-0.0, -1.0) == -pi
    assert L.oracle_pm_atan2f(0.0, 0.0) == 0.0
    assert L.oracle_pm_atan2f(1.0, 0.0) == np.float32(np.pi / 2)
    assert np.isnan(L.oracle_pm_asinf(1.0000001))  # normalised directions can overshoot 1 by an ulp
    assert L.oracle_pm_asinf(-1.0) == -np.float32(np.pi / 2)


def test_expf(oracle):
    L = oracle.lib()
    L.oracle_pm_expf_array.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_int64]
    rng = np.random.default_rng(4)
    x = np.concatenate([-rng.uniform(0, 60, 1_000_000), rng.uniform(0, 5, 100_000)]).astype(np.float32)  # absorption: -colour * dist
    out = np.empty_like(x)
    L.oracle_pm_expf_array(_fp(x), _fp(out), x.size)
    ref = np.exp(x.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(out, ref).max() <= 1 and (out == ref).mean() > 0.9999
    assert _ulp_diff(out, np.exp(x)).max() <= 2  # numpy's vectorised float32 exp is itself only faithful to ~2 ulp
    e = np.array([0.0, -200.0, 100.0, -103.0], dtype=np.float32)
    o = np.empty_like(e)
    L.oracle_pm_expf_array(_fp(e), _fp(o), e.size)
    assert o[0] == 1.0 and o[1] == 0.0 and np.isinf(o[2]) and o[3] == np.float32(np.exp(np.float64(-103.0)))


def test_powf(oracle):
    """pm_powf (the pow_ps of the non-fast gamma, v4.cpp:185): the float64 power rounded to binary32 on the gamma's domain and beyond"""
    L = oracle.lib()
    L.oracle_pm_powf_array.argtypes = [ctypes.POINTER(ctypes.c_float)] * 3 + [ctypes.c_int64]
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.random(2_000_000, dtype=np.float32), (rng.random(300_000, dtype=np.float32) * 100).astype(np.float32),
                        np.array([0.0031308, 1e-38, 1e-45, 3e38, 0.99999994, 1.0000001], np.float32)])
    for yv in (np.float32(1.0) / np.float32(2.4), np.float32(2.4), np.float32(-3.5), np.float32(0.5), np.float32(17.25)):
        y = np.full_like(x, yv)
        out = np.empty_like(x)
        L.oracle_pm_powf_array(_fp(x), _fp(y), _fp(out), x.size)
        with np.errstate(all="ignore"):
            ref = np.power(x.astype(np.float64), np.float64(yv)).astype(np.float32)
        assert _ulp_diff(out, ref).max() <= 1 and (out == ref).mean() > 0.99999, float(yv)
    # special values
    e = np.array([0, 0, 1, 5, np.inf, np.inf, -1, np.nan, 2, 0.5, 2, 0.5, 7], np.float32)
    p = np.array([2, -2, np.nan, 0, 2, -2, 0.5, 1, np.inf, np.inf, -np.inf, -np.inf, 1], np.float32)
    o = np.empty_like(e)
    L.oracle_pm_powf_array(_fp(e), _fp(p), _fp(o), e.size)
    want = [0, np.inf, 1, 1, np.inf, 0, np.nan, np.nan, np.inf, 0, 0, np.inf, 7]
    for got, w in zip(o, want):
        assert (np.isnan(got) and np.isnan(w)) or got == np.float32(w)
