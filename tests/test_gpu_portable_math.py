"""The parity kernels' sin/cos/atan2/asin/exp (csrc/pm_math.cuh) against the oracle's definitions
(oracle/portable_math.h), value by value through the C ABI, and the device-side comparison of the
short first-tier atan2/asin against the literal algorithm (exhaustive for asin)."""
import ctypes
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpuperformanceraytracer_b200 import api  # noqa: E402

pytestmark = pytest.mark.gpu


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _same(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def _inputs(seed, n):
    rng = np.random.default_rng(seed)
    unit = rng.uniform(-1, 1, n).astype(np.float32)
    bits = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)  # any bit pattern
    edge = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 0.49999997, 0.50000006, 0.99999994, 1.0000001, np.inf, -np.inf, np.nan,
                     1e-45, -1e-45, 1e-38, 3e38, 1e-30, 0.70710677, 0.25, 0.75], dtype=np.float32)
    return unit, bits, edge


def test_device_values_equal_the_oracle_definitions(oracle):
    L = oracle.lib()
    i64 = ctypes.c_int64
    unit, bits, edge = _inputs(11, 3_000_000)
    with api.Renderer(profile=api.PROFILE_V2) as r:
        # asin
        a = np.concatenate([unit, bits[:500_000], edge])
        want = np.empty_like(a)
        L.oracle_pm_asinf_array(_fp(a), _fp(want), i64(a.size))
        assert _same(r.eval_portable(api.FN_ASIN, a), want)
        # atan2: direction components, arbitrary patterns, all pairs of edge values
        ey, ex = np.meshgrid(edge, edge)
        y = np.concatenate([unit, bits[:500_000], ey.ravel(), unit[:1000] * 0])
        x = np.concatenate([unit[::-1], bits[500_000:1_000_000], ex.ravel(), unit[:1000]])
        y, x = np.ascontiguousarray(y, dtype=np.float32), np.ascontiguousarray(x, dtype=np.float32)
        want = np.empty_like(y)
        L.oracle_pm_atan2f_array(_fp(y), _fp(x), _fp(want), i64(y.size))
        assert _same(r.eval_portable(api.FN_ATAN2, y, x), want)
        # sin / cos on the hot path's domain u * 2 pi (v2.cpp:79-86) and beyond
        k = np.random.default_rng(5).integers(0, 2 ** 31, 2_000_000, dtype=np.int64)
        ang = (k.astype(np.int32).astype(np.float32) / np.float32(2147483648.0)) * np.float32(2.0 * np.float32(3.14159265359))
        ang = np.concatenate([ang.astype(np.float32), np.float32(100.0) * unit[:100_000]])
        s, c = np.empty_like(ang), np.empty_like(ang)
        L.oracle_pm_sincosf_array(_fp(ang), _fp(s), _fp(c), i64(ang.size))
        assert _same(r.eval_portable(api.FN_SIN, ang), s)
        assert _same(r.eval_portable(api.FN_COS, ang), c)
        # exp (v3_redo absorption)
        e = np.concatenate([-60 * np.abs(unit), 5 * unit[:100_000], np.array([0, -200, 100, -103, np.nan, np.inf, -np.inf], np.float32)]).astype(np.float32)
        want = np.empty_like(e)
        L.oracle_pm_expf_array(_fp(e), _fp(want), i64(e.size))
        assert _same(r.eval_portable(api.FN_EXP, e), want)
        # pow (the exact-gamma tone map, v4.cpp:185): the gamma's domain, other exponents, special values
        L.oracle_pm_powf_array.argtypes = [ctypes.POINTER(ctypes.c_float)] * 3 + [i64]
        px = np.concatenate([np.abs(unit), np.abs(bits[:500_000]), edge, np.abs(unit[:200_000]) * 100]).astype(np.float32)
        for yv in (np.float32(1.0) / np.float32(2.4), np.float32(2.4), np.float32(-3.5), np.float32(0.0), np.float32(np.inf), np.float32(np.nan)):
            py = np.full_like(px, yv)
            want = np.empty_like(px)
            L.oracle_pm_powf_array(_fp(px), _fp(py), _fp(want), i64(px.size))
            assert _same(r.eval_portable(api.FN_POW, px, py), want), float(yv)
        with np.errstate(invalid="ignore"):
            py = np.abs(bits[500_000:500_000 + px.size]) % np.float32(40.0)  # arbitrary exponents in [0, 40)
        py = np.where(np.isfinite(py), py, np.float32(1.5)).astype(np.float32)
        want = np.empty_like(px)
        L.oracle_pm_powf_array(_fp(px), _fp(py), _fp(want), i64(px.size))
        assert _same(r.eval_portable(api.FN_POW, px, py), want)


def test_first_tier_never_changes_a_result():
    with api.Renderer(profile=api.PROFILE_V2) as r:
        bad, literal = r.check_portable_tiers(api.FN_ASIN, 0, 2 ** 32)  # every binary32 input
        assert bad == 0
        # not served by the first tier: |v| >= 1 and NaN (2 * (2^31 - 0x3f800000) patterns), results below 2^-120
        # (2 * 0x03800000 patterns incl. zeros and denormals) and ~2^-18 of the rest (near a rounding boundary)
        expected = 2 * (2 ** 31 - 0x3F800000) + 2 * 0x03800000
        assert expected <= literal <= expected + 2 ** 32 // 50000
        bad, _ = r.check_portable_tiers(api.FN_ATAN2, 0, 2 ** 32)
        assert bad == 0
        bad, _ = r.check_portable_tiers(api.FN_ATAN2, 2 ** 40, 2 ** 30)
        assert bad == 0


def test_unchecked_sqrt_rcp_div_sequences_are_the_ieee_operations():
    """every binary32 operand of the valid range: sqrt_mid == __fsqrt_rn, rcp_mid == __frcp_rn; div_mid == __fdiv_rn
    on 2^33 hashed pairs"""
    with api.Renderer(profile=api.PROFILE_V2) as r:
        bad, skipped = r.check_portable_tiers(api.FN_SQRT, 0, 2 ** 32)
        assert bad == 0 and skipped == 2 ** 32 - (0x7F7FFFFF - 0x0D000000 + 1)
        bad, skipped = r.check_portable_tiers(api.FN_RCP, 0, 2 ** 32)
        assert bad == 0 and skipped == 2 ** 32 - 2 * (0x7E800000 - 0x00800000)
        bad, _ = r.check_portable_tiers(api.FN_DIV, 0, 2 ** 33)
        assert bad == 0


def test_bracketed_equirect_texel_index_is_the_exact_one():
    """2^32 hashed (direction, jitter, map size) samples: whenever the binary32 bracket is decisive its texel index
    equals the one from the exact angles; the approximate angles stay within a third of the bracket half-width;
    the bracket decides all but ~1 % of the lookups"""
    with api.Renderer(profile=api.PROFILE_V2) as r:
        n = 2 ** 32
        bad, undecided = r.check_portable_tiers(api.FN_EQUIRECT_TEXEL, 0, n)
        assert bad == 0
        assert undecided < 0.02 * n
