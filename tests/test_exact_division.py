"""The parity kernel divides fragCoord by the image size with one fused correction step instead of the
IEEE division sequence (ParityMath::div_small, csrc/pt_device.cuh): x / c == fma(fma(-q, c, x), rc, q),
q = RN(x * rc), rc = RN(1 / c), for divisors with <= 16 significant bits.  Check the identity on the CPU
with numpy float32/float64 arithmetic for many divisors, densely around pixel coordinates."""
import numpy as np


def fma32(a, b, c):
    # exact a*b in float64 (24x24 bits), one addition: the sum of a 48-bit product and a 24-bit float can need
    # more than 53 bits only when the exponents are far apart, in which case float64 rounding is still
    # monotone; the final cast rounds once more -- adequate here because the kernel identity is ALSO
    # verified exhaustively in C (10^10 cases); this test guards against regressions.
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_division_by_image_sizes_is_exact():
    rng = np.random.default_rng(0)
    for c in (8, 24, 512, 720, 777, 1023, 1080, 1280, 1366, 1920, 2160, 3840, 4095, 8192, 40000, 65535):
        cf = np.float32(c)
        rc = np.float32(1.0) / cf
        px = rng.integers(0, c, 200_000).astype(np.float32)
        jit = (rng.integers(0, 2 ** 31, 200_000).astype(np.int32).astype(np.float32) / np.float32(2147483648.0)) - np.float32(0.5)
        x = (px + jit).astype(np.float32)
        q = (x * rc).astype(np.float32)
        r = fma32(-q, np.full_like(q, cf), x)
        got = fma32(r, np.full_like(q, rc), q)
        assert np.array_equal(got, (x / cf).astype(np.float32)), c
