"""gen_minimax.py -- coefficients of the first-tier atan / asin polynomials in csrc/pm_math.cuh.

The parity kernels evaluate atan2f/asinf in two tiers: a short binary64 evaluation (this file's
polynomials, approximate reciprocal / square root refined by Newton steps) decides the binary32
rounding whenever the value is not within 2^-43 (relative) of a rounding boundary; otherwise the
kernel falls back to the literal algorithm of oracle/portable_math.h.  Both tiers approximate the
same real number to better than 2^-48, so they round to the same float outside that band.

Chebyshev interpolation in 80-digit arithmetic, converted to the monomial basis and rounded to
binary64.  usage: python scripts/gen_minimax.py
"""
import mpmath as mp

mp.mp.dps = 80


def cheb_fit(f, a, b, n):
    """degree-n interpolant of f on [a, b] at Chebyshev nodes, monomial coefficients in u"""
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
    A = mp.matrix(n + 1, n + 1)
    y = mp.matrix(n + 1, 1)
    for i, u in enumerate(nodes):
        for j in range(n + 1):
            A[i, j] = u ** j
        y[i] = f(u)
    return list(mp.lu_solve(A, y))


def check(f, coef, a, b, samples=4000):
    c = [mp.mpf(float(x)) for x in coef]
    worst = mp.mpf(0)
    for i in range(samples + 1):
        u = a + (b - a) * mp.mpf(i) / samples
        p = mp.mpf(0)
        for x in reversed(c):
            p = p * u + x
        worst = max(worst, abs(p - f(u)) / abs(f(u)))
    return worst


def f_atan(u):  # atan(t)/t with u = t^2
    if u == 0:
        return mp.mpf(1)
    t = mp.sqrt(u)
    return mp.atan(t) / t


def f_asin(u):  # (asin(t) - t) / t^3 with u = t^2
    if u == 0:
        return mp.mpf(1) / 6
    t = mp.sqrt(u)
    return (mp.asin(t) - t) / (t * u)


if __name__ == "__main__":
    for name, f, a, b, n in (("ATAN", f_atan, mp.mpf(0), mp.mpf(1), 21), ("ASIN", f_asin, mp.mpf(0), mp.mpf(1) / 4, 13)):
        coef = cheb_fit(f, a, b, n)
        err = check(f, coef, a, b)
        print("// %s: degree %d in u, max relative error of the rounded polynomial %.2e (2^%.1f)" % (name, n, float(err), float(mp.log(err, 2))))
        for i, c in enumerate(coef):
            print("    %+.17e,  // u^%d  %s" % (float(c), i, float(c).hex()))
