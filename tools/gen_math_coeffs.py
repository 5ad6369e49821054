"""Generates the polynomial coefficients of the branch-free FP64 asin / atan kernels in csrc/fw_math.cuh.

Chebyshev interpolation in 60-digit arithmetic (mpmath) of
    g(z) = (asin(sqrt z)/sqrt z - 1)/z   on z in [0, 1/4]          -> asin(x) = x + x z g(z),   z = x^2, |x| <= 1/2
    q(w) = (atan(sqrt w)/sqrt w - 1)/w   on w in [0, tan(pi/8)^2]  -> atan(t) = t + t w q(w),   w = t^2
converted to monomial coefficients and rounded to double; the script also reports the worst error of the double
evaluation against mpmath on a dense grid, in ulp of the result."""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_monomial(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    fx = [f(x) for x in xs]
    # Chebyshev coefficients
    c = []
    for j in range(n):
        s = mp.fsum(fx[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n)) for k in range(n)) * 2 / n
        c.append(s)
    c[0] /= 2
    # to monomial in t, then substitute t = (2x - a - b)/(b - a)
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        prev, prev2 = T[-1], T[-2]
        new = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            new[i] -= v
        T.append(new)
    pt = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            pt[i] += c[j] * v
    # t = alpha x + beta
    alpha, beta = 2 / (b - a), -(a + b) / (b - a)
    px = [mp.mpf(0)] * n
    # expand sum pt[i] (alpha x + beta)^i
    for i in range(n):
        for k in range(i + 1):
            px[k] += pt[i] * mp.binomial(i, k) * alpha ** k * beta ** (i - k)
    return px


def g_asin(z):
    if z == 0:
        return mp.mpf(1) / 6
    s = mp.sqrt(z)
    return (mp.asin(s) / s - 1) / z


def q_atan(w):
    if w == 0:
        return -mp.mpf(1) / 3
    s = mp.sqrt(w)
    return (mp.atan(s) / s - 1) / w


def horner(coef, x):
    r = np.full_like(x, coef[-1])
    for c in coef[-2::-1]:
        r = r * x + c
    return r


def report(name, coef, fexact, xs, build):
    cd = np.array([float(c) for c in coef])
    got = build(cd, xs)
    worst = 0.0
    for x, gv in zip(xs[::97], got[::97]):
        ex = fexact(mp.mpf(float(x)))
        if ex == 0:
            continue
        ulp = abs(float(np.spacing(abs(float(ex)))))
        worst = max(worst, abs(float(mp.mpf(float(gv)) - ex)) / ulp)
    print("// %s: degree %d, worst error %.3f ulp on the sampled grid" % (name, len(coef) - 1, worst))
    print("static __device__ __constant__ const double %s[%d] = {" % (name, len(coef)))
    print(",\n".join("    %s" % repr(float(c)) for c in coef))
    print("};")


if __name__ == "__main__":
    deg_asin = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    deg_atan = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ca = cheb_monomial(g_asin, mp.mpf(0), mp.mpf(1) / 4, deg_asin)
    xs = np.linspace(-0.5, 0.5, 200001)
    report("FW_ASIN_P", ca, mp.asin, xs, lambda c, x: x + x * (x * x) * horner(c, x * x))
    t8 = mp.tan(mp.pi / 8)
    cq = cheb_monomial(q_atan, mp.mpf(0), t8 * t8, deg_atan)
    ts = np.linspace(-float(t8), float(t8), 200001)
    report("FW_ATAN_P", cq, mp.atan, ts, lambda c, t: t + t * (t * t) * horner(c, t * t))
