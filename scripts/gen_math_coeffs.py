#!/usr/bin/env python
"""Generate the polynomial coefficients of csrc/fwmath.cuh (branch-free fp64 atan / exp / log kernels).

Near-minimax by Chebyshev interpolation at 60-digit precision (mpmath), converted to the monomial basis.  For each
function the degree is the smallest whose interpolation error is below the stated bound.  Run:  python
scripts/gen_math_coeffs.py  -> prints C arrays (pasted into fwmath.cuh) with the achieved max error.
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    """Degree n-1 interpolant of f on [a,b] at n Chebyshev nodes -> monomial coefficients (low to high)."""
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def max_err(f, c, a, b, m=2000):
    e = mp.mpf(0)
    for k in range(m + 1):
        x = a + (b - a) * k / m
        p = mp.polyval(list(reversed(c)), x)
        e = max(e, abs(p - f(x)))
    return e


def fit(name, f, a, b, tol, nmin=4, nmax=24):
    for n in range(nmin, nmax):
        c = cheb_fit(f, a, b, n)
        e = max_err(f, c, a, b)
        if e < tol:
            print("// %s: %d coefficients on [%s, %s], max abs error %s" % (name, n, mp.nstr(a, 8), mp.nstr(b, 8),
                                                                            mp.nstr(e, 3)))
            print("static const double %s[%d] = {" % (name, n))
            print(",\n".join("    %s" % float(x).hex() + "  /* %s */" % mp.nstr(x, 20) for x in c))
            print("};")
            return c
    raise RuntimeError(name)


def atan_R(z):      # atan(t) = t + t*z*R(z), z = t*t
    if abs(z) < mp.mpf(10) ** -20:
        return -mp.mpf(1) / 3 + z / 5
    t = mp.sqrt(z)
    return (mp.atan(t) / t - 1) / z


def exp_P(r):       # exp(r) = 1 + r + r*r*P(r)
    if abs(r) < mp.mpf(10) ** -20:
        return mp.mpf(1) / 2 + r / 6
    return (mp.exp(r) - 1 - r) / (r * r)


def log_L(w):       # ln(m) = 2f + 2f*w*L(w), f = (m-1)/(m+1), w = f*f   (atanh series: L = 1/3 + w/5 + ...)
    if abs(w) < mp.mpf(10) ** -20:
        return mp.mpf(1) / 3 + w / 5
    f = mp.sqrt(w)
    return (mp.atanh(f) / f - 1) / w


def sinpi_S(z):     # sin(pi r) = r*(pi + z*S(z)), z = r*r
    if abs(z) < mp.mpf(10) ** -20:
        return -mp.pi ** 3 / 6 + z * mp.pi ** 5 / 120
    r = mp.sqrt(z)
    return (mp.sin(mp.pi * r) / r - mp.pi) / z


def cospi_C(z):     # cos(pi r) = 1 + z*C(z)
    if abs(z) < mp.mpf(10) ** -20:
        return -mp.pi ** 2 / 2 + z * mp.pi ** 4 / 24
    return (mp.cos(mp.pi * mp.sqrt(z)) - 1) / z


if __name__ == "__main__":
    tz = mp.tan(mp.pi / 8) ** 2
    fit("FW_ATAN_R", atan_R, mp.mpf(0), tz * (1 + mp.mpf(10) ** -6), mp.mpf(2) ** -54)
    h = mp.log(2) / 2 * (1 + mp.mpf(10) ** -6)
    fit("FW_EXP_P", exp_P, -h, h, mp.mpf(2) ** -52)
    fw = ((mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)) ** 2
    fit("FW_LOG_L", log_L, mp.mpf(0), fw * (1 + mp.mpf(10) ** -6), mp.mpf(2) ** -49)
    q = mp.mpf(1) / 16 * (1 + mp.mpf(10) ** -6)
    fit("FW_SINPI_S", sinpi_S, mp.mpf(0), q, mp.mpf(2) ** -50)
    fit("FW_COSPI_C", cospi_C, mp.mpf(0), q, mp.mpf(2) ** -51)
