"""Accuracy of the dynamics kernel's branch-free fp64 math (csrc/fwmath.cuh) against numpy/libm.

The header is plain C++ under g++ (the device build only swaps the ~20-bit MUFU seeds and the constant tables), so the
polynomials, reductions and Newton budgets are checked here on the CPU; the GPU parity tests then cover the device
seeds end to end.  Bound: <= 4 ulp everywhere tested (the parity budget is 1e-9 relative per env step).
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "fwmath_host.cpp")
OUT = os.path.join(HERE, "native", "libfwmath_host.so")
HDR = os.path.join(HERE, "..", "fixed-wing-gym_b200", "csrc", "fwmath.cuh")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def call(lib, name, *arrs, scalar=None):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    out = np.empty_like(arrs[0])
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    args = [ptr(a) for a in arrs]
    if scalar is not None:
        args.append(ctypes.c_double(scalar))
    getattr(lib, name)(*args, ptr(out), ctypes.c_long(out.size))
    return out


def ulps(a, ref):
    return np.abs(a - ref) / np.spacing(np.abs(ref))


def test_atan2(lib):
    rng = np.random.RandomState(0)
    n = 400000
    x = np.concatenate([rng.normal(0, 20, n), rng.uniform(5, 40, n), [0.0, 1.0, -1.0, 0.0, 3.0, -0.0]])
    y = np.concatenate([rng.normal(0, 20, n), rng.normal(0, 5, n), [0.0, 0.0, 0.0, 2.0, 3.0, 1.0]])
    got, ref = call(lib, "t_atan2", y, x), np.arctan2(y, x)
    assert ulps(got, ref).max() <= 4
    assert call(lib, "t_atan2", [0.0], [0.0])[0] == 0.0
    assert np.isnan(call(lib, "t_atan2", [np.nan], [1.0])[0])


def test_exp(lib):
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.uniform(-180, 180, 400000), rng.uniform(-1, 1, 100000), [0.0, 700.0, -700.0]])
    assert ulps(call(lib, "t_exp", x), np.exp(x)).max() <= 2
    assert np.isnan(call(lib, "t_exp", [np.nan])[0])


def test_log_pow(lib):
    rng = np.random.RandomState(2)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 300000), rng.uniform(0.5, 2, 100000), [1.0, 2.0, 0.5]])
    got, ref = call(lib, "t_log", x), np.log(x)
    assert (np.abs(got - ref) <= 4 * np.spacing(np.maximum(np.abs(ref), 1.0))).all()
    for p in (-0.2, 0.2):
        got, ref = call(lib, "t_pow", x, scalar=p), x ** p
        assert (np.abs(got / ref - 1) <= 1e-13).all()   # |p ln x| <= 140 amplifies the log's ulp
    assert np.isnan(call(lib, "t_pow", [np.nan], scalar=-0.2)[0])
    # err = inf: NaN, which the step controller's `fac > 0.2 ? fac : 0.2` turns into 0.2 like Python's max(0.2, nan)
    assert not call(lib, "t_pow", [np.inf], scalar=-0.2)[0] > 0.2


def test_sqrt_rcp_div(lib):
    rng = np.random.RandomState(3)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 300000), rng.uniform(0, 1000, 100000)])
    assert ulps(call(lib, "t_sqrt", x), np.sqrt(x)).max() <= 1
    assert ulps(call(lib, "t_rsqrt", x), 1 / np.sqrt(x)).max() <= 2
    assert ulps(call(lib, "t_rcp", x), 1 / x).max() <= 1
    b = np.concatenate([x, -x])
    a = rng.normal(0, 100, b.size)
    assert ulps(call(lib, "t_div", a, b), a / b).max() <= 1
    assert call(lib, "t_sqrt", [0.0])[0] == 0.0
    assert np.isnan(call(lib, "t_sqrt", [np.nan])[0])


def test_sincospi(lib):
    import mpmath as mp
    rng = np.random.RandomState(4)
    x = np.concatenate([rng.uniform(0, 2, 200000), [0.0, 0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 1.9999999]])
    s, c = call(lib, "t_sinpi", x), call(lib, "t_cospi", x)
    # reference: exact sin/cos of pi*x (quadrant reduction done in exact arithmetic)
    k = np.rint(2 * x)
    r = x - k / 2
    sr, cr = np.sin(np.pi * r), np.cos(np.pi * r)
    q = k.astype(int) % 4
    s_ref = np.where(q == 0, sr, np.where(q == 1, cr, np.where(q == 2, -sr, -cr)))
    c_ref = np.where(q == 0, cr, np.where(q == 1, -sr, np.where(q == 2, -cr, sr)))
    assert np.abs(s - s_ref).max() <= 4e-16 and np.abs(c - c_ref).max() <= 4e-16
    for xv in (0.3, 1.7, 0.9999):
        assert abs(call(lib, "t_sinpi", [xv])[0] - float(mp.sin(mp.pi * mp.mpf(xv)))) < 4e-16
