"""Elements of Julia float ranges, for the host glue that sizes and fills time grids.

`t0:dt:tf` and `range(a, b, length = n)` are StepRangeLen{T, TwicePrecision{T}, TwicePrecision{T}} in Julia
(base/twiceprecision.jl): start and step are lifted to exact rationals when they have short continued fractions
(`Base.rat`), and element k is evaluated in twice the working precision and rounded ONCE -- so (0:0.1:1)[4] == 0.3
although 0.1 * 3 != 0.3 in Float64.  The reference relies on this in two places of
src/ensemblegpukernel/lowerlevel_solve.jl: `count(x -> x in tstops, timeseries)` (:74-77) decides the number of output
rows, and `Tt.(collect(range(...)))` (:90-91, :104, :276-303) produces the saveat grid.  A naive `t0 + k*dt` in the
working precision disagrees with both in Float64 (one row too many; save times off by an ulp).

Here the elements are computed with exact rational arithmetic and rounded once, which is what the twice-precision
evaluation amounts to.  The range LENGTH stays with libdegk (`degk_output_rows`, the same rational lift in C++)."""
from fractions import Fraction

import numpy as np


def _rat(x, dtype):
    """Base.rat(x): continued fraction of x until it reproduces x (or the terms leave the exact range)."""
    T = np.dtype(dtype).type
    mn = 2048.0 if np.dtype(dtype) == np.float32 else 16777216.0
    y = T(x)
    a, d, b, c = 1, 1, 0, 0
    while abs(float(y)) <= mn:
        f = int(np.trunc(float(y)))
        y = T(y - T(f))
        a, b, c, d = f * a + c, f * b + d, a, b
        if max(abs(a), abs(b)) > mn:
            return c, d
        if b != 0 and T(T(a) / T(b)) == T(x):
            break
        if y == 0:
            break
        y = T(T(1) / y)
    return a, b


def _lift(x, dtype):
    """exact rational Julia uses for x: the short continued fraction if it reproduces x, else x's binary value"""
    T = np.dtype(dtype).type
    n, d = _rat(x, dtype)
    if d != 0 and T(T(n) / T(d)) == T(x):
        return Fraction(n, d)
    return Fraction(float(T(x)))


def _round(fr, dtype):
    return np.dtype(dtype).type(float(fr))          # float(Fraction) is correctly rounded to Float64


def step_range(start, step, length, dtype):
    """collect(start:step:stop) given its length (from degk_output_rows)"""
    s, p = _lift(start, dtype), _lift(step, dtype)
    return np.array([_round(s + k * p, dtype) for k in range(int(length))], dtype=dtype)


def lin_range(start, stop, length, dtype):
    """collect(range(start, stop, length = n))"""
    n = int(length)
    if n == 1:
        return np.array([np.dtype(dtype).type(start)], dtype=dtype)
    a, b = _lift(start, dtype), _lift(stop, dtype)
    out = [_round(a + (b - a) * Fraction(k, n - 1), dtype) for k in range(n)]
    out[0], out[-1] = np.dtype(dtype).type(start), np.dtype(dtype).type(stop)
    return np.array(out, dtype=dtype)
