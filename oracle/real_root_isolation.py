"""
Restatement, in plain Python floats, of the CERTIFIED REAL-ROOT ISOLATION the CUDA follow-up kernel of polynomial uses
instead of Durand-Kerner (multiple-quadrotor-slam_b200/csrc/trgl_hartley_sturm.cuh: hs_interval_test, hs_refine_root,
hs_scan_roots; k_polynomial_general runs the subdivision level by level over a CTA, here it is a depth-first walk -- the
set of recorded intervals is the same).  Test infrastructure only: it lets the CPU suite check, without a GPU, that
scanning the real roots of g selects the same t as the reference's scan over the real parts of all six Durand-Kerner roots
(cv2.correctMatches; oracle/triangulation_oracle.py::correct_matches follows that line by line).

Why it must: the real part of a complex root is some real t, s(t) at any real t is at least the global minimum of s over
the reals, and that minimum is attained at a real root of g = numerator of s' (or at infinity).
"""
import math

import numpy as np

DBL_MAX = 1.7976931348623157e308
MAX_DEPTH = 23
MAX_ROOTS = 8


def cost(t, a, b, c, d, f1, f2):
    return t * t / (1 + f1 * f1 * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f2 * f2 * (c * t + d) ** 2)


def taylor_shift(p, mid):
    c = list(p)
    for i in range(6):
        for j in range(5, i - 1, -1):
            c[j] += mid * c[j + 1]
    return c


def interval_geometry(R, depth, pos):
    w = math.ldexp(2.0 * R, -depth)
    h = 0.5 * w
    return w * pos - R + h, h


def interval_test(p, R, depth, pos):
    """0: no root of p in the interval; 1: exactly one; 2: undecided."""
    mid, h = interval_geometry(R, depth, pos)
    c = taylor_shift(p, mid)
    rest0 = 0.0
    for j in range(6, 0, -1):
        rest0 = (rest0 + abs(c[j])) * h
    noise = 0.0
    for j in range(6, -1, -1):
        noise = noise * abs(mid) + abs(p[j])
    rest1 = 0.0
    for j in range(6, 1, -1):
        rest1 = rest1 * h + j * abs(c[j])
    rest1 *= h
    glo = ghi = c[6]
    for j in range(5, -1, -1):
        glo = glo * (-h) + c[j]
        ghi = ghi * h + c[j]
    if abs(c[0]) > rest0 * (1.0 + 1e-9) + 1e-13 * noise:
        return 0
    if not abs(c[1]) > rest1 * (1.0 + 1e-9):
        return 2
    return 1 if ((glo < 0.0) != (ghi < 0.0)) or glo == 0.0 or ghi == 0.0 else 0


def refine_root(p, mid, h):
    c = taylor_shift(p, mid)
    glo = c[6]
    for j in range(5, -1, -1):
        glo = glo * (-h) + c[j]
    xl, xh = -h, h
    x = min(max(-c[0] / c[1], -h), h)
    for _ in range(64):
        g = c[6]; dg = 0.0
        for j in range(5, -1, -1):
            dg = dg * x + g
            g = g * x + c[j]
        if g == 0.0:
            break
        if (g < 0.0) == (glo < 0.0):
            xl = x
        else:
            xh = x
        xn = x - g / dg
        if xn == x:
            break
        newton = xl <= xn <= xh
        if not newton:
            xn = 0.5 * (xl + xh)
        step = abs(xn - x)
        x = xn
        if (newton and step <= 1e-8 * abs(mid + xn)) or (xh - xl) <= 4e-16 * abs(mid + xn):
            break
    return mid + x


def select_t(k, a, b, c, d, f1, f2):
    """t of the reference's cost scan, from the real roots of g only.  Returns (t, intervals visited) or (None, visited)
    when the subdivision cannot certify (the CUDA kernel then runs Durand-Kerner)."""
    k = [float(v) for v in k]
    with np.errstate(all='ignore'):
        s0 = float(np.float64(d * d) / np.float64(b * b + f2 * f2 * d * d))
    bounded = f1 * f1 * s0 < 1.0
    T0 = 0.0
    if bounded:
        T0s = s0 / (1.0 - f1 * f1 * s0)
        T0 = math.sqrt(T0s) * (1.0 + 1e-9) if T0s > 0.0 else T0s
    R0 = min(T0, 1.0) if bounded else 1.0
    recorded = []
    visited = 0
    for dom in (0, 1):
        if dom == 1 and bounded and T0 <= 1.0:
            break
        p = k[::-1] if dom else k
        R = 1.0 if dom else R0
        stack = [(0, 0)]
        while stack:
            depth, pos = stack.pop()
            visited += 1
            verdict = interval_test(p, R, depth, pos)
            if verdict == 2:
                if depth == MAX_DEPTH:
                    return None, visited
                stack.append((depth + 1, 2 * pos + 1)); stack.append((depth + 1, 2 * pos))
            elif verdict == 1:
                if len(recorded) == MAX_ROOTS:
                    return None, visited
                recorded.append((dom, depth, pos))
    with np.errstate(all='ignore'):                    # IEEE semantics like the C / CUDA code: x/0 = inf, 0/0 = NaN
        s_val = float(np.float64(1.0) / np.float64(f1 * f1) + np.float64(c * c) / np.float64(a * a + f2 * f2 * c * c))
    t_min = DBL_MAX
    found = 0
    for dom, depth, pos in recorded:
        mid, h = interval_geometry(1.0 if dom else R0, depth, pos)
        t = refine_root(k[::-1] if dom else k, mid, h)
        if dom:
            if t == 0.0:
                continue
            t = 1.0 / t
        found += 1
        sv = cost(t, a, b, c, d, f1, f2)
        if sv < s_val or (sv == s_val and t_min != DBL_MAX and t < t_min):
            s_val = sv; t_min = t
    if bounded and found == 0:
        return None, visited
    return t_min, visited
