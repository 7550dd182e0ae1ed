"""The certified real-root isolation of the CUDA follow-up kernel of polynomial (oracle/real_root_isolation.py restates it
in plain Python) selects the same t as the reference's scan over all six Durand-Kerner roots (cv2.correctMatches, restated
line by line in oracle/triangulation_oracle.py::correct_matches) -- on every rig, including the forward-motion rig at heavy
noise where most points lack the fast-path certificate and 11 % have no finite search bound."""
import numpy as np
import pytest

import synthetic_rig as rig
from oracle import real_root_isolation as iso
from oracle import triangulation_oracle as orc


@pytest.mark.parametrize("rig_name,sigma,n", [("forward", 0.8, 6000), ("forward", 8.0, 6000), ("forward", 20.0, 4000),
                                              ("rotating", 8.0, 2000), ("general", 8.0, 2000), ("translating", 20.0, 2000)])
def test_real_roots_give_the_reference_selection(rig_name, sigma, n):
    u1, P1, u2, P2, _ = rig.make_correspondences(n, rig_name, sigma)
    F = orc.fundamental_from_P(P1, P2)
    _, _, t_ref, k, (a, b, c, d, f1, f2) = orc.correct_matches(F, u1, u2, return_t='system')
    gave_up = mismatches = 0
    visited = []
    for i in range(n):
        t, v = iso.select_t(k[i], a[i], b[i], c[i], d[i], f1[i], f2[i])
        if t is None:
            gave_up += 1
            continue
        visited.append(v)
        if t != t_ref[i] and not abs(t - t_ref[i]) <= 1e-9 * abs(t_ref[i]):
            mismatches += 1
    assert mismatches == 0
    assert gave_up <= n // 200                  # measured: none
    assert np.mean(visited) < 40 and max(visited) < 400


def test_interval_test_on_known_polynomials():
    # (t - 0.3)(t + 0.45)(t^2 + 1)(t^2 + 4): two real roots in [-1, 1], none outside.  (A root exactly on a dyadic point is
    # recorded by both neighbours -- harmless for the cost scan -- so the test roots are not dyadic.)
    p = np.poly1d([1, -0.3]) * np.poly1d([1, 0.45]) * np.poly1d([1, 0, 1]) * np.poly1d([1, 0, 4])
    k = list(p.coeffs[::-1])
    assert iso.interval_test(k, 1.0, 0, 0) == 2                      # two roots: undecided at the top
    roots = []
    stack = [(0, 0)]
    while stack:
        depth, pos = stack.pop()
        v = iso.interval_test(k, 1.0, depth, pos)
        if v == 2:
            stack += [(depth + 1, 2 * pos), (depth + 1, 2 * pos + 1)]
        elif v == 1:
            mid, h = iso.interval_geometry(1.0, depth, pos)
            roots.append(iso.refine_root(k, mid, h))
    assert sorted(roots) == pytest.approx([-0.45, 0.3], abs=1e-15)
    # reversed polynomial (u = 1/t): no root with |t| >= 1
    assert all(iso.interval_test(k[::-1], 1.0, 3, pos) in (0,) for pos in range(8))
    # a double real root is never "exactly one root": the subdivision runs to its depth limit there (-> Durand-Kerner)
    q = np.poly1d([1, -0.35]) ** 2 * np.poly1d([1, 0, 1]) * np.poly1d([1, 0, 3])
    kq = list(q.coeffs[::-1])
    t, _ = iso.select_t(kq, 1.0, 0.5, 0.2, 0.1, 0.3, 0.4)
    assert t is None
