"""The two-ray closed form of the CUDA kernels (oracle/two_ray_form.py restates it in NumPy) against the line-by-line
oracle of the reference: same status vector and points for iterative_LS, the least-squares point for linear_LS, and the
smallest singular vector of the DLT system once the match satisfies the epipolar constraint (polynomial)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multiple-quadrotor-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "harness"))
import synthetic_rig as rig                                     # noqa: E402
from oracle import triangulation_oracle as orc                  # noqa: E402
from oracle import two_ray_form as trf                           # noqa: E402

RIGS = ["translating", "rotating", "forward", "general"]


def rel(x, ref):
    with np.errstate(all='ignore'):
        return np.max(np.abs(x - ref), axis=1) / np.max(np.abs(ref), axis=1)


@pytest.mark.parametrize("rig_name", RIGS)
@pytest.mark.parametrize("sigma", [0.8, 20.0])
def test_closed_form_is_the_reference_loop(rig_name, sigma):
    u1, P1, u2, P2, X = rig.make_correspondences(20011, rig_name, sigma)
    xo, so, n_solves, margin = orc.iterative_LS_core(u1, P1, u2, P2)
    x, st, ns, s = trf.iterative_ls(u1, P1, u2, P2)
    certified = s["k2"] < 1e10                                   # the kernel's limit; beyond it the reference loop runs
    assert certified.mean() > 0.999
    knife = margin < 1e-9                                        # convergence test decided in the last bits
    keep = certified & ~knife
    assert np.array_equal(st[keep], so[keep])
    assert np.array_equal(ns[keep], n_solves[keep])
    assert rel(x, xo)[keep].max() < 1e-9
    assert np.percentile(rel(x, xo)[keep], 99.9) < 1e-11


@pytest.mark.parametrize("semantics", ["c", "py"])
def test_closed_form_both_control_flows(semantics):
    u1, P1, u2, P2, X = rig.make_correspondences(5003, "rotating", 4.0)
    xo, so, _, margin = orc.iterative_LS_core(u1, P1, u2, P2, 3e-5, semantics)
    x, st, _, s = trf.iterative_ls(u1, P1, u2, P2, 3e-5, semantics)
    keep = margin > 1e-9
    assert np.array_equal(st[keep], so[keep]) and rel(x, xo)[keep].max() < 1e-9
    assert (0 in so) == (semantics == 'c')                       # status 0 is unreachable in the Python flow


@pytest.mark.parametrize("rig_name", RIGS)
def test_unweighted_point_is_linear_ls(rig_name):
    u1, P1, u2, P2, X = rig.make_correspondences(20011, rig_name, 8.0)
    xo, _ = orc.linear_LS_triangulation(u1, P1, u2, P2)
    x, s = trf.linear_ls(u1, P1, u2, P2)
    keep = s["k2"] < 1e10
    assert keep.mean() > 0.999 and rel(x, xo)[keep].max() < 1e-9


@pytest.mark.parametrize("rig_name", RIGS)
def test_ray_intersection_is_the_eigen_solution_after_correction(rig_name):
    """polynomial: after cv2.correctMatches the rays meet (certificate: residual ratio <= 1e-11) and the dehomogenised
    smallest singular vector equals the least-squares point."""
    u1, P1, u2, P2, X = rig.make_correspondences(20011, rig_name, 4.0)
    n1, n2 = orc.correct_matches(orc.fundamental_from_P(P1, P2), u1, u2)
    Xh = np.asarray(orc.eigen_homogeneous(n1, P1, n2, P2, rows=4))
    xo = Xh[:, 0:3] / Xh[:, 3:4]
    x, s = trf.linear_ls(n1, P1, n2, P2)
    certified = (s["res"] <= 1e-11) & (s["k2"] < 1e10) & np.isfinite(xo).all(axis=1)
    assert certified.mean() > 0.99
    assert rel(x, xo)[certified].max() < 1e-9
    # without the correction the certificate must fail: the noisy rays do not meet
    _, s_raw = trf.linear_ls(u1, P1, u2, P2)
    assert (s_raw["res"] <= 1e-11).mean() < 0.01
