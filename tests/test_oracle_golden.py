"""
Pins the CPU oracle against the reference's own golden result files:
sampled cells of test_1and2.mat / test_3.mat / figures_scene/test_1and2.mat must be reproduced to
>= 9 significant digits by linear_LS, iterative_LS (C semantics, incl. false-pos/neg of its status) and
linear_eigen with the OpenCV-2 6x4 system.
"""
import json
import os
from functools import partial

import numpy as np
import pytest

from oracle import triangulation_oracle as orc
from harness_replay import replay_cell

RTOL = 2e-9


@pytest.fixture(scope="module")
def cells(golden_dir):
    with open(os.path.join(golden_dir, "golden_cells.json")) as f:
        return json.load(f)


SOLVERS = [partial(orc.linear_eigen_triangulation, rows=6), orc.linear_LS_triangulation,
           orc.iterative_LS_triangulation]


def _pose(tab, t, p):
    tr = tab[t]
    return tr["sideways_values"][p], tr["towards_values"][p], tr["angle_values"][p]


def _check(got, cell, methods, keys, rtol=RTOL):
    for k in keys:
        for ti in methods:
            want = cell[k][ti]
            if want is None:
                continue
            assert got[k][ti] == pytest.approx(want, rel=rtol, abs=1e-12), (k, ti, got[k][ti], want)


ALL_KEYS = ("err3D_mean_summary", "err3D_median_summary", "err2D_mean_summary", "err2D_median_summary",
            "false_pos_summary", "false_neg_summary")


@pytest.mark.parametrize("ci", range(10))
def test_test_1and2_cells(cells, ci):
    g = cells["test_1and2"]
    cell = g["cells"][ci]
    got = replay_cell(SOLVERS, _pose(g["trajectories"], cell["traj"], cell["pose"]), num_trials=g["num_trials"],
                      rseed=g["rseed"])
    # linear_LS (err2D mean on the forward trajectory is 0/0 at the camera centre: see below)
    _check(got, cell, (1,), [k for k in ALL_KEYS if not (cell["traj"] == 1 and k.startswith("err2D"))])
    if cell["traj"] != 1:
        _check(got, cell, (2,), ALL_KEYS)                                 # iterative_LS (C semantics)
        _check(got, cell, (0,), ALL_KEYS)                                 # linear_eigen, OpenCV-2 6x4 system
    else:
        # Forward motion: 9 of the 257 lattice points lie exactly on the baseline.  With integer pixels their
        # depth in camera 2 is 0 up to rounding, so the C code's `d2_new == 0` break (triangulation.c:138) is
        # decided by the last bit of the SVD -- not reproducible across SVD implementations.  Those 3.5 % of
        # the points move the statistics by < 1 %; linear_eigen's means are NaN/huge for the same reason.
        _check(got, cell, (2,), [k for k in ALL_KEYS if not k.startswith("err2D")], rtol=1e-2)
        _check(got, cell, (0,), ("err3D_median_summary",))


def test_iterative_status_semantics_differ(cells):
    """F2: the golden false-negative ratio is only reproduced by the C control flow, not the Python one."""
    g = cells["test_1and2"]
    cell = [c for c in g["cells"] if (c["traj"], c["pose"]) == (4, 39)][0]
    py = partial(orc.iterative_LS_triangulation, semantics='py')
    got = replay_cell([orc.iterative_LS_triangulation, py], _pose(g["trajectories"], 4, 39),
                      num_trials=g["num_trials"], rseed=g["rseed"])
    assert got["false_neg_summary"][0] == pytest.approx(cell["false_neg_summary"][2], rel=RTOL)
    assert got["false_neg_summary"][1] == 0.0
    assert cell["false_neg_summary"][2] > 0.4


@pytest.mark.parametrize("ci", range(4))
def test_test_3_cells(cells, ci):
    g = cells["test_3"]
    cell = g["cells"][ci]
    nt = cell["ntype"]
    got = replay_cell(SOLVERS, _pose(g["trajectories"], cell["traj"], -1), num_trials=g["num_trials"],
                      rseed=g["rseed"], sigma=g["noise_sigma_values"][cell["nidx"]], discretized=nt >= 1,
                      k1=0.3 if nt == 2 else 0.)
    _check(got, cell, (1, 2), ALL_KEYS)


@pytest.mark.parametrize("ci", range(2))
def test_scene_cells(cells, ci):
    g = cells["scene_test_1and2"]
    cell = g["cells"][ci]
    got = replay_cell(SOLVERS, _pose(g["trajectories"], cell["traj"], cell["pose"]), num_trials=g["num_trials"],
                      rseed=g["rseed"], points_3D=np.array(g["points_3D"]))
    _check(got, cell, (1, 2), ALL_KEYS)


@pytest.mark.parametrize("traj", [0, 3, 4])
def test_vector_stat_reproduces_the_reference_results(cells, traj, golden_dir):
    """vector_stat (triangulation_comparison.py:219-240): the per-point mean vectors / covariance matrices the reference
    stored for the last pose of a trajectory (test_1and2.mat, p_err3Dv_*_summary) are reproduced by the harness replay
    (oracle solvers + NumPy restatement of vector_stat) -- the pin the device kernel is tested against on the GPU."""
    from harness_replay import vector_stat
    g = cells["test_1and2"]
    vs = np.load(os.path.join(golden_dir, "vector_stat_cells.npz"))
    got = replay_cell([orc.linear_LS_triangulation, orc.iterative_LS_triangulation], _pose(g["trajectories"], traj, 39),
                      num_trials=g["num_trials"], rseed=g["rseed"], return_vectors=True)
    for i, name in enumerate(("linear_LS_triangulation", "iterative_LS_triangulation")):
        means, covars = vector_stat(got["error_vectors_3D"][i])
        gm, gc = vs["mean_%d_%s" % (traj, name)], vs["covar_%d_%s" % (traj, name)]
        scale = np.sqrt(np.trace(gc, axis1=1, axis2=2))[:, None]
        assert np.max(np.abs(means - gm) / scale) < 1e-9
        assert np.max(np.abs(covars - gc) / (scale ** 2)[:, :, None]) < 1e-9


def test_multiview_oracle_matches_independent_lstsq():
    """Independent pin of the m-view oracle for m > 2 (the reference has no m-view call): per point, the stacked 2m x 3
    system of triangulation.c:30-40 solved by np.linalg.lstsq (LAPACK gelsd, minimum norm) -- a different code path from the
    oracle's own batched SVD with OpenCV's rank rule -- on fully and partially observed points."""
    import synthetic_rig as rig
    for m, p_vis in ((3, 1.0), (5, 0.7), (8, 0.6), (16, 1.0)):
        us, Ps, X, valid = rig.make_multiview(300, m, 0.8, seed=rig.RSEED + m, p_visible=p_vis)
        xo, so = orc.multiview_LS_triangulation(us, Ps, valid if p_vis < 1.0 else None)
        for i in range(us.shape[1]):
            rows, rhs = [], []
            for v in range(m):
                if not valid[v, i]:
                    continue
                P = np.asarray(Ps[v])
                for k in range(2):
                    r = us[v, i, k] * P[2] - P[k]
                    rows.append(r[0:3]); rhs.append(-r[3])
            if len(rows) < 4:
                continue                                   # seen by < 2 views: status False, minimum-norm by definition
            want = np.linalg.lstsq(np.array(rows), np.array(rhs), rcond=None)[0]
            assert so[i]
            assert np.max(np.abs(xo[i] - want)) <= 1e-10 * max(1.0, np.max(np.abs(want))), (m, i)


def test_multiview_oracle_reduces_to_linear_ls_and_ignores_masked_views():
    """The m-view oracle is the stacked system of the two-view one: identical for m = 2; masked
    views change nothing; more views of the same noise level move the estimate towards the ground truth."""
    import synthetic_rig as rig
    u1, P1, u2, P2, X = rig.make_correspondences(2000, "rotating", 0.8)
    x2, _ = orc.linear_LS_triangulation(u1, P1, u2, P2)
    xm, st = orc.multiview_LS_triangulation(np.stack([u1, u2]), [P1, P2])
    assert np.array_equal(x2, xm) and st.all()
    us, Ps, Xm, valid = rig.make_multiview(2000, 6, 0.8, p_visible=0.7)
    xa, sa = orc.multiview_LS_triangulation(us, Ps, valid)
    junk = us.copy(); junk[~valid] = 1e6
    xb, sb = orc.multiview_LS_triangulation(junk, Ps, valid)
    assert np.array_equal(xa, xb) and np.array_equal(sa, valid.sum(axis=0) >= 2)
    full, _ = orc.multiview_LS_triangulation(us, Ps)
    pair, _ = orc.linear_LS_triangulation(us[0], Ps[0], us[-1], Ps[-1])
    assert np.linalg.norm(full - Xm, axis=1).mean() < np.linalg.norm(pair - Xm, axis=1).mean()
