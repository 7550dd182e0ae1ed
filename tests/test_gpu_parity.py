"""
GPU parity tests (run with -m gpu on the B200 box): every solver, called through the reference-facing Python API
(`triangulation`, which goes through the C ABI of libtriangl_cuda.so), against the CPU oracle on the same seeded
inputs, against the fixtures produced by the reference's own Python code, and against the reference's golden cells.

Bars (BASELINE.json north_star): status / good mask bit-exact; x within 1e-9 relative in FP64, 1e-4 in FP32 mode.
"rel" is per point: max_k |x_k - ref_k| / max_k |ref_k|.
Two documented classes of points are excluded from the bit-exact claim because no two correct implementations agree
on them (DESIGN.md "Ill-posed points"): (i) iterative_LS points whose convergence test |d_new - d| <= tol is decided
by less than 1e-9 (knife edge), (ii) linear_eigen / polynomial points whose homogeneous w or singular-value gap is at
rounding level (points at infinity, on the baseline).
"""
import glob
import json
import os

import numpy as np
import pytest

import synthetic_rig as rig
from oracle import triangulation_oracle as orc
from harness_replay import replay_cell

pytestmark = pytest.mark.gpu

TOL64 = 1e-9
TOL32 = 1e-4
RIG_NAMES = ["translating", "rotating", "forward", "general"]


@pytest.fixture(scope="module")
def tri():
    import triangl_cuda
    triangl_cuda.require_device()
    import triangulation
    yield triangulation
    triangulation.set_triangl_output_dtype(float)
    triangulation.set_triangl_compute_dtype(np.float64)
    triangulation.set_triangl_semantics('c', 4)


def rel_err(x, ref):
    with np.errstate(all='ignore'):
        return np.max(np.abs(np.asarray(x, dtype=np.float64) - ref), axis=1) / np.max(np.abs(ref), axis=1)


def eigen_well_posed(u1, P1, u2, P2, rows=4, tol=1e-9):
    """Points where the smallest singular vector is determined to `tol` in double: eps*s1/(s3-s4) and |w| margins."""
    X = orc.eigen_homogeneous(u1, P1, u2, P2, rows)
    A = np.empty((len(u1), 4, 4))
    A[:, 0] = u1[:, 0:1] * P1[2] - P1[0]; A[:, 1] = u1[:, 1:2] * P1[2] - P1[1]
    A[:, 2] = u2[:, 0:1] * P2[2] - P2[0]; A[:, 3] = u2[:, 1:2] * P2[2] - P2[1]
    s = np.linalg.svd(A, compute_uv=False)
    with np.errstate(all='ignore'):
        amp = s[:, 0] / (s[:, 2] - s[:, 3]) / np.abs(X[:, 3])
    return amp * 2.2e-16 * 50 < tol


def iterative_margin(u1, P1, u2, P2, semantics='c'):
    """Distance of the deciding convergence test |d_new - d| <= tol from its threshold (knife-edge points < 1e-9)."""
    return orc.iterative_LS_core(np.asarray(u1, dtype=np.float64), P1, np.asarray(u2, dtype=np.float64), P2, 3e-5, semantics)[3]


def ls_well_posed(u1, P1, u2, P2, tol=1e-9):
    A, _ = orc.build_Ab(u1, P1, u2, P2)
    s = np.linalg.svd(A, compute_uv=False)
    return (s[:, 0] / s[:, 2]) * 2.2e-16 * 50 < tol


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.0, 0.8, 8.0])
def test_linear_ls_vs_oracle(tri, rig_name, sigma):
    u1, P1, u2, P2, X = rig.make_correspondences(30011, rig_name, sigma)
    x, st = tri.linear_LS_triangulation(u1, P1, u2, P2)
    xo, so = orc.linear_LS_triangulation(u1, P1, u2, P2)
    assert x.shape == (30011, 3) and x.dtype == np.float64 and st.dtype == np.bool_
    assert st.all() and np.array_equal(st, so)
    ok = ls_well_posed(u1, P1, u2, P2)
    assert ok.mean() >= 0.995
    assert rel_err(x, xo)[ok].max() < TOL64
    # ill-conditioned remainder still agrees to its conditioning
    A, _ = orc.build_Ab(u1, P1, u2, P2)
    s = np.linalg.svd(A, compute_uv=False)
    assert np.all(rel_err(x, xo) < np.maximum(TOL64, (s[:, 0] / s[:, 2]) ** 2 * 1e-14))


@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.0, 0.8, 8.0])
@pytest.mark.parametrize("semantics", ["c", "py"])
def test_iterative_ls_vs_oracle(tri, rig_name, sigma, semantics):
    u1, P1, u2, P2, X = rig.make_correspondences(30011, rig_name, sigma)
    tri.set_triangl_semantics(semantics)
    try:
        x, st = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    finally:
        tri.set_triangl_semantics('c')
    xo, so, nsolves, margin = orc.iterative_LS_core(u1, P1, u2, P2, 3e-5, semantics)
    assert st.dtype == (np.int32 if semantics == 'c' else np.int64)
    knife = margin < 1e-9
    assert knife.mean() < 1e-3
    ok = ls_well_posed(u1, P1, u2, P2) & ~knife
    assert ok.mean() >= 0.995
    assert np.array_equal(st[ok], so[ok])
    assert rel_err(x, xo)[ok].max() < TOL64
    assert set(np.unique(st)).issubset({1, 0, -1, -2, -3})


@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.8, 20.0])
def test_iterative_ls_closed_form_equals_reference_loop(tri, rig_name, sigma):
    """The two-ray closed form (default) and the reference's loop as written (trgl_set_two_ray(0)) are the same
    function: identical status vector away from knife-edge convergence tests, points within 1e-9."""
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(50021, rig_name, sigma)
    x, st = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    old = tc.set_two_ray(0)
    try:
        xg, stg = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    finally:
        tc.set_two_ray(old)
    assert old == 1
    knife = iterative_margin(u1, P1, u2, P2) < 1e-9
    ok = ls_well_posed(u1, P1, u2, P2) & ~knife
    assert ok.mean() >= 0.995
    assert np.array_equal(st[ok], stg[ok])
    assert rel_err(x, xg)[ok].max() < TOL64
    # away from the certified set both still agree to the conditioning of the system
    A, _ = orc.build_Ab(u1, P1, u2, P2)
    s = np.linalg.svd(A, compute_uv=False)
    same = st == stg
    assert np.all(rel_err(x, xg)[same & ~knife] < np.maximum(TOL64, (s[:, 0] / s[:, 2])[same & ~knife] ** 2 * 1e-13))


def test_deferred_list_overflow_redoes_every_point(tri):
    """More uncertified points than the deferred list holds: the follow-up kernels fall back to redoing every point.
    The list is limited to 1000 slots; identical cameras make every system rank 2, i.e. every point uncertified, and the
    forward-motion rig under 20 px noise defers most of its points in polynomial (Durand-Kerner) and some in linear_eigen."""
    import triangl_cuda as tc
    old = tc.set_deferred_capacity(1000)
    try:
        n = 40013
        u1, P1, u2, P2, X = rig.make_correspondences(n, "translating", 0.5)
        xi, sti = tri.iterative_LS_triangulation(u1, P1, u1, P1)
        xo, so, _, margin = orc.iterative_LS_core(u1, P1, u1, P1)
        keep = margin > 1e-9
        assert np.isfinite(xi).all()
        assert np.array_equal(sti[keep], so[keep]) and rel_err(xi, xo)[keep].max() < 1e-8
        xe, ste = tri.linear_eigen_triangulation(u1, P1, u1, P1)          # two-dimensional null space: well-formedness only
        assert xe.shape == (n, 3) and ste.shape == (n,)
        us = np.stack([u1, u1, u1])
        xm, stm = tri.multiview_LS_triangulation(us, [P1, P1, P1])
        xmo, _ = orc.multiview_LS_triangulation(us, [P1, P1, P1])
        assert stm.all() and rel_err(xm, xmo).max() < 1e-8
        u1, P1, u2, P2, X = rig.make_correspondences(n, "forward", 20.0)
        for name in ("polynomial", "linear_eigen"):
            x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
            tc.set_deferred_capacity(old)
            x_full, st_full = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)      # same points through the list
            tc.set_deferred_capacity(1000)
            xo, so = orc.SOLVERS[name](u1, P1, u2, P2)
            if name == "polynomial":
                n1, n2 = orc.correct_matches(orc.fundamental_from_P(P1, P2), u1, u2)
                ok = eigen_well_posed(n1, P1, n2, P2)
            else:
                ok = eigen_well_posed(u1, P1, u2, P2)
            ok &= np.isfinite(xo).all(axis=1)
            assert ok.mean() >= 0.995, name
            assert np.array_equal(st[ok], so[ok]) and np.array_equal(st_full[ok], so[ok]), name
            assert rel_err(x, xo)[ok].max() < TOL64 and rel_err(x_full, xo)[ok].max() < TOL64, name
    finally:
        tc.set_deferred_capacity(old)
    # the list is re-armed: a normal call right after it is unaffected
    u1, P1, u2, P2, X = rig.make_correspondences(30011, "rotating", 0.8)
    x, st = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    xo, so, _, margin = orc.iterative_LS_core(u1, P1, u2, P2)
    keep = margin > 1e-9
    assert np.array_equal(st[keep], so[keep]) and rel_err(x, xo)[keep].max() < TOL64


@pytest.mark.gpu
def test_followup_grid_hint_does_not_change_results(tri):
    """polynomial sizes its follow-up grid from the number of points its previous call on the same stream deferred (a
    page-locked hint, no synchronisation).  First call: default grid; later calls: the hinted one.  Same bits."""
    import triangl_cuda as tc
    n = 300007
    u1, P1, u2, P2, X = rig.make_correspondences(n, "forward", 20.0)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    outs, deferred = [], []
    for call in range(3):
        x = tc.DeviceArray((n, 3), np.float64); st = tc.DeviceArray((n,), np.uint8)
        before = tc.deferred_total()
        tc.polynomial(d1, P1, d2, P2, x=x, status=st, check_all_nan=False)
        deferred.append(tc.deferred_total() - before)
        outs.append((x.to_host(), st.to_host()))
    # enough deferred points for the hint to enlarge the grid beyond 2 CTAs per SM (148 x 2 x 256 = 75776 points)
    assert min(deferred) > 80000 and len(set(deferred)) == 1, deferred
    for x, st in outs[1:]:
        assert np.array_equal(st, outs[0][1]) and np.array_equal(x, outs[0][0], equal_nan=True)
    xo, so = orc.SOLVERS["polynomial"](u1, P1, u2, P2)
    n1, n2 = orc.correct_matches(orc.fundamental_from_P(P1, P2), u1, u2)
    ok = eigen_well_posed(n1, P1, n2, P2) & np.isfinite(xo).all(axis=1)
    assert np.array_equal(outs[2][1].astype(bool)[ok], so[ok]) and rel_err(outs[2][0], xo)[ok].max() < TOL64


@pytest.mark.parametrize("name", ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"])
def test_deferred_list_overflow_with_fused_evaluation(tri, name):
    """List overflow + evaluation epilogue: the follow-up kernel redoes and re-evaluates EVERY point, so the hot kernel's
    own sums must not be counted twice.  Fused sums / mask / errors == the stand-alone pass over the stored result."""
    import triangl_cuda as tc
    n = 60013
    # forward motion under heavy noise defers thousands of points in every solver but linear_LS; a 1e-5 baseline does for it
    if name == "linear_LS":
        u1, P1, u2, P2, _ = rig.make_correspondences(n, (1e-5, 0., 0.), 0.1)
    else:
        u1, P1, u2, P2, _ = rig.make_correspondences(n, "forward", 20.0)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    fn = {"linear_eigen": tc.linear_eigen, "linear_LS": tc.linear_ls, "iterative_LS": tc.iterative_ls,
          "polynomial": lambda *a, **k: tc.polynomial(*a, check_all_nan=False, **k)[:2]}[name]
    thr = (30.0 / 480) ** 2
    if name == "iterative_LS":                   # identical cameras: every system has rank 2, nothing is certified
        d2, u2, P2 = d1, u1, P1
    x_ref, st_ref = fn(d1, P1, d2, P2)
    old = tc.set_deferred_capacity(10)
    try:
        fe = tc.FusedEval(n, np.float64, -5, thr, want_errors=True, want_good=True)
        x, st = fn(d1, P1, d2, P2, evaluate=fe)
        e1, e2, good, sums = tc.pair_reproj(x, d1, P1, d2, P2, st, -5, thr)
        tc.synchronize()
        fs = fe.sums.to_host()
    finally:
        tc.set_deferred_capacity(old)
    if name in ("linear_LS", "iterative_LS"):
        # the follow-up kernel redoes every point with the arithmetic the point had before: same bits
        assert np.array_equal(x.to_host(), x_ref.to_host(), equal_nan=True) and np.array_equal(st.to_host(), st_ref.to_host())
    else:
        # every point through the Jacobi SVD / the complete correction instead of the certified fast paths: same points
        # to rounding, not the same bits
        with np.errstate(all="ignore"):
            assert np.nanmedian(rel_err(x.to_host(), x_ref.to_host())) < 1e-12
        assert (st.to_host() != st_ref.to_host()).mean() < 1e-3
    assert np.array_equal(fe.good.to_host(), good.to_host())
    assert np.array_equal(fe.err1.to_host(), e1.to_host(), equal_nan=True)
    assert sums[3] > 0 and fs[2] == sums[2] and fs[3] == sums[3]          # counts exactly once
    assert fs[0] == pytest.approx(sums[0], rel=1e-10) and fs[1] == pytest.approx(sums[1], rel=1e-10)
    # and the list is re-armed: the same call without the limit gives the same sums
    fe2 = tc.FusedEval(n, np.float64, -5, thr, want_errors=False, want_good=False)
    fn(d1, P1, d2, P2, evaluate=fe2)
    tc.synchronize()
    f2 = fe2.sums.to_host()
    assert f2[2] == sums[2] and f2[3] == sums[3] and f2[0] == pytest.approx(sums[0], rel=1e-10)


def test_iterative_ls_uncertified_points_take_the_reference_loop(tri):
    """Cameras without a finite centre (affine P), identical cameras (rank 2) and a 1e-5 baseline: none of them is
    certified by the closed form, all must still match the oracle."""
    u1, P1, u2, P2, X = rig.make_correspondences(4001, "rotating", 0.5)
    # (a) affine cameras: third row (0,0,0,1) -> no centre; depths are exactly 1, every point stops at the first solve
    Pa1 = P1.copy(); Pa2 = P2.copy()
    Pa1[2] = (0, 0, 0, 1); Pa2[2] = (0, 0, 0, 1)
    x, st = tri.iterative_LS_triangulation(u1, Pa1, u2, Pa2)
    xo, so, _, margin = orc.iterative_LS_core(u1, Pa1, u2, Pa2)
    ok = ls_well_posed(u1, Pa1, u2, Pa2) & (margin > 1e-9)
    assert ok.mean() >= 0.995
    assert np.array_equal(st[ok], so[ok]) and rel_err(x, xo)[ok].max() < TOL64
    # (b) identical cameras: minimum-norm solutions of rank-2 systems
    x, st = tri.iterative_LS_triangulation(u1, P1, u1, P1)
    xo, so, _, margin = orc.iterative_LS_core(u1, P1, u1, P1)
    keep = margin > 1e-9
    assert np.isfinite(x).all()
    assert np.array_equal(st[keep], so[keep]) and rel_err(x, xo)[keep].max() < 1e-8
    # (c) 1e-5 baseline: kappa^2 ~ 1e7..1e10, straddles the certification limit
    u1, P1, u2, P2, X = rig.make_correspondences(20011, (1e-5, 0., 0.), 0.1)
    x, st = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    xo, so, _, margin = orc.iterative_LS_core(u1, P1, u2, P2)
    keep = margin > 1e-9
    assert np.array_equal(st[keep], so[keep])
    A, _ = orc.build_Ab(u1, P1, u2, P2)
    s = np.linalg.svd(A, compute_uv=False)
    assert np.all(rel_err(x, xo)[keep] < np.maximum(TOL64, (s[:, 0] / s[:, 2])[keep] * 1e-13))


@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.0, 0.8, 8.0])
@pytest.mark.parametrize("rows", [4, 6])
def test_linear_eigen_vs_oracle(tri, rig_name, sigma, rows):
    u1, P1, u2, P2, X = rig.make_correspondences(30011, rig_name, sigma)
    tri.set_triangl_semantics(eigen_rows_=rows)
    try:
        x, st = tri.linear_eigen_triangulation(u1, P1, u2, P2)
    finally:
        tri.set_triangl_semantics(eigen_rows_=4)
    xo, so = orc.linear_eigen_triangulation(u1, P1, u2, P2, rows=rows)
    ok = eigen_well_posed(u1, P1, u2, P2)
    assert ok.mean() >= 0.995           # measured: >= 0.9956 on every rig and noise level
    assert np.array_equal(st[ok], so[ok])
    assert rel_err(x, xo)[ok].max() < TOL64


@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.0, 0.8, 8.0, 20.0])
def test_polynomial_vs_oracle(tri, rig_name, sigma):
    u1, P1, u2, P2, X = rig.make_correspondences(20011, rig_name, sigma)
    x, st = tri.polynomial_triangulation(u1, P1, u2, P2)
    xo, so = orc.polynomial_triangulation(u1, P1, u2, P2)
    F = orc.fundamental_from_P(P1, P2)
    c1, c2 = orc.correct_matches(F, u1, u2)
    ok = eigen_well_posed(c1, P1, c2, P2) & np.isfinite(xo).all(axis=1)
    assert ok.mean() >= 0.995           # measured: >= 0.9956 on every rig and noise level
    assert np.array_equal(st[ok], so[ok])
    assert rel_err(x, xo)[ok].max() < TOL64


@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_polynomial_ray_intersection_equals_eigen_solver(tri, rig_name, dtype):
    """After the Hartley-Sturm correction the two viewing rays meet: the certified closed-form intersection (default)
    and cv2.triangulatePoints' smallest singular vector (trgl_set_two_ray(0)) are the same point, same mask."""
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(50021, rig_name, 4.0, dtype=dtype)
    x, st = tri.polynomial_triangulation(u1, P1, u2, P2)
    old = tc.set_two_ray(0)
    try:
        xg, stg = tri.polynomial_triangulation(u1, P1, u2, P2)
    finally:
        tc.set_two_ray(old)
    assert np.array_equal(st, stg)
    n1, n2 = orc.correct_matches(orc.fundamental_from_P(P1, P2), np.asarray(u1, np.float64), np.asarray(u2, np.float64))
    ok = eigen_well_posed(n1, P1, n2, P2) & st
    assert ok.mean() >= 0.995
    assert rel_err(x, xg)[ok].max() < (TOL64 if dtype == np.float64 else TOL32)


@pytest.mark.parametrize("rig_name", RIG_NAMES)
def test_corrected_matches_vs_oracle(tri, rig_name):
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(20011, rig_name, 4.0)
    _, _, all_nan, c1, c2 = tc.polynomial(u1, P1, u2, P2, want_corrected=True)
    F = orc.fundamental_from_P(P1, P2)
    o1, o2 = orc.correct_matches(F, u1, u2)
    assert not all_nan
    assert np.nanmax(np.abs(c1 - o1)) < 1e-12 and np.nanmax(np.abs(c2 - o2)) < 1e-12
    # the corrected matches satisfy the epipolar constraint exactly (the defining property of the method)
    h1 = np.concatenate([c1, np.ones((len(c1), 1))], axis=1); h2 = np.concatenate([c2, np.ones((len(c2), 1))], axis=1)
    assert np.max(np.abs(np.einsum('ni,ij,nj->n', h2, F, h1))) < 1e-12 * np.abs(F).max()


# ---------------------------------------------------------------------------------------------------------------
def test_reference_python_fixtures(tri, golden_dir):
    """Outputs of the reference's own (unmodified, exec'd) Python solvers, committed by oracle/make_golden.py."""
    files = sorted(glob.glob(os.path.join(golden_dir, "ref_py_[0-9]*.npz")))
    assert len(files) >= 5
    tri.set_triangl_semantics('py')
    try:
        for f in files:
            d = np.load(f)
            u1, P1, u2, P2 = d["u1"], d["P1"], d["u2"], d["P2"]
            for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
                x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
                xr, sr = d["x_" + name], d["status_" + name]
                ok = ls_well_posed(u1, P1, u2, P2) if "LS" in name else eigen_well_posed(u1, P1, u2, P2)
                if name == "polynomial":
                    F = orc.fundamental_from_P(P1, P2)
                    c1, c2 = orc.correct_matches(F, u1, u2)
                    ok = eigen_well_posed(c1, P1, c2, P2)
                assert ok.mean() >= 0.985, (f, name)      # 400-point fixtures: the forward rig has 5 ill-posed points
                assert np.array_equal(np.asarray(st)[ok], sr[ok]), (f, name)
                assert rel_err(x, xr)[ok].max() < TOL64, (f, name, rel_err(x, xr)[ok].max())
    finally:
        tri.set_triangl_semantics('c')


def test_slam_convention_fixture(tri, golden_dir):
    """float32 points, 4x4 P, float32 output dtype (slam2.py:19,551-555); reference computes in float64."""
    d = np.load(os.path.join(golden_dir, "ref_py_slam_f32.npz"))
    tri.set_triangl_output_dtype(np.float32)
    tri.set_triangl_semantics('py')
    try:
        x, st = tri.iterative_LS_triangulation(d["u1"], d["P1"], d["u2"], d["P2"])
    finally:
        tri.set_triangl_output_dtype(float)
        tri.set_triangl_semantics('c')
    assert x.dtype == np.float32
    assert np.array_equal(st, d["status_iterative_LS"])
    assert np.max(np.abs(x.astype(np.float64) - d["x_iterative_LS"].astype(np.float64))
                  / np.abs(d["x_iterative_LS"]).max(axis=1, keepdims=True)) < 3e-7       # one float32 ulp-ish


def test_golden_cells_on_gpu(tri, golden_dir):
    """The GPU drop-in reproduces cells of the reference's own golden result files (test_1and2.mat)."""
    with open(os.path.join(golden_dir, "golden_cells.json")) as f:
        g = json.load(f)["test_1and2"]
    tri.set_triangl_semantics('c', 6)
    try:
        for cell in g["cells"]:
            if cell["traj"] == 1:
                continue        # forward trajectory: on-baseline lattice points, see tests/test_oracle_golden.py
            tr = g["trajectories"][cell["traj"]]
            pose = (tr["sideways_values"][cell["pose"]], tr["towards_values"][cell["pose"]], tr["angle_values"][cell["pose"]])
            got = replay_cell([tri.linear_eigen_triangulation, tri.linear_LS_triangulation, tri.iterative_LS_triangulation],
                              pose, num_trials=g["num_trials"], rseed=g["rseed"])
            for k in got:
                for ti in range(3):
                    want = cell[k][ti]
                    if want is not None:
                        assert got[k][ti] == pytest.approx(want, rel=2e-8, abs=1e-12), (cell["traj"], cell["pose"], k, ti)
    finally:
        tri.set_triangl_semantics('c', 4)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rig_name", RIG_NAMES)
@pytest.mark.parametrize("sigma", [0.0, 0.8, 8.0])
@pytest.mark.parametrize("name", ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"])
def test_fp32_mode(tri, name, rig_name, sigma):
    """FP32 mode: float32 storage (and float32 arithmetic where it exists: linear_LS); the comparator is the float64
    reference run on the float32-rounded inputs (the reference is all-float64: triangulation_c/__init__.py:32-33), bar 1e-4.
    Points whose float64 answer is itself ill-posed at the 1e-4 level (conditioning x float32 output rounding) are a
    separate, counted class; status vectors are compared on every other point."""
    from oracle import oracle_c
    n = 30011
    u1, P1, u2, P2, X = rig.make_correspondences(n, rig_name, sigma, dtype=np.float32)
    tri.set_triangl_output_dtype(np.float32)
    tri.set_triangl_compute_dtype(np.float32)
    try:
        x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
    finally:
        tri.set_triangl_output_dtype(float)
        tri.set_triangl_compute_dtype(np.float64)
    w1, w2 = u1.astype(np.float64), u2.astype(np.float64)
    assert x.dtype == np.float32
    flip = np.zeros(n, dtype=bool)
    if name == "linear_LS":
        xo, so = oracle_c.linear_LS_triangulation(w1, P1, w2, P2)
        well = oracle_c.ls_condition(w1, P1, w2, P2) * 2.2e-16 * 50 < TOL64
    elif name == "iterative_LS":
        xo, so, margin = oracle_c.iterative_LS_triangulation(w1, P1, w2, P2, return_margin=True)
        well = margin > 1e-9
    elif name == "linear_eigen":
        xo, so, amp = oracle_c.linear_eigen_triangulation(w1, P1, w2, P2, return_amp=True)
        well = amp * 2.2e-16 * 50 < TOL64
    else:
        # cv2.correctMatches returns the dtype of its input: the corrected match is rounded to float32
        # before the triangulation (triangulation.py:224,232) -- in the oracle as in the kernel
        c1, c2 = oracle_c.correct_matches(orc.fundamental_from_P(P1, P2), w1, w2)
        c1 = c1.astype(np.float32).astype(np.float64); c2 = c2.astype(np.float32).astype(np.float64)
        xo, so, amp = oracle_c.linear_eigen_triangulation(c1, P1, c2, P2, return_amp=True)
        well = (amp * 2.2e-16 * 50 < TOL64) & np.isfinite(xo).all(axis=1)
        # a last-bit difference in the float64 correction can flip that float32 rounding: one float32 ulp of the match
        # moves the point by amp x 6e-8 -- counted separately where that exceeds the bar (rare: measured 0 per 30 k points)
        flip = amp * 1.2e-7 >= TOL32
    rel = rel_err(x, xo)
    mism = np.asarray(st) != np.asarray(so)
    from conftest import PARITY_REPORTS
    PARITY_REPORTS.append("fp32 %-13s %-11s sigma %4.1f: n %d, separate class %d, status mismatches %d (outside the class %d), "
                          "max rel err outside the class %.2e" % (name, rig_name, sigma, n, int((~well).sum()), int(mism.sum()),
                                                                  int((mism & well).sum()), rel[well].max()))
    assert well.mean() >= 0.99
    assert not (mism & well).any()
    assert rel[well & ~flip].max() < TOL32
    assert (rel[well & flip] >= TOL32).sum() <= 1e-3 * n


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 257, 1000])
def test_ragged_sizes(tri, n):
    u1, P1, u2, P2, X = rig.make_correspondences(max(n, 1), "general", 0.8)
    u1, u2 = u1[:n], u2[:n]
    for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
        x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
        assert x.shape == (n, 3) and st.shape == (n,)
        if n:
            xo, so = orc.SOLVERS[name](u1, P1, u2, P2)
            assert np.array_equal(st, so) and rel_err(x, xo).max() < TOL64


def test_single_point_identity_P(tri):
    """calibrate.py:337-339 calls iterative_LS with P1 = eye(4) on one point."""
    P1 = np.eye(4)
    P2 = rig.P_from_R_and_t(rig.rot_y(0.1), [-1.0, 0.0, 0.2])
    Xw = np.array([[0.3, -0.2, 5.0, 1.0]])
    u1 = (Xw @ P1[0:3].T); u1 = u1[:, 0:2] / u1[:, 2:3]
    u2 = (Xw @ P2[0:3].T); u2 = u2[:, 0:2] / u2[:, 2:3]
    for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
        x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
        assert np.allclose(x, Xw[:, 0:3], rtol=1e-10, atol=1e-12), name
        assert st[0] == 1


def test_nan_inf_inputs_propagate(tri):
    u1, P1, u2, P2, X = rig.make_correspondences(64, "rotating", 0.8)
    u1 = u1.copy(); u1[3, 0] = np.nan; u1[7, 1] = np.inf
    for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
        x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
        xo, so = orc.SOLVERS[name](u1, P1, u2, P2)
        assert not np.isfinite(x[3]).all() and not np.isfinite(x[7]).all(), name
        assert np.array_equal(st, so), (name, st[[3, 7]], so[[3, 7]])
        keep = np.ones(64, bool); keep[[3, 7]] = False
        assert rel_err(x, xo)[keep].max() < TOL64


def test_rank_deficient_min_norm(tri):
    """Identical cameras: A has rank 2, cvSolve(DECOMP_SVD) returns the minimum-norm solution."""
    u1, P1, u2, P2, X = rig.make_correspondences(500, "translating", 0.0)
    x, st = tri.linear_LS_triangulation(u1, P1, u1, P1)
    xo, _ = orc.linear_LS_triangulation(u1, P1, u1, P1)
    assert np.isfinite(x).all()
    assert rel_err(x, xo).max() < 1e-9


def test_polynomial_all_nan_fallback(tri):
    """F == 0 (identical cameras) -> every corrected point NaN -> 8-point F from the matches (triangulation.py:227-229)."""
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(2000, "general", 0.5)
    _, _, all_nan = tc.polynomial(u1, P1, u2, P1)           # P2 == P1 -> F == 0
    assert all_nan
    F = tc.fundamental_8point(u1, u2)
    Fo = orc.find_fundamental_8point(u1, u2)
    F /= np.linalg.norm(F); Fo /= np.linalg.norm(Fo)
    assert min(np.abs(F - Fo).max(), np.abs(F + Fo).max()) < 1e-8
    x, st = tri.polynomial_triangulation(u1, P1, u2, P1)    # takes the fallback internally, must not raise
    assert x.shape == (2000, 3)


def test_device_resident_api(tri):
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(100000, "rotating", 0.8)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    xh, sh = tri.linear_LS_triangulation(u1, P1, u2, P2)
    xd, sd = tri.linear_LS_triangulation(d1, P1, d2, P2)
    tc.synchronize()
    assert isinstance(xd, tc.DeviceArray)
    assert np.array_equal(xd.to_host(), xh) and np.array_equal(sd.to_host(), sh)
    for ppt in (1, 2, 4):
        old = tc.set_points_per_thread(ppt)
        xp, _ = tc.linear_ls(d1, P1, d2, P2)
        tc.synchronize()
        assert np.array_equal(xp.to_host(), xh)
        tc.set_points_per_thread(old)


def test_torch_tensor_frontend(tri):
    torch = pytest.importorskip("torch")
    u1, P1, u2, P2, X = rig.make_correspondences(10000, "general", 0.8)
    t1 = torch.from_numpy(u1).cuda(); t2 = torch.from_numpy(u2).cuda()
    xo = torch.empty((len(u1), 3), dtype=torch.float64, device="cuda")
    so = torch.empty(len(u1), dtype=torch.uint8, device="cuda")
    import triangl_cuda as tc
    torch.cuda.synchronize()
    tc.linear_ls(t1, P1, t2, P2, x=xo, status=so, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    xh, _ = tri.linear_LS_triangulation(u1, P1, u2, P2)
    assert np.array_equal(xo.cpu().numpy(), xh) and bool(so.all())


def test_pinned_host_buffers(tri):
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(300000, "general", 0.8)
    p1, p2 = tc.pinned_copy(u1), tc.pinned_copy(u2)
    xa, _ = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    xb, _ = tri.iterative_LS_triangulation(p1, P1, p2, P2)
    assert np.array_equal(xa, xb)
    # linear_LS through the chunked host pipeline (1.2 M points = 3 chunks) into a caller-provided status array
    u1, P1, u2, P2, X = rig.make_correspondences(1_200_007, "rotating", 0.8)
    st = np.zeros(len(u1), dtype=np.bool_)
    x, st2 = tc.linear_ls(u1, P1, u2, P2, status=st)
    xo, so = orc.linear_LS_triangulation(u1[::97], P1, u2[::97], P2)
    assert st2 is st and st.all() and so.all()
    assert rel_err(x[::97], xo).max() < TOL64


# ---------------------------------------------------------------------------------------------------------------
def test_reprojection_error_vs_oracle(tri):
    import calibration_tools as ct
    u1, P1, u2, P2, X = rig.make_correspondences(50000, "rotating", 0.8)
    K = np.array([[480., 0, 320], [0, 470., 240], [0, 0, 1]])
    dist = np.array([0.1, -0.05, 0.001, -0.002, 0.01])
    rvec = np.array([0.02, 0.29, -0.01]); tvec = np.array([-1.0, 0.1, 40.0])
    imgp = orc.project_points(X, rvec, tvec, K, dist) + np.random.RandomState(1).normal(0, 0.5, (len(X), 2))
    rms, proj = ct.reprojection_error(X, imgp, K, dist, rvec, tvec)
    rmso, projo = orc.reprojection_error(X, imgp, K, dist, rvec, tvec)
    assert proj.shape == (len(X), 1, 2)
    assert abs(rms - rmso) < 1e-12 * rmso and np.max(np.abs(proj - projo)) < 1e-10
    m, s = ct.reprojection_error_ext([X, X[:100]], [imgp, imgp[:100]], K, dist, [rvec, rvec], [tvec, tvec])
    mo, so = orc.reprojection_error_ext([X, X[:100]], [imgp, imgp[:100]], K, dist, [rvec, rvec], [tvec, tvec])
    assert abs(m - mo) < 1e-12 * mo and abs(s - so) < 1e-12 * so
    # float32 object points (the SLAM convention)
    rms32, _ = ct.reprojection_error(X.astype(np.float32), imgp.astype(np.float32), K, dist, rvec, tvec)
    assert abs(rms32 - rmso) < 1e-4 * rmso


def test_fused_pair_reprojection_and_good_mask(tri):
    u1, P1, u2, P2, X = rig.make_correspondences(50000, "rotating", 2.0)
    x, st, good, rms = tri.triangulate_and_evaluate(tri.iterative_LS_triangulation, u1, P1, u2, P2, min_status=0,
                                                    max_sq_err=(2.0 / 480) ** 2)
    xh = np.concatenate([x, np.ones((len(x), 1))], axis=1)
    e = []
    for u, P in ((u1, P1), (u2, P2)):
        p = xh @ P.T
        e.append(((p[:, 0:2] / p[:, 2:3] - u) ** 2).sum(axis=1))
    want = (st > 0) & (e[0] <= (2.0 / 480) ** 2) & (e[1] <= (2.0 / 480) ** 2)
    margin = np.minimum(np.abs(e[0] - (2.0 / 480) ** 2), np.abs(e[1] - (2.0 / 480) ** 2)) > 1e-15
    assert np.array_equal(good[margin], want[margin])
    assert rms[0] == pytest.approx(np.sqrt(e[0][want].mean()), rel=1e-9)
    assert rms[1] == pytest.approx(np.sqrt(e[1][want].mean()), rel=1e-9)


# ---------------------------------------------------------------------------------------------------------------
def _classify(name, w1, P1, w2, P2):
    """Oracle result + the class of points no two correct implementations agree on, from the C oracle."""
    from oracle import oracle_c
    if name == "linear_LS":
        xo, so = oracle_c.linear_LS_triangulation(w1, P1, w2, P2)
        special = oracle_c.ls_condition(w1, P1, w2, P2) * 2.2e-16 * 50 >= TOL64
    elif name == "iterative_LS":
        xo, so, margin = oracle_c.iterative_LS_triangulation(w1, P1, w2, P2, return_margin=True)
        special = (margin < 1e-9) | (oracle_c.ls_condition(w1, P1, w2, P2) * 2.2e-16 * 50 >= TOL64)
    elif name == "linear_eigen":
        xo, so, amp = oracle_c.linear_eigen_triangulation(w1, P1, w2, P2, return_amp=True)
        special = ~(amp * 2.2e-16 * 50 < TOL64)
    else:
        xo, so, amp = oracle_c.polynomial_triangulation(w1, P1, w2, P2, return_amp=True)
        special = ~(amp * 2.2e-16 * 50 < TOL64) | ~np.isfinite(xo).all(axis=1)
    return xo, np.asarray(so), special


@pytest.mark.parametrize("rig_name,tiled", [("rotating", True), ("rotating", False), ("forward", False)])
def test_bench_input_parity(tri, rig_name, tiled):
    """
    Parity on what is benchmarked (BASELINE configs[1]): 10 M correspondences, 0.8 px noise, all four solvers, device-resident
    arrays, the evaluation fused / separate exactly as bench.py's timed step runs it -- against the C oracle
    (triangulation.c:104-161 / triangulation.py:6-25,198-232 restated; pinned to the NumPy oracle and through it to the
    reference's golden files) on EVERY point.
      tiled = True : the very arrays bench.py times (harness/synthetic_rig.bench_batch: 2 M seeded points tiled to 10 M);
                     the oracle runs on the 2 M distinct points, every tile of the GPU result is compared with it;
      tiled = False: 10 M distinct points (rotating and forward-motion rigs).
    Nothing is filtered: status mismatches and the largest relative error are counted over all points and reported with the
    knife-edge / ill-posed class separated; outside that class mismatches must be 0 and errors <= 1e-9.
    """
    import triangl_cuda as tc
    from conftest import PARITY_REPORTS
    n = 10_000_000
    if tiled:
        u1, P1, u2, P2, base = rig.bench_batch(n, rig_name, rank=0)
    else:
        u1, P1, u2, P2, _ = rig.make_correspondences(n, rig_name, sigma=0.8, seed=rig.RSEED + 17)
        base = n
    d_u1, d_u2 = tc.to_device(u1), tc.to_device(u2)
    thr = (2.0 / 480) ** 2
    for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
        fe = None
        if name != "linear_LS":                  # bench.py default: epilogue for the three FP64-bound solvers ...
            fe = tc.FusedEval(n, np.float64, 0, thr, want_errors=False, want_good=True)
        if name == "linear_eigen":
            x, st = tc.linear_eigen(d_u1, P1, d_u2, P2, evaluate=fe)
        elif name == "linear_LS":
            x, st = tc.linear_ls(d_u1, P1, d_u2, P2)
        elif name == "iterative_LS":
            x, st = tc.iterative_ls(d_u1, P1, d_u2, P2, evaluate=fe)
        else:
            x, st, all_nan = tc.polynomial(d_u1, P1, d_u2, P2, evaluate=fe)
            assert not all_nan
        if fe is None:                           # ... and the stand-alone pass after linear_LS
            _, _, good, sums = tc.pair_reproj(x, d_u1, P1, d_u2, P2, st, 0, thr, want_errors=False, want_good=True)
        else:
            good, sums = fe.good, fe.sums.to_host()
        tc.synchronize()
        x, st, good = x.to_host(), st.to_host(), good.to_host()
        xo, so, special = _classify(name, u1[:base], P1, u2[:base], P2)
        if base < n:                             # every tile against the oracle on the distinct points
            reps = -(-n // base)
            xo = np.tile(xo, (reps, 1))[:n]; so = np.tile(so, reps)[:n]; special = np.tile(special, reps)[:n]
        rel = rel_err(x, xo)
        with np.errstate(invalid="ignore"):
            bad = ~(rel <= TOL64) & ~(np.isnan(x).all(axis=1) & np.isnan(xo).all(axis=1))
        mism = st.astype(np.int64) != so.astype(np.int64)
        # good mask of the harness / SLAM step from the ORACLE's x (triangulation_comparison.py:190-217; slam2.py:556,589)
        xh = np.concatenate([xo, np.ones((n, 1))], axis=1)
        e = []
        dep = []
        with np.errstate(all="ignore"):
            for u, P in ((u1, P1), (u2, P2)):
                pr = xh @ P.T
                e.append(((pr[:, 0:2] / pr[:, 2:3] - u) ** 2).sum(axis=1)); dep.append(pr[:, 2])
            want = (so > 0) & (e[0] <= thr) & (e[1] <= thr) & (dep[0] > 0) & (dep[1] > 0)
            edge = (np.minimum(np.abs(e[0] - thr), np.abs(e[1] - thr)) <= 1e-9 * thr) | special
        gm = good.astype(bool) != want
        PARITY_REPORTS.append(
            "bench-input %-13s %-8s %s n %d: separate class %d (%.1e of all), status mismatches %d (outside the class %d), "
            "points over 1e-9 %d (outside the class %d), max rel err outside the class %.2e, good-mask mismatches %d "
            "(outside the class / threshold edge %d), good %d" %
            (name, rig_name, "tiled-2M" if tiled else "distinct", n, int(special.sum()), special.mean(), int(mism.sum()),
             int((mism & ~special).sum()), int(bad.sum()), int((bad & ~special).sum()), np.nanmax(rel[~special]),
             int(gm.sum()), int((gm & ~edge).sum()), int(good.sum())))
        assert special.mean() < 5e-3, name
        assert not (mism & ~special).any(), name
        assert not (bad & ~special).any(), name
        assert not (gm & ~edge).any(), name
        assert sums[2] == good.sum(), name               # the fused count is the mask's population
        # the ill-posed class is bounded too: finite where the oracle is finite, status mismatches only inside it
        assert mism.sum() <= special.sum(), name


def test_full_size_properties(tri):
    """BASELINE config-2 size (10 M points): size-independent properties instead of an oracle run.
       (1) exact projections triangulate back to the cloud (round trip) for every solver;
       (2) shard-invariance: solving a slice equals slicing the solution (what multi-GPU sharding relies on)."""
    n = 10_000_000
    rng = np.random.RandomState(7)
    u1s, P1, u2s, P2, Xs = rig.make_correspondences(1_000_000, "rotating", 0.0, seed=99)
    reps = n // len(u1s)
    u1 = np.tile(u1s, (reps, 1)); u2 = np.tile(u2s, (reps, 1)); X = np.tile(Xs, (reps, 1))
    for name in ("linear_LS", "iterative_LS", "linear_eigen", "polynomial"):
        x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
        assert x.shape == (n, 3)
        err = np.max(np.abs(x - X), axis=1)
        assert err.max() < 1e-9 * 44, (name, err.max())
        assert (st == 1).all(), name
        lo = int(rng.randint(0, n - 70000)); hi = lo + 65537
        xs, ss = getattr(tri, name + "_triangulation")(u1[lo:hi], P1, u2[lo:hi], P2)
        assert np.array_equal(xs, x[lo:hi]) and np.array_equal(ss, st[lo:hi]), name


# ---- input normalisation fused in front of the solvers -----------------------------------------
def _pixel_batch(n, rig_name="rotating", dtype=np.float64, seed=5):
    """Pixel observations of the synthetic rig with a distorting camera: normalised noisy points pushed through the
    forward distortion model + K (what a real detector would hand to cv2.undistortPoints)."""
    u1, P1, u2, P2, X = rig.make_correspondences(n, rig_name, sigma=0.8, seed=rig.RSEED + seed)
    K = np.array([[480., 0, 320], [0, 480., 240], [0, 0, 1]])
    dist = np.array([-0.28, 0.07, 2e-4, -1e-4, 0.01])

    def to_px(u):
        z = np.zeros(3)
        return orc.project_points(np.column_stack([u, np.ones(len(u))]), z, z, K, dist)
    return to_px(u1).astype(dtype), P1, to_px(u2).astype(dtype), P2, K, dist


def test_undistort_kernel_bit_identical_to_cv2_fixture(tri, golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_undistort.npz"))
    px, K = g["px"], g["K"]
    for k, d in enumerate(g["dists"]):
        out = tri.undistort_points(px, K, d)
        assert out.dtype == np.float64 and np.array_equal(out, g["n64_%d" % k]), "float64 model %d" % k
        out32 = tri.undistort_points(px.astype(np.float32), K, d)
        assert out32.dtype == np.float32 and np.array_equal(out32, g["n32_%d" % k]), "float32 model %d" % k
    assert np.array_equal(tri.undistort_points(px, K, None), g["n64_none"])
    assert np.array_equal(tri.undistort_points(px.reshape(1, -1, 2), K, g["dists"][0]), g["n64_0"])   # cv2-style shape
    assert tri.undistort_points(np.zeros((0, 2)), K, None).shape == (0, 2)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"])
def test_pixel_input_solvers_equal_undistort_then_solve(tri, name, dtype):
    """*_px == undistort kernel followed by the plain solver, bit for bit (x and status), and both agree with the
    oracle chain cv2.undistortPoints-restatement -> solver restatement within the FP64 bar."""
    px1, P1, px2, P2, K, dist = _pixel_batch(20011, dtype=dtype)
    tri.set_triangl_output_dtype(dtype)
    try:
        xf, sf = getattr(tri, name + "_triangulation_px")(px1, P1, px2, P2, K, dist)
        u1 = tri.undistort_points(px1, K, dist); u2 = tri.undistort_points(px2, K, dist)
        xs, ss = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
    finally:
        tri.set_triangl_output_dtype(float)
    assert xf.dtype == dtype
    assert np.array_equal(np.asarray(sf), np.asarray(ss))
    assert np.array_equal(xf, xs, equal_nan=True)
    o1 = orc.undistort_points(px1, K, dist); o2 = orc.undistort_points(px2, K, dist)
    assert np.array_equal(u1, o1) and np.array_equal(u2, o2)
    xo, so = orc.SOLVERS[name](o1, P1, o2, P2)
    ok = eigen_well_posed(o1.astype(np.float64), P1, o2.astype(np.float64), P2) if name in ("linear_eigen", "polynomial") \
        else np.ones(len(xo), bool)
    if name == "iterative_LS":
        ok &= iterative_margin(o1, P1, o2, P2) > 1e-9
    assert np.array_equal(np.asarray(sf)[ok], np.asarray(so)[ok])
    tol = TOL64 if dtype == np.float64 else 1e-6          # float32 storage of x: 6e-8 rounding
    assert rel_err(xf, xo)[ok].max() < tol


def test_pixel_input_two_cameras_and_no_distortion(tri):
    px1, P1, px2, P2, K, dist = _pixel_batch(4099)
    K2 = K.copy(); K2[0, 0] = 500.; K2[0, 2] = 300.
    x, st = tri.iterative_LS_triangulation_px(px1, P1, px2, P2, K, dist, cameraMatrix2=K2, distCoeffs2=None)
    xo, so = orc.iterative_LS_triangulation(orc.undistort_points(px1, K, dist), P1, orc.undistort_points(px2, K2, None), P2)
    ok = iterative_margin(orc.undistort_points(px1, K, dist), P1, orc.undistort_points(px2, K2, None), P2) > 1e-9
    assert np.array_equal(st[ok], so[ok]) and rel_err(x, xo)[ok].max() < TOL64
    # no distortion at all: (p - c) / f shortcut of the harness (triangulation_comparison.py:168-172) up to 1 ulp
    u = tri.undistort_points(px1, K, None)
    assert np.allclose(u, (px1 - K[0:2, 2]) / 480., rtol=0, atol=1e-15)


# ---- device-side harness statistics -------------------------------------------------------------
def test_device_median_is_np_median(tri):
    import triangl_cuda as tc
    rng = np.random.RandomState(9)
    cases = [rng.rand(1), rng.rand(2), rng.rand(1001) ** 4, rng.rand(1000) * 1e-12, np.repeat(rng.rand(7), 150),
             np.r_[rng.rand(500), np.inf, np.inf], np.zeros(64), rng.rand(300000), np.r_[rng.rand(10), np.nan]]
    for v in cases:
        want = np.median(v)
        for arr in (v, tc.to_device(v)):
            got = tc.median(arr)
            assert (np.isnan(want) and np.isnan(got)) or got == want, (len(v), got, want)
    assert np.isnan(tc.median(np.zeros(0)))
    with pytest.raises(tc.TrianglCudaError):
        tc.median(np.array([1.0, -2.0, 3.0]))


def test_device_error_statistics_match_numpy(tri):
    import harness_stats as hs
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(50021, "rotating", sigma=4.0)
    x, st = tri.iterative_LS_triangulation(u1, P1, u2, P2)
    errors = np.sum((x - X[:, 0:3]) ** 2, axis=1)
    rms, rmed, e_dev = hs.error_rms_3D(X, x)
    assert np.array_equal(e_dev, errors)                                       # same operation order as numpy
    assert rmed == np.sqrt(np.median(errors)) and rms == pytest.approx(np.sqrt(np.mean(errors)), rel=1e-12)
    hs.robustness_thresh_max = hs.robustness_thresh_min = float(np.percentile(errors, 80))
    try:
        fp, fn = hs.robustness_stat_3D(X, x, st)
        est = st > 0
        assert fp == np.mean(~(errors <= hs.robustness_thresh_max) & est) and fn == np.mean((errors <= hs.robustness_thresh_min) & ~est)
        assert fp > 0 and fn > 0
        # device-resident route gives the same numbers
        dx, dst = tc.iterative_ls(tc.to_device(u1), P1, tc.to_device(u2), P2)
        fp2, fn2 = hs.robustness_stat_3D(tc.to_device(X), dx, dst)
        assert (fp2, fn2) == (fp, fn)
    finally:
        hs.robustness_thresh_max = hs.robustness_thresh_min = 1.0


def test_golden_cells_device_resident(tri, golden_dir):
    """Cells of the reference's golden test_1and2.mat with everything after the RNG on the GPU: cv2.undistortPoints
    (k1 = 0.3), the three solvers, the 3-D / 2-D error evaluation, medians and robustness ratios."""
    import harness_stats as hs
    import triangl_cuda as tc
    with open(os.path.join(golden_dir, "golden_cells.json")) as f:
        g = json.load(f)["test_1and2"]
    points_3D = rig.finite_3D_points(4)
    for cell in g["cells"]:
        if cell["traj"] == 1 or cell["pose"] not in (20, 39):
            continue
        tr = g["trajectories"][cell["traj"]]
        pose = (tr["sideways_values"][cell["pose"]], tr["towards_values"][cell["pose"]], tr["angle_values"][cell["pose"]])
        cam1, cam2 = rig.Camera(), rig.Camera()
        cam1.camera_pose(40.); cam2.camera_pose(40., *pose)
        cams = []
        for cam, ang, side, tow in ((cam1, 0., 0., 0.), (cam2, pose[2], pose[0], pose[1])):
            cam.camera_intrinsics((640, 480), 0.3)
            cam.project_points(points_3D)
            cams.append(dict(K=cam.K, dist=cam.dist_coeffs, rvec=np.array([0., ang, 0.]), tvec=cam.P[:, 3].copy(),
                             points_2D_exact=cam.points_2D_exact))
        solvers = [lambda a, b: tc.linear_eigen(a, cam1.P, b, cam2.P, rows=6), lambda a, b: tc.linear_ls(a, cam1.P, b, cam2.P),
                   lambda a, b: tc.iterative_ls(a, cam1.P, b, cam2.P)]
        last_pose = cell["pose"] == 39 and cell["traj"] in (0, 3, 4)      # the reference runs vector_stat there (:483-487)
        acc = [hs.CellStatistics(points_3D, cams, g["num_trials"], keep_vectors=last_pose) for _ in solvers]
        np.random.seed(g["rseed"])
        for _ in range(g["num_trials"]):
            cam1.apply_noise(0.8, True); cam2.apply_noise(0.8, True)
            d1 = tc.undistort_points(tc.to_device(cam1.points_2D), cam1.K, cam1.dist_coeffs)
            d2 = tc.undistort_points(tc.to_device(cam2.points_2D), cam2.K, cam2.dist_coeffs)
            for a, solve in zip(acc, solvers):
                x, st = solve(d1, d2)
                a.add_trial(x, st)
        keys = ("err3D_mean_summary", "err3D_median_summary", "err2D_mean_summary", "err2D_median_summary",
                "false_pos_summary", "false_neg_summary")
        for ti, a in enumerate(acc):
            for k, got in zip(keys, a.summary()):
                want = cell[k][ti]
                if want is not None:
                    assert got == pytest.approx(want, rel=2e-8, abs=1e-12), (cell["traj"], cell["pose"], k, ti)
        if last_pose:
            # vector_stat on the device against the reference's own stored results (p_err3Dv_*_summary of test_1and2.mat)
            vs = np.load(os.path.join(golden_dir, "vector_stat_cells.npz"))
            for ti, name in ((1, "linear_LS_triangulation"), (2, "iterative_LS_triangulation")):
                means, covars = acc[ti].vector_stat()
                gm, gc = vs["mean_%d_%s" % (cell["traj"], name)], vs["covar_%d_%s" % (cell["traj"], name)]
                assert means.shape == (257, 3) and covars.shape == (257, 3, 3)
                scale = np.sqrt(np.trace(gc, axis1=1, axis2=2))[:, None]              # per-point error spread
                assert np.nanmax(np.abs(means - gm) / scale) < 1e-8, (cell["traj"], name)
                assert np.nanmax(np.abs(covars - gc) / (scale ** 2)[:, :, None]) < 1e-8, (cell["traj"], name)
                assert np.array_equal(np.isnan(covars), np.isnan(gc))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_vector_stat_matches_numpy(tri, dtype):
    """trgl_vector_stat (triangulation_comparison.py:219-240) against the NumPy statement, host and device arrays, ragged n."""
    import harness_stats as hs
    import triangl_cuda as tc
    from harness_replay import vector_stat
    rng = np.random.RandomState(11)
    trials, n = 37, 10007
    exact = np.concatenate([rng.normal(0, 3, (n, 3)), np.ones((n, 1))], axis=1)
    x = (exact[None, :, 0:3] + rng.normal(0, 0.1, (trials, n, 3)) * rng.uniform(0.1, 5, (1, n, 1))).astype(dtype)
    x[3, 17] = np.nan
    wm, wc = vector_stat(x.astype(np.float64) - exact[None, :, 0:3])
    for dev in (False, True):
        if dev:
            m, c = hs.vector_stat(tc.to_device(exact), tc.to_device(x))
            tc.synchronize()
            m, c = m.to_host(), c.to_host()
        else:
            m, c = hs.vector_stat(exact, x)
        assert m.shape == (n, 3) and c.shape == (n, 3, 3)
        assert np.isnan(m[17]).all() and np.isnan(c[17]).all()
        keep = np.arange(n) != 17
        assert np.allclose(m[keep], wm[keep], rtol=1e-10, atol=1e-12) and np.allclose(c[keep], wc[keep], rtol=1e-10, atol=1e-14)
        assert np.array_equal(c, np.transpose(c, (0, 2, 1)), equal_nan=True)


def test_async_pair_reprojection_matches_synchronous(tri):
    """trgl_pair_reproj_async finishes the sums inside the kernel (last-block reduction in block order): same bits as the
    host-summed synchronous variant, launch after launch (the ticket counter resets itself)."""
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(123457, "rotating", sigma=0.8)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    x, st = tc.iterative_ls(d1, P1, d2, P2)
    _, _, good_s, sums = tc.pair_reproj(x, d1, P1, d2, P2, st, 0, (2.0 / 480) ** 2)
    dsum = tc.DeviceArray((4,), np.float64)
    for _ in range(3):
        _, _, good_a, out = tc.pair_reproj(x, d1, P1, d2, P2, st, 0, (2.0 / 480) ** 2, sums_device=dsum)
        tc.synchronize()
        assert np.array_equal(out.to_host(), sums)
        assert np.array_equal(good_a.to_host(), good_s.to_host())
    assert sums[2] > 1000


def test_device_mode_is_reentrant_per_stream(tri):
    """Four host threads, each on its own stream, hammer the entry points that use reduction scratch (polynomial NaN
    flags, pair_reproj sums, median): scratch is per (device, stream), so every thread must get the single-thread answer."""
    import ctypes
    import threading
    import triangl_cuda as tc
    u1, P1, u2, P2, X = rig.make_correspondences(200003, "general", sigma=0.8)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    x0, st0, nan0 = tc.polynomial(d1, P1, d2, P2)
    tc.synchronize()
    _, _, _, sums0 = tc.pair_reproj(x0, d1, P1, d2, P2, st0, 0, 1e-5, want_errors=False, want_good=False)
    e0, _ = tc.eval_errors_3d(x0, tc.to_device(X))
    med0 = tc.median(e0)
    x0h = x0.to_host()
    errors = []

    def worker(k):
        try:
            s = ctypes.c_void_p()
            tc.check(tc.lib().trgl_stream_create(ctypes.byref(s)))
            dX = tc.to_device(X)
            for _ in range(15):
                x, st, nan = tc.polynomial(d1, P1, d2, P2, stream=s)
                _, _, _, sums = tc.pair_reproj(x, d1, P1, d2, P2, st, 0, 1e-5, want_errors=False, want_good=False, stream=s)
                e, _ = tc.eval_errors_3d(x, dX, stream=s)
                med = tc.median(e, stream=s)
                assert nan == nan0 and np.array_equal(sums, sums0) and med == med0
            tc.check(tc.lib().trgl_stream_synchronize(s))
            assert np.array_equal(x.to_host(), x0h, equal_nan=True)
            tc.check(tc.lib().trgl_stream_destroy(s))
        except Exception as exc:        # noqa: BLE001
            errors.append((k, repr(exc)))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ---- result mirrors: the multi-GPU gather fused into the solver stores -------------------------------------------------
@pytest.mark.parametrize("name", ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"])
def test_result_mirrors_receive_every_store(tri, name):
    """Two mirror buffers (here on the same GPU; across GPUs they are IPC-mapped peer memory) must end up bit-identical to
    the primary output, for every solver, at a ragged size, at an offset inside a larger 'gathered' array."""
    import triangl_cuda as tc
    n, lo, total = 70001, 12345, 100000
    u1, P1, u2, P2, _ = rig.make_correspondences(n, "rotating", sigma=0.8)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    sdt = np.int32 if name == "iterative_LS" else np.uint8
    fn = {"linear_eigen": tc.linear_eigen, "linear_LS": tc.linear_ls, "iterative_LS": tc.iterative_ls,
          "polynomial": lambda *a, **k: tc.polynomial(*a, check_all_nan=False, **k)[:2]}[name]
    gathered = [(tc.DeviceArray((total, 3), np.float64), tc.DeviceArray((total,), sdt)) for _ in range(2)]
    for gx, gs in gathered:
        tc.check(tc.lib().trgl_memset_d(gx.ptr, 0xff, gx.nbytes, None)); tc.check(tc.lib().trgl_memset_d(gs.ptr, 0x7f, gs.nbytes, None))
    tc.set_result_mirrors([(gx.ptr + lo * 24, gs.ptr + lo * np.dtype(sdt).itemsize) for gx, gs in gathered])
    x, st = fn(d1, P1, d2, P2)
    x2, st2 = fn(d1, P1, d2, P2)                       # the table is consumed by ONE call: this one has no mirrors
    tc.synchronize()
    xh, sth = x.to_host(), st.to_host()
    assert np.array_equal(x2.to_host(), xh, equal_nan=True)
    for gx, gs in gathered:
        gxh, gsh = gx.to_host(), gs.to_host()
        assert np.array_equal(gxh[lo:lo + n], xh, equal_nan=True) and np.array_equal(gsh[lo:lo + n], sth)
        assert np.isnan(gxh[:lo]).all() and np.isnan(gxh[lo + n:]).all()          # nothing outside the shard was touched
        assert (gsh[:lo].view(np.uint8) == 0x7f).all() and (gsh[lo + n:].view(np.uint8) == 0x7f).all()
    with pytest.raises(tc.TrianglCudaError):            # host buffers cannot be mirrored
        tc.set_result_mirrors([(gathered[0][0].ptr, gathered[0][1].ptr)])
        tri.linear_LS_triangulation(u1, P1, u2, P2)


def _ipc_child(handles, n, lo, q):
    import numpy as np
    import synthetic_rig as rig
    import triangl_cuda as tc
    try:
        px, ps = tc.ipc_import(handles[0]), tc.ipc_import(handles[1])
        u1, P1, u2, P2, _ = rig.make_correspondences(n, "rotating", sigma=0.8)
        tc.set_result_mirrors([(px + lo * 24, ps + lo * 4)])
        x, st = tc.iterative_ls(tc.to_device(u1), P1, tc.to_device(u2), P2)
        tc.synchronize()
        q.put(("ok", x.to_host(), st.to_host()))
        tc.ipc_close(px); tc.ipc_close(ps)
    except Exception as exc:        # noqa: BLE001
        q.put(("error", repr(exc), None))


def test_result_mirrors_across_processes_via_cuda_ipc(tri):
    """The exporter / importer pair of the multi-GPU gather: another PROCESS maps this process's gathered arrays through a
    CUDA IPC handle and its solver kernel stores its shard into them (same GPU here; peer GPUs in bench.py --p2p-gather)."""
    import multiprocessing as mp
    import triangl_cuda as tc
    n, lo, total = 30001, 5000, 40000
    gx, gs = tc.DeviceArray((total, 3), np.float64), tc.DeviceArray((total,), np.int32)
    tc.check(tc.lib().trgl_memset_d(gx.ptr, 0xff, gx.nbytes, None)); tc.check(tc.lib().trgl_memset_d(gs.ptr, 0x7f, gs.nbytes, None))
    tc.synchronize()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_ipc_child, args=((tc.ipc_export(gx), tc.ipc_export(gs)), n, lo, q))
    p.start()
    tag, xc, stc = q.get(timeout=120)
    p.join(timeout=60)
    assert tag == "ok", xc
    tc.synchronize()
    assert np.array_equal(gx.to_host()[lo:lo + n], xc, equal_nan=True) and np.array_equal(gs.to_host()[lo:lo + n], stc)
    assert np.isnan(gx.to_host()[:lo]).all()


def _peer_gather_rank(rank, world, port, n_total, tmp):
    """One process per GPU: solve the local shard with the peers' shard addresses as result mirrors (sharding.PeerGather),
    then compare the gathered arrays with a single-GPU solve of the whole batch done on this rank."""
    import os
    import sys
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "multiple-quadrotor-slam_b200"), os.path.join(root, "harness")):
        sys.path.insert(0, p)
    import numpy as np
    import torch.distributed as dist
    import sharding
    import synthetic_rig as rig
    import triangl_cuda as tc
    tc.check(tc.lib().trgl_set_device(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)          # handle exchange + barrier only (plumbing)
    result = {}
    try:
        u1, P1, u2, P2, _ = rig.make_correspondences(n_total, "rotating", sigma=0.8)
        d1, d2 = tc.to_device(u1), tc.to_device(u2)
        lo, hi = sharding.shard_range(n_total, rank, world)
        s1 = d1.view(2 * lo, (hi - lo, 2)); s2 = d2.view(2 * lo, (hi - lo, 2))
        for name in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
            sdt = np.int32 if name == "iterative_LS" else np.uint8
            fn = {"linear_eigen": tc.linear_eigen, "linear_LS": tc.linear_ls, "iterative_LS": tc.iterative_ls,
                  "polynomial": lambda *a, **k: tc.polynomial(*a, check_all_nan=False, **k)[:2]}[name]
            pg = sharding.PeerGather(n_total, np.float64, sdt)
            tc.check(tc.lib().trgl_memset_d(pg.x_all.ptr, 0xff, pg.x_all.nbytes, None))
            tc.check(tc.lib().trgl_memset_d(pg.status_all.ptr, 0x7f, pg.status_all.nbytes, None))
            tc.synchronize(); dist.barrier()                               # nobody stores into a buffer still being cleared
            xs, ss = pg.shard_outputs()
            pg.arm()
            fn(s1, P1, s2, P2, x=xs, status=ss)
            pg.finish()
            x_full, st_full = fn(d1, P1, d2, P2)                           # single-GPU result of the whole batch
            tc.synchronize()
            result[name] = bool(np.array_equal(pg.x_all.to_host(), x_full.to_host(), equal_nan=True) and
                                np.array_equal(pg.status_all.to_host(), st_full.to_host()))
            dist.barrier()
            pg.close()
    except Exception as exc:        # noqa: BLE001
        result["error"] = repr(exc)
    import json
    with open(os.path.join(tmp, "peer_%d.json" % rank), "w") as f:
        json.dump(result, f)
    dist.destroy_process_group()


def test_peer_gather_across_real_gpus(tri, tmp_path):
    """sharding.PeerGather across DIFFERENT GPUs (>= 2 devices): every rank's solver kernels store their shard into every
    peer's gathered arrays over NVLink; afterwards x_all / status_all on every rank equal the single-GPU result bit for bit."""
    import triangl_cuda as tc
    import torch.multiprocessing as mp
    world = min(tc.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    port = 29700 + os.getpid() % 1000
    mp.spawn(_peer_gather_rank, args=(world, port, 300_007, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        with open(os.path.join(str(tmp_path), "peer_%d.json" % r)) as f:
            res = json.load(f)
        assert "error" not in res, res
        assert res == {"linear_eigen": True, "linear_LS": True, "iterative_LS": True, "polynomial": True}, (r, res)


# ---- evaluation fused into the solver kernels ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"])
def test_fused_evaluation_equals_separate_pass(tri, name, dtype):
    """trgl_set_fused_eval: reprojection errors and good mask computed in the solver's epilogue are bit-identical to the
    stand-alone pass over the stored results (trgl_pair_reproj); the sums agree to rounding, the counts exactly."""
    import triangl_cuda as tc
    n = 150001
    u1, P1, u2, P2, _ = rig.make_correspondences(n, "rotating", sigma=2.0, dtype=dtype)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    fn = {"linear_eigen": tc.linear_eigen, "linear_LS": tc.linear_ls, "iterative_LS": tc.iterative_ls,
          "polynomial": lambda *a, **k: tc.polynomial(*a, check_all_nan=False, **k)[:2]}[name]
    thr = (2.0 / 480) ** 2
    fe = tc.FusedEval(n, dtype, 0, thr, want_errors=True, want_good=True)
    x, st = fn(d1, P1, d2, P2, out_dtype=dtype, evaluate=fe)
    x_plain, st_plain = fn(d1, P1, d2, P2, out_dtype=dtype)                   # the request applied to ONE call
    e1, e2, good, sums = tc.pair_reproj(x, d1, P1, d2, P2, st, 0, thr)
    tc.synchronize()
    assert np.array_equal(x.to_host(), x_plain.to_host(), equal_nan=True) and np.array_equal(st.to_host(), st_plain.to_host())
    assert np.array_equal(fe.good.to_host(), good.to_host())
    assert np.array_equal(fe.err1.to_host(), e1.to_host(), equal_nan=True) and np.array_equal(fe.err2.to_host(), e2.to_host(), equal_nan=True)
    fs = fe.sums.to_host()
    assert fs[2] == sums[2] and fs[3] == sums[3] and 0 < sums[2] < n
    assert fs[0] == pytest.approx(sums[0], rel=1e-11) and fs[1] == pytest.approx(sums[1], rel=1e-11)


def test_fused_evaluation_with_pixel_inputs_and_small_batches(tri):
    """Pixel-input solve + evaluation in one kernel (the whole SLAM keyframe triangulation step, slam2.py:551-563), and
    batch sizes around the tile / queue boundaries."""
    import triangl_cuda as tc
    for n in (1, 31, 257, 1000, 20011):
        px1, P1, px2, P2, K, dist = _pixel_batch(n, dtype=np.float32)
        intr = tc.Intrinsics(K, dist)
        d1, d2 = tc.to_device(px1), tc.to_device(px2)
        fe = tc.FusedEval(n, np.float32, 0, 1e-4, want_errors=True)
        x, st = tc.iterative_ls(d1, P1, d2, P2, out_dtype=np.float32, pixel=intr, evaluate=fe)
        n1 = tc.undistort_points(d1, K, dist); n2 = tc.undistort_points(d2, K, dist)
        e1, e2, good, sums = tc.pair_reproj(x, n1, P1, n2, P2, st, 0, 1e-4)
        tc.synchronize()
        assert np.array_equal(fe.good.to_host(), good.to_host()) and np.array_equal(fe.err1.to_host(), e1.to_host(), equal_nan=True)
        fs = fe.sums.to_host()
        assert fs[2] == sums[2] and fs[3] == sums[3] and fs[0] == pytest.approx(sums[0], rel=1e-11, abs=1e-30)
    with pytest.raises((tc.TrianglCudaError, ValueError)):        # host arrays cannot carry a fused evaluation
        tc.linear_ls(px1, P1, px2, P2, evaluate=fe)


def test_per_call_requests_do_not_outlive_a_failed_call(tri):
    """Result mirrors and the fused evaluation are consumed by the next solver call even when that call fails."""
    import triangl_cuda as tc
    n = 5000
    u1, P1, u2, P2, _ = rig.make_correspondences(n, "rotating", sigma=0.8)
    d1, d2 = tc.to_device(u1), tc.to_device(u2)
    gx, gs = tc.DeviceArray((n, 3), np.float64), tc.DeviceArray((n,), np.uint8)
    tc.check(tc.lib().trgl_memset_d(gx.ptr, 0xff, gx.nbytes, None))
    fe = tc.FusedEval(n, np.float64)
    tc.check(tc.lib().trgl_memset_d(fe.sums.ptr, 0xff, 32, None))
    tc.set_result_mirrors([(gx.ptr, gs.ptr)])
    with pytest.raises(tc.TrianglCudaError):
        tc.linear_eigen(d1, P1, d2, P2, rows=5, evaluate=fe)            # bad argument: rows must be 4 or 6
    x, st = tc.linear_eigen(d1, P1, d2, P2)                              # must run without mirrors / evaluation
    tc.synchronize()
    assert np.isnan(gx.to_host()).all() and np.isnan(fe.sums.to_host()).all()
    assert np.isfinite(x.to_host()).all()


# ---- multi-view linear LS -----------------------------------------------------------------------
def multiview_well_posed(us, Ps, valid, tol=1e-9):
    m, n = us.shape[0], us.shape[1]
    A = np.zeros((n, 2 * m, 3))
    for v in range(m):
        P = np.asarray(Ps[v])
        A[:, 2 * v] = (us[v][:, 0:1] * P[2] - P[0])[:, 0:3] * valid[v][:, None]
        A[:, 2 * v + 1] = (us[v][:, 1:2] * P[2] - P[1])[:, 0:3] * valid[v][:, None]
    s = np.linalg.svd(A, compute_uv=False)
    with np.errstate(all='ignore'):
        return (s[:, 0] / s[:, 2]) * 2.2e-16 * 50 < tol


@pytest.mark.gpu
@pytest.mark.parametrize("num_cams", [2, 3, 8, 16])
@pytest.mark.parametrize("p_visible", [1.0, 0.6])
def test_multiview_ls_vs_oracle(tri, num_cams, p_visible):
    us, Ps, X, valid = rig.make_multiview(20011, num_cams, 0.8, p_visible=p_visible)
    mask = None if p_visible >= 1.0 else valid
    x, st = tri.multiview_LS_triangulation(us, Ps, mask)
    xo, so = orc.multiview_LS_triangulation(us, Ps, mask)
    assert x.shape == (20011, 3) and x.dtype == np.float64 and st.dtype == np.bool_
    assert np.array_equal(st, so)
    ok = multiview_well_posed(us, Ps, valid) & so
    assert ok.mean() > 0.95 * so.mean() > 0.3
    assert rel_err(x, xo)[ok].max() < TOL64
    # points seen by 0 or 1 view: minimum-norm solutions of rank <= 2 systems, like the oracle
    few = valid.sum(axis=0) < 2
    if few.any():
        assert np.isfinite(x[few]).all()
        assert np.abs(x[few] - xo[few]).max() < 1e-8 * max(1.0, np.abs(xo[few]).max())
    if num_cams >= 8 and p_visible >= 1.0:           # more views -> closer to the ground truth than any pair
        x2, _ = tri.linear_LS_triangulation(us[0], Ps[0], us[-1], Ps[-1])
        assert np.linalg.norm(x - X, axis=1).mean() < np.linalg.norm(x2 - X, axis=1).mean()


@pytest.mark.gpu
def test_multiview_two_views_is_linear_ls(tri):
    import triangl_cuda as tc
    for dtype in (np.float64, np.float32):
        u1, P1, u2, P2, X = rig.make_correspondences(30011, "rotating", 0.8, dtype=dtype)
        x2, s2 = tri.linear_LS_triangulation(u1, P1, u2, P2)
        xm, sm = tri.multiview_LS_triangulation(np.stack([u1, u2]), [P1, P2])
        assert np.array_equal(sm, s2)
        assert rel_err(xm, x2).max() < (1e-12 if dtype == np.float64 else 1e-6)
    # device-resident call, float32 output, 4x4 matrices, ragged size
    us, Ps, X, valid = rig.make_multiview(1000 + 37, 5, 0.8, p_visible=0.8)
    P4 = [np.vstack([P, [0, 0, 0, 1]]) for P in Ps]
    xd, sd = tc.multiview_ls(tc.to_device(us), P4, tc.to_device(valid.astype(np.uint8)), out_dtype=np.float32)
    xh, sh = tc.multiview_ls(us, Ps, valid, out_dtype=np.float32)
    assert np.array_equal(xd.to_host(), xh) and np.array_equal(sd.to_host().astype(bool), sh.astype(bool))
    for n in (0, 1, 33):
        x, st = tri.multiview_LS_triangulation(us[:, :n], Ps, valid[:, :n])
        assert x.shape == (n, 3) and st.shape == (n,)


@pytest.mark.gpu
def test_multiview_degenerate_and_nan(tri):
    us, Ps, X, valid = rig.make_multiview(600, 4, 0.5)
    # all four "cameras" identical: every system has rank 2 -> minimum-norm solution, as cvSolve(DECOMP_SVD)
    same = np.stack([us[0]] * 4)
    x, st = tri.multiview_LS_triangulation(same, [Ps[0]] * 4)
    xo, so = orc.multiview_LS_triangulation(same, [Ps[0]] * 4)
    assert np.isfinite(x).all() and st.all()
    assert rel_err(x, xo).max() < 1e-8
    # NaN / Inf observations poison their point only -- unless the view is masked out
    bad = us.copy(); bad[1, 5, 0] = np.nan; bad[2, 9, 1] = np.inf
    x, st = tri.multiview_LS_triangulation(bad, Ps)
    assert not np.isfinite(x[5]).all() and not np.isfinite(x[9]).all()
    keep = np.ones(600, bool); keep[[5, 9]] = False
    xo, _ = orc.multiview_LS_triangulation(us, Ps)
    assert rel_err(x, xo)[keep].max() < TOL64
    valid2 = np.ones((4, 600), bool); valid2[1, 5] = False; valid2[2, 9] = False
    x, st = tri.multiview_LS_triangulation(bad, Ps, valid2)
    xo, so = orc.multiview_LS_triangulation(us, Ps, valid2)
    assert np.isfinite(x).all() and rel_err(x, xo).max() < TOL64 and np.array_equal(st, so)


@pytest.mark.gpu
def test_fp64_fma_rate_probe():
    """bench.py's FP64 roofline denominator: the two-register-source rate is near the nominal 64 lanes per clock and SM, and
    a DFMA with three distinct register sources is slower by about the 2 : 3 the register-file banks allow."""
    import triangl_cuda as tc
    tc.require_device()
    r2 = tc.fp64_fma_rate(2, 8, 4)
    r3 = tc.fp64_fma_rate(3, 8, 4)
    nominal = 148 * 2 * 1.965e9                    # warp instructions per second at the boost clock
    assert 0.6 * nominal < r2 < 1.1 * nominal
    assert 0.55 < r3 / r2 < 0.9
    one_chain = tc.fp64_fma_rate(3, 1, 2)          # 16 warps per SM, one dependent chain each: latency-bound
    assert one_chain < r3
    with pytest.raises(RuntimeError):
        tc.fp64_fma_rate(4, 8, 4)
