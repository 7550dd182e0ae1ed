"""
CPU-side tests (-m "not gpu"): the C ABI library loads and exports every symbol include/triangl_cuda.h declares,
the product path fails loudly without a GPU (no CPU fallback), host-side logic (input coercion, sharding, rig),
and the oracles agree with each other and with cv2 (the executable statement of the third-party OpenCV calls).
"""
import ctypes
import os
import re

import numpy as np
import pytest

import synthetic_rig as rig
from oracle import triangulation_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import triangl_cuda as tc
    if not os.path.isfile(tc.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return tc


def test_header_symbols_exported():
    tc = _lib()
    header = open(os.path.join(ROOT, "include", "triangl_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(trgl_[A-Za-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    L = ctypes.CDLL(tc.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "libtriangl_cuda.so does not export %s" % name
    assert sorted(tc.EXPORTS) == declared
    m = re.search(r"#define TRGL_VERSION (\d+)", header)
    assert L.trgl_version() == int(m.group(1))


def test_kernels_are_sm100a_native():
    """The shipped library contains sm_100a SASS for every solver, including the bulk-async (TMA engine) pipeline."""
    import subprocess
    tc = _lib()
    out = subprocess.run(["cuobjdump", "-lelf", tc.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", tc.LIB_PATH], capture_output=True, text=True).stdout
    for kern in ("k_linear_ls", "k_linear_ls_tma", "k_iterative_ls", "k_linear_eigen", "k_polynomial",
                 "k_reproj_error", "k_pair_reproj"):
        assert kern in sass, kern
    assert "UBLKCP" in sass and "SYNCS" in sass        # cp.async.bulk + mbarrier transaction


def test_no_cpu_fallback_without_gpu():
    tc = _lib()
    if tc.device_count() > 0:
        pytest.skip("a GPU is present")
    import triangulation
    import calibration_tools
    u = np.zeros((4, 2))
    for fn in (triangulation.linear_eigen_triangulation, triangulation.linear_LS_triangulation,
               triangulation.iterative_LS_triangulation, triangulation.polynomial_triangulation):
        with pytest.raises(tc.TrianglCudaError):
            fn(u, np.eye(4), u, np.eye(4))
    with pytest.raises(tc.TrianglCudaError):
        calibration_tools.reprojection_error(np.zeros((4, 3)), u, np.eye(3), np.zeros(5), np.zeros(3), np.zeros(3))
    with pytest.raises(tc.TrianglCudaError):
        tc.require_device()


def test_bad_arguments_rejected_before_any_device_work():
    tc = _lib()
    u = np.zeros((4, 2))
    with pytest.raises(ValueError):
        tc.linear_ls(u, np.eye(3), u, np.eye(4))              # not 3x4 / 4x4
    with pytest.raises(ValueError):
        tc.linear_ls(u, np.eye(4), np.zeros((5, 2)), np.eye(4))
    us = np.zeros((3, 4, 2))
    with pytest.raises(ValueError):
        tc.multiview_ls(us[:, :, 0], [np.eye(4)] * 3)                 # not (m, n, 2)
    with pytest.raises(ValueError):
        tc.multiview_ls(us, [np.eye(4)] * 2)                          # one camera matrix per view
    with pytest.raises(ValueError):
        tc.multiview_ls(us, [np.eye(4)] * 3, valid=np.ones((3, 5)))   # mask must be (m, n)
    L = tc.lib()
    P = (ctypes.c_double * 12)()
    assert L.trgl_multiview_ls(None, None, P, 17, None, None, 0, 2, 0, 0, None) == -1   # more than 16 views
    assert L.trgl_multiview_ls(None, None, P, 2, None, None, 4, 2, 0, 0, None) == -1    # NULL arrays, n > 0
    assert L.trgl_set_two_ray(1) == 1 and L.trgl_set_deferred_capacity(1 << 26) == 1 << 26     # knobs return the old value
    assert L.trgl_linear_ls(None, None, P, P, None, None, 4, 0, 0, None) == -1      # NULL arrays, n > 0
    assert L.trgl_linear_ls(None, None, P, P, None, None, 4, 99, 0, None) == -1     # unknown mode
    assert b"mode" in L.trgl_last_error_string()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multiple-quadrotor-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, "%s references the oracle" % f


def test_module_api_matches_reference_signatures():
    import inspect
    import triangulation as tri
    sig = {name: list(inspect.signature(getattr(tri, name)).parameters.items()) for name in
           ("linear_eigen_triangulation", "linear_LS_triangulation", "iterative_LS_triangulation",
            "polynomial_triangulation", "set_triangl_output_dtype")}
    assert [k for k, _ in sig["linear_eigen_triangulation"]] == ["u1", "P1", "u2", "P2", "max_coordinate_value"]
    assert sig["linear_eigen_triangulation"][4][1].default == 1.e16
    assert [k for k, _ in sig["linear_LS_triangulation"]] == ["u1", "P1", "u2", "P2"]
    assert [k for k, _ in sig["iterative_LS_triangulation"]] == ["u1", "P1", "u2", "P2", "tolerance"]
    assert sig["iterative_LS_triangulation"][4][1].default == 3.e-5
    assert [k for k, _ in sig["polynomial_triangulation"]] == ["u1", "P1", "u2", "P2"]
    assert tri.output_dtype is float


# ---- oracle cross-checks -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("rig_name", list(rig.RIGS))
def test_c_oracle_matches_numpy_oracle(rig_name):
    from oracle import oracle_c as oc
    u1, P1, u2, P2, X = rig.make_correspondences(4000, rig_name, 0.8)
    for name in orc.SOLVERS:
        x, st = oc.SOLVERS[name](u1, P1, u2, P2)
        xo, so = orc.SOLVERS[name](u1, P1, u2, P2)
        assert np.array_equal(st, so), name
        with np.errstate(all='ignore'):
            rel = np.abs(x - xo).max(axis=1) / np.abs(xo).max(axis=1)
        assert np.nanmax(rel) < (1e-8 if rig_name == "forward" else 1e-11), (name, np.nanmax(rel))


def test_oracle_matches_cv2():
    """cv2 4.x is the only executable statement of triangulatePoints / correctMatches / projectPoints / solve."""
    cv2 = pytest.importorskip("cv2")
    for rig_name in rig.RIGS:
        u1, P1, u2, P2, X = rig.make_correspondences(3000, rig_name, 2.0)
        Xh = cv2.triangulatePoints(P1, P2, u1.T.copy(), u2.T.copy())
        xo = orc.eigen_homogeneous(u1, P1, u2, P2, 4)
        a = (Xh[0:3] / Xh[3:4]).T; b = xo[:, 0:3] / xo[:, 3:4]
        well = np.abs(b).max(axis=1) < 1e3
        assert np.max(np.abs(a - b)[well] / np.abs(b)[well].max(axis=1, keepdims=True)) < 1e-8
        F = orc.fundamental_from_P(P1, P2)
        c1, c2 = cv2.correctMatches(F, u1.reshape(1, -1, 2), u2.reshape(1, -1, 2))
        o1, o2 = orc.correct_matches(F, u1, u2)
        assert np.nanmax(np.abs(c1[0] - o1)) < 1e-10 and np.nanmax(np.abs(c2[0] - o2)) < 1e-10
    A, b = orc.build_Ab(u1[:50], P1, u2[:50], P2)
    for k in range(50):
        x = np.zeros((3, 1))
        cv2.solve(A[k], b[k].reshape(4, 1), x, cv2.DECOMP_SVD)
        assert np.allclose(x[:, 0], orc.lstsq_minnorm(A[k:k + 1], b[k:k + 1])[0], rtol=1e-9, atol=1e-12)
    K = np.array([[480., 0, 320], [0, 480., 240], [0, 0, 1]])
    dist = np.array([0.1, -0.05, 0.001, -0.002, 0.01]); rvec = np.array([0.02, 0.29, -0.01]); tvec = np.array([-1., 0.1, 40.])
    p, _ = cv2.projectPoints(X, rvec, tvec, K, dist)
    assert np.max(np.abs(p.reshape(-1, 2) - orc.project_points(X, rvec, tvec, K, dist))) < 1e-9


def test_rank_deficient_oracle_is_min_norm():
    u1, P1, u2, P2, X = rig.make_correspondences(20, "translating", 0.0)
    A, b = orc.build_Ab(u1, P1, u1, P1)
    x = orc.lstsq_minnorm(A, b)
    for k in range(20):
        assert np.allclose(x[k], np.linalg.lstsq(A[k], b[k], rcond=None)[0], atol=1e-10)


def test_rig_is_deterministic_and_matches_reference_parameters():
    a = rig.make_correspondences(1000, "rotating", 0.8)
    b = rig.make_correspondences(1000, "rotating", 0.8)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert len(rig.finite_3D_points(4)) == 257                       # the reference's sphere cloud
    cam1, cam2 = rig.make_cameras("translating")
    assert np.allclose(cam1.P, [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 40]])
    assert np.allclose(cam2.P, [[1, 0, 0, -5], [0, 1, 0, 0], [0, 0, 1, 40]])


# ---- sharding ----------------------------------------------------------------------------------------------------
def test_shard_ranges_cover_everything():
    import sharding
    for n in (0, 1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    segs = sharding.pair_segments(8, 800_000_001)
    assert len(segs) == 28 and sum(c for *_, c in segs) == 800_000_001
    pieces = [p for k in range(8) for p in sharding.intersect_segments(segs, *sharding.shard_range(800_000_001, k, 8))]
    assert sum(c for *_, c in pieces) == 800_000_001


def _gloo_worker(rank, world, port, tmp):
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
    from oracle import triangulation_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u1, P1, u2, P2, X = rig.make_correspondences(1001, "rotating", 0.8)
    if rank != 0:                      # only rank 0 knows the cameras: they must arrive by broadcast
        P1 = np.zeros_like(P1); P2 = np.zeros_like(P2)
    x, st = sharding.triangulate_sharded(orc.iterative_LS_triangulation, u1, P1, u2, P2)
    cams = rig.circle_cameras(4)
    segs = sharding.pair_segments(4, 999)
    rng = np.random.RandomState(5)
    u_by = {(i, j): (rng.normal(0, 0.05, (c, 2)), rng.normal(0, 0.05, (c, 2))) for (i, j, o, c) in segs}
    xp, sp = sharding.triangulate_pairs_sharded(orc.linear_LS_triangulation, u_by, cams if rank == 0 else
                                                [np.zeros((3, 4))] * 4, segs)
    us, Ps, Xm, valid = rig.make_multiview(1003, 5, 0.8, p_visible=0.7)
    xm, sm = sharding.triangulate_multiview_sharded(orc.multiview_LS_triangulation, us,
                                                    Ps if rank == 0 else [np.zeros((3, 4))] * 5, valid)
    np.savez(os.path.join(tmp, "r%d.npz" % rank), x=x, st=st, xp=xp, sp=sp, xm=xm, sm=sm)
    dist.destroy_process_group()


def test_sharded_path_world_size_2_gloo(tmp_path):
    """N > 1 host logic on CPU: shard, broadcast cameras from rank 0, solve (oracle injected as the solver), gather."""
    import torch.multiprocessing as mp
    import sharding
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    u1, P1, u2, P2, X = rig.make_correspondences(1001, "rotating", 0.8)
    xo, so = orc.iterative_LS_triangulation(u1, P1, u2, P2)
    cams = rig.circle_cameras(4)
    segs = sharding.pair_segments(4, 999)
    rng = np.random.RandomState(5)
    u_by = {(i, j): (rng.normal(0, 0.05, (c, 2)), rng.normal(0, 0.05, (c, 2))) for (i, j, o, c) in segs}
    xp = np.concatenate([orc.linear_LS_triangulation(u_by[(i, j)][0], cams[i], u_by[(i, j)][1], cams[j])[0]
                         for (i, j, o, c) in segs])
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(d["st"], so) and np.allclose(d["x"], xo, rtol=1e-12, atol=1e-12)
        assert d["xp"].shape == (999, 3) and np.allclose(d["xp"], xp, rtol=1e-12, atol=1e-12) and d["sp"].all()
    # m-view scene: points sharded, all cameras broadcast
    us, Ps, Xm, valid = rig.make_multiview(1003, 5, 0.8, p_visible=0.7)
    xmo, smo = orc.multiview_LS_triangulation(us, Ps, valid)
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(d["sm"], smo) and np.allclose(d["xm"], xmo, rtol=1e-12, atol=1e-12)


# ---- input normalisation: oracle pinned to cv2 ------------------------------------------------
def test_oracle_undistort_bit_identical_to_cv2_fixture():
    """tests/golden/cv2_undistort.npz holds cv2 4.13 `undistortPoints` output (oracle/make_golden.py): the NumPy
    restatement must reproduce every bit, float64 and float32, for all distortion models incl. the icdist < 0 branch."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_undistort.npz"))
    px, K = g["px"], g["K"]
    for k, d in enumerate(g["dists"]):
        assert np.array_equal(orc.undistort_points(px, K, d), g["n64_%d" % k])
        out32 = orc.undistort_points(px.astype(np.float32), K, d)
        assert out32.dtype == np.float32 and np.array_equal(out32, g["n32_%d" % k])
    assert np.array_equal(orc.undistort_points(px, K, None), g["n64_none"])


def test_oracle_undistort_matches_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    K = np.array([[525., 0, 319.5], [0, 525., 239.5], [0, 0, 1]])
    px = rng.uniform(0, 640, (5000, 2))
    for d in ([0.2, -0.3, 1e-3, -2e-3, 0.1], [0.1, 0.01, 0, 0]):
        ref = cv2.undistortPoints(px.reshape(-1, 1, 2), K, np.array(d)).reshape(-1, 2)
        assert np.array_equal(orc.undistort_points(px, K, d), ref)


def test_P_from_rvec_and_tvec_matches_cv2_rodrigues():
    cv2 = pytest.importorskip("cv2")
    import triangulation
    rng = np.random.RandomState(11)
    for rvec in list(rng.normal(0, 1.5, (20, 3))) + [np.zeros(3), np.array([1e-20, 0, 0])]:
        tvec = rng.normal(0, 3, (3, 1))
        P = triangulation.P_from_rvec_and_tvec(rvec, tvec)
        assert P.shape == (4, 4) and np.array_equal(P[3], [0, 0, 0, 1])
        assert np.allclose(P[0:3, 0:3], cv2.Rodrigues(rvec)[0], atol=1e-14) and np.array_equal(P[0:3, 3:4], tvec)


# ---- bench.py contract (CPU-checkable part) ----------------------------------------------------------------------------
def _bench_lines(cmd):
    import subprocess
    import sys
    out = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_bench_reference_arm_prints_one_json_line():
    import json
    lines = _bench_lines(["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "20000"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "triangulated_points_per_sec" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_under_torchrun_only_rank0_prints():
    import json
    lines = _bench_lines(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "20000"])
    js = [ln for ln in lines if ln.startswith("{")]
    assert len(js) == 1 and json.loads(js[0])["n_gpus"] == 2
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every core of the box
    assert json.loads(js[0])["cpu_baseline"]["cores"] == max(1, len(os.sched_getaffinity(0)))
