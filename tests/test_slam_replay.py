"""
SLAM keyframe replay: loader of the reference's BA_info / trajectory logs, and the recorded SVO run
(tests/golden/slam_replay_svo.npz, written by oracle/make_golden.py from the dataset shipped with the reference together
with the REFERENCE's own results on every keyframe batch).
"""
import os

import numpy as np
import pytest

import slam_replay
from oracle import triangulation_oracle as orc


def _fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "slam_replay_svo.npz"))
    off = g["offsets"]
    poses = []
    kfs = []
    for k in range(len(g["steps"])):
        poses += [g["P0"][k], g["P1"][k]]
        kfs.append({"step": int(g["steps"][k]), "frame0": 2 * k, "frame1": 2 * k + 1, "idx3d": np.arange(off[k], off[k + 1]),
                    "px0": g["px0"][off[k]:off[k + 1]], "px1": g["px1"][off[k]:off[k + 1]]})
    return g, {"K": g["K"], "dist": g["dist"], "poses": poses, "keyframes": kfs}


def _write_dataset(tmp, K, dist, poses_tum, frames_px, steps):
    """Restates the writer of slam2.py:795-858 for a toy run: frames_px[frame] = (n,2); steps[s] = (added ids, assocs)."""
    with open(os.path.join(tmp, "BA_info.calibrations.cam0.txt"), "w") as f:
        f.write("# Format: fx fy shear u0 v0 k1 k2 p1 p2\n")
        f.write(" ".join("%.16e" % v for v in (K[0, 0], K[1, 1], K[0, 1], K[0, 2], K[1, 2]) + tuple(dist[:4])) + "\n")
    with open(os.path.join(tmp, "BA_info.measurements.points2D.cam0-slam2.txt"), "w") as f:
        f.write("# Format: x y\n# Newline means next feature; Empty line means next frame, first feature\n")
        f.write("\n\n".join("\n".join("%.16e %.16e" % tuple(p) for p in fr) for fr in frames_px) + "\n")
    with open(os.path.join(tmp, "BA_info.measurements.point2D3DAssocs.cam0-slam2.txt"), "w") as f:
        f.write("# Format: frameIdx point2DIdx point3DIdx\n# Newline means next feature; Empty line means next step, first feature\n")
        f.write("\n\n".join("\n".join("%d %d %d" % a for a in st[1]) for st in steps) + "\n")
    with open(os.path.join(tmp, "BA_info.measurements.point3DAddedIdxs-slam2.txt"), "w") as f:
        f.write("# Format: point3DIdx\n# Newline means next point; Empty line means next step\n")
        f.write("\n\n".join("\n".join(str(i) for i in st[0]) for st in steps) + "\n")
    with open(os.path.join(tmp, "traj_out.cam0-slam2.txt"), "w") as f:
        f.write("# Format: timestamp tx ty tz qx qy qz qw\n")
        for k, (q, l) in enumerate(poses_tum):
            f.write("%.2f %s %s\n" % (0.02 * (k + 1), " ".join(repr(float(v)) for v in l), " ".join(repr(float(v)) for v in q)))


def test_loader_rebuilds_keyframe_batches(tmp_path):
    K = np.array([[300., 0, 320.], [0, 310., 240.], [0, 0, 1]]); dist = np.array([0.1, -0.05, 1e-3, 2e-3, 0.])
    poses = [((0., 0., 0., 1.), (0., 0., 0.)), ((0., np.sin(0.1), 0., np.cos(0.1)), (1., 0., 0.)), ((0., 0., 0., 1.), (2., 0.5, 0.))]
    rng = np.random.RandomState(0)
    frames = [rng.uniform(0, 640, (4, 2)), rng.uniform(0, 640, (5, 2)), rng.uniform(0, 640, (3, 2))]
    steps = [([0, 1], [(0, 0, 0), (0, 1, 1)]),                                   # initial map: one frame only -> no batch
             ([2, 3], [(0, 2, 2), (1, 0, 2), (0, 3, 3), (1, 1, 3), (1, 2, 0)]),  # keyframe (frames 0,1) + a re-observation
             ([4], [(1, 4, 4), (2, 2, 4), (2, 0, 3)])]                           # keyframe (frames 1,2)
    _write_dataset(str(tmp_path), K, dist, poses, frames, steps)
    ds = slam_replay.load_dataset(str(tmp_path))
    assert np.array_equal(ds["K"], K) and np.array_equal(ds["dist"], dist)
    assert len(ds["poses"]) == 3 and len(ds["keyframes"]) == 2
    k0, k1 = ds["keyframes"]
    assert (k0["step"], k0["frame0"], k0["frame1"]) == (1, 0, 1) and list(k0["idx3d"]) == [2, 3]
    assert np.array_equal(k0["px0"], frames[0][[2, 3]]) and np.array_equal(k0["px1"], frames[1][[0, 1]])
    assert (k1["frame0"], k1["frame1"]) == (1, 2) and np.array_equal(k1["px0"], frames[1][[4]]) and np.array_equal(k1["px1"], frames[2][[2]])
    # pose: world-to-camera = inverse of the TUM pose
    P1 = ds["poses"][1]
    assert np.allclose(P1[0:3, 0:3].dot(P1[0:3, 0:3].T), np.eye(3), atol=1e-15)
    assert np.allclose(P1.dot([1., 0., 0., 1.]), [0, 0, 0, 1], atol=1e-15)           # the camera centre maps to the origin
    assert np.allclose(P1[0:3, 0:3].T, orc.rodrigues([0, 0.2, 0]), atol=1e-15)


def test_oracle_reproduces_reference_results_on_recorded_keyframes(golden_dir):
    """The oracle chain (undistort restatement -> iterative_LS, Python semantics) against the reference's own output on
    all 172 recorded keyframes: status identical, x within float32 storage rounding."""
    g, ds = _fixture(golden_dir)
    assert len(ds["keyframes"]) == 172 and g["offsets"][-1] == 946
    sizes = np.diff(g["offsets"])
    assert sizes.min() >= 1 and sizes.max() <= 13                      # real keyframe batches are tiny
    for k, kf in enumerate(ds["keyframes"]):
        lo, hi = g["offsets"][k], g["offsets"][k + 1]
        n0 = orc.undistort_points(kf["px0"], g["K"], g["dist"]); n1 = orc.undistort_points(kf["px1"], g["K"], g["dist"])
        x, st = orc.iterative_LS_triangulation(n0, g["P0"][k], n1, g["P1"][k], semantics='py')
        assert np.array_equal(st, g["status_first"][lo:hi])
        ref = g["x_first"][lo:hi].astype(np.float64)
        assert np.max(np.abs(x - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) < 2e-7


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_gpu_replay_matches_reference_results(golden_dir, fused):
    import triangl_cuda
    triangl_cuda.require_device()
    import triangulation as tri
    g, ds = _fixture(golden_dir)
    tri.set_triangl_semantics('py')
    try:
        res = slam_replay.replay(ds, fused=fused)
    finally:
        tri.set_triangl_semantics('c')
    assert len(res) == 172
    for k, r in enumerate(res):
        lo, hi = g["offsets"][k], g["offsets"][k + 1]
        assert r["x"].dtype == np.float32
        assert np.array_equal(r["status"], g["status_first"][lo:hi])
        ref = g["x_first"][lo:hi].astype(np.float64)
        assert np.max(np.abs(r["x"] - ref) / np.max(np.abs(ref), axis=1, keepdims=True)) < 2e-7
        lo2, hi2 = g["offsets_second"][k], g["offsets_second"][k + 1]
        assert np.array_equal(r["inliers"], np.where(g["status_first"][lo:hi] == 1)[0])
        ref2 = g["x_second"][lo2:hi2].astype(np.float64)
        assert r["x_final"].shape == ref2.shape
        assert np.max(np.abs(r["x_final"] - ref2) / np.max(np.abs(ref2), axis=1, keepdims=True)) < 2e-7
