"""
Generate the committed parity fixtures under tests/golden/ (run in the build container only):

  python oracle/make_golden.py

1. ref_py_<rig>.npz : per-point inputs and the outputs of the reference's own pure-Python solvers
   (Work/python_libs/triangulation.py:1-233 exec'd unmodified, cv2 4.13) on seeded synthetic batches.
6. vector_stat_cells.npz : the reference's stored vector_stat results (per-point mean / covariance of the 3-D error vectors
   over the trials of a trajectory's last pose) from test_1and2.mat.
2. golden_cells.json : sampled cells of the reference's golden result files
   Work/triangulation_comparison/{test_1and2,test_3}.mat and figures_scene/test_1and2.mat (numbers
   only), with the trajectory tables needed to replay them.
3. cv2_undistort.npz : cv2.undistortPoints outputs (cv2 4.13) pinning the input-normalisation stage.
4. slam_replay_svo.npz : per-keyframe batches of the recorded slam2 run on the reference's SVO dataset + the reference's
   own results on them.
5. formats.npz : what the reference's own dataset_tools.py (lines 1-269 exec'd, with the two Python-2 idioms
   `colors != None` / `colors == None` read as identity tests) loads from one of its shipped .pcd maps and TUM
   trajectories, and writes back for them.
"""
import json
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multiple-quadrotor-slam_b200"))
sys.path.insert(0, os.path.join(ROOT, "harness"))

from oracle.ref_exec import REFERENCE_ROOT, load_reference_triangulation   # noqa: E402
import synthetic_rig as rig                                                  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def per_point_fixtures():
    ref = load_reference_triangulation()
    cases = [("translating", 0.8, False, 400), ("rotating", 0.8, False, 400), ("forward", 0.8, True, 400),
             ("general", 4.0, False, 400), ("translating", 0.0, False, 64)]
    for ci, (name, sigma, disc, n) in enumerate(cases):
        u1, P1, u2, P2, X = rig.make_correspondences(n, name, sigma, disc, seed=rig.RSEED + ci)
        out = {"u1": u1, "P1": P1, "u2": u2, "P2": P2, "X": X, "sigma": sigma}
        for fn in ("linear_eigen", "linear_LS", "iterative_LS", "polynomial"):
            x, st = getattr(ref, fn + "_triangulation")(u1, P1, u2, P2)
            out["x_" + fn] = x
            out["status_" + fn] = st
        np.savez_compressed(os.path.join(GOLDEN, "ref_py_%d_%s.npz" % (ci, name)), **out)
        print("wrote fixture", ci, name, n)
    # 4x4 P, float32 inputs, float32 output dtype: the SLAM calling convention (slam2.py:19,551-555)
    u1, P1, u2, P2, X = rig.make_correspondences(300, "rotating", 0.8, False, seed=rig.RSEED + 77, dtype=np.float32)
    P1f = np.eye(4); P1f[0:3] = P1
    P2f = np.eye(4); P2f[0:3] = P2
    ref.set_triangl_output_dtype(np.float32)
    x, st = ref.iterative_LS_triangulation(u1, P1f, u2, P2f)
    ref.set_triangl_output_dtype(float)
    np.savez_compressed(os.path.join(GOLDEN, "ref_py_slam_f32.npz"), u1=u1, P1=P1f, u2=u2, P2=P2f, X=X,
                        x_iterative_LS=x, status_iterative_LS=st)


def golden_cells():
    base = os.path.join(REFERENCE_ROOT, "Work/triangulation_comparison")
    out = {}
    keys = ["err3D_mean_summary", "err3D_median_summary", "err2D_mean_summary", "err2D_median_summary",
            "false_pos_summary", "false_neg_summary"]

    def traj_table(m):
        t = m["trajectories"]
        tab = []
        t = t.ravel()
        for i in range(len(t)):
            e = t[i]
            tab.append({k: e[k][0, 0].ravel().tolist() for k in ("sideways_values", "towards_values", "angle_values")})
        return tab

    m = sio.loadmat(os.path.join(base, "test_1and2.mat"))
    cells = [(0, 5), (0, 10), (0, 39), (1, 10), (1, 39), (2, 20), (3, 20), (3, 39), (4, 20), (4, 39)]
    out["test_1and2"] = {
        "num_trials": int(m["num_trials"][0, 0]), "rseed": int(m["rseed"][0, 0]),
        "methods": [s.strip() for s in m["triangl_methods"]],
        "trajectories": traj_table(m),
        "cells": [{"traj": t, "pose": p, **{k: m[k][t, p].tolist() for k in keys}} for t, p in cells]}

    m = sio.loadmat(os.path.join(base, "test_3.mat"))
    cells3 = [(0, 0, 20), (3, 2, 39), (4, 1, 10), (2, 0, 5)]
    out["test_3"] = {
        "num_trials": int(m["num_trials"][0, 0]), "rseed": int(m["rseed"][0, 0]),
        "trajectories": traj_table(m),
        "noise_sigma_values": m["noise_sigma_values"].ravel().tolist(),
        "cells": [{"traj": t, "ntype": nt, "nidx": ni, **{k: m[k][t, nt, ni].tolist() for k in keys}}
                  for t, nt, ni in cells3]}

    m = sio.loadmat(os.path.join(base, "figures_scene", "test_1and2.mat"))
    out["scene_test_1and2"] = {
        "num_trials": int(m["num_trials"][0, 0]), "rseed": int(m["rseed"][0, 0]),
        "trajectories": traj_table(m),
        "points_3D": m["points_3D"].tolist(),
        "cells": [{"traj": t, "pose": p, **{k: m[k][t, p].tolist() for k in keys}} for t, p in [(0, 20), (3, 39)]]}

    def clean(o):       # NaN -> None for strict JSON
        if isinstance(o, float):
            return None if o != o else o
        if isinstance(o, list):
            return [clean(v) for v in o]
        if isinstance(o, dict):
            return {k: clean(v) for k, v in o.items()}
        return o
    with open(os.path.join(GOLDEN, "golden_cells.json"), "w") as f:
        json.dump(clean(out), f)
    print("wrote golden_cells.json")


def vector_stat_fixture():
    """6. vector_stat_cells.npz : per-point mean vector / covariance matrix of the 3-D error vectors over the 100 trials of
    the LAST pose of a trajectory (triangulation_comparison.py:483-487, vector_stat :219-240), as stored by the reference
    in test_1and2.mat (p_err3Dv_mean_summary, p_err3Dv_covar_summary), for linear_LS and iterative_LS."""
    m = sio.loadmat(os.path.join(REFERENCE_ROOT, "Work/triangulation_comparison", "test_1and2.mat"))
    methods = [s.strip() for s in m["triangl_methods"]]
    out = {"methods": np.array(methods), "trajs": np.array([0, 3, 4])}
    for t in (0, 3, 4):
        for name in ("linear_LS_triangulation", "iterative_LS_triangulation"):
            ti = methods.index(name)
            out["mean_%d_%s" % (t, name)] = m["p_err3Dv_mean_summary"][t, ti]
            out["covar_%d_%s" % (t, name)] = m["p_err3Dv_covar_summary"][t, ti]
    np.savez_compressed(os.path.join(GOLDEN, "vector_stat_cells.npz"), **out)
    print("wrote vector_stat_cells.npz", methods)


def cv2_undistort_fixture():
    """3. cv2_undistort.npz : cv2.undistortPoints (OpenCV 4.13, the executable third-party statement of the call at
    slam2.py:551-552) on seeded pixel points, float64 and float32, for several distortion models."""
    import cv2
    rng = np.random.RandomState(rig.RSEED + 500)
    K = np.array([[480., 0, 320], [0, 470., 240], [0, 0, 1]])
    px = rng.uniform([-200, -200], [840, 680], (2000, 2))
    dists = np.array([[0.3, -0.1, 0.001, 0.002, 0.05], [-0.6, 0.1, 0, 0, 0], [0.3, 0, 0, 0, 0], [0, 0, 0, 0, 0],
                      [-0.28, 0.07, 2e-4, -1e-4, 0.0]])
    out = {"K": K, "px": px, "dists": dists, "cv2_version": np.array(cv2.__version__)}
    for k, d in enumerate(dists):
        out["n64_%d" % k] = cv2.undistortPoints(px.reshape(1, -1, 2), K, d).reshape(-1, 2)
        out["n32_%d" % k] = cv2.undistortPoints(px.astype(np.float32).reshape(1, -1, 2), K, d).reshape(-1, 2)
    out["n64_none"] = cv2.undistortPoints(px.reshape(1, -1, 2), K, None).reshape(-1, 2)
    np.savez_compressed(os.path.join(GOLDEN, "cv2_undistort.npz"), **out)
    print("wrote cv2_undistort.npz (cv2 %s)" % cv2.__version__)


def slam_replay_fixture():
    """4. slam_replay_svo.npz : the recorded slam2 run on the SVO dataset shipped with the reference
    (Work/SLAM/datasets/SVO/sin2_tex2_h1_v8_d: BA_info.* + traj_out.cam0-slam2.txt), turned into per-keyframe
    triangulation batches by multiple-quadrotor-slam_b200/slam_replay.py, plus what the REFERENCE computes on them:
    cv2.undistortPoints and the reference's own iterative_LS_triangulation (exec'd, float32 points, float32 output,
    the convention of slam2.py:19,551-555), first pass and the re-triangulation of the status == 1 inliers."""
    import cv2
    import slam_replay
    ref = load_reference_triangulation()
    ds = slam_replay.load_dataset(os.path.join(REFERENCE_ROOT, "Work/SLAM/datasets/SVO/sin2_tex2_h1_v8_d"))
    K, dist = ds["K"], ds["dist"]
    off = [0]; px0 = []; px1 = []; P0 = []; P1 = []; steps = []; x1 = []; s1 = []; off2 = [0]; x2 = []; s2 = []
    ref.set_triangl_output_dtype(np.float32)
    for kf in ds["keyframes"]:
        a = kf["px0"].astype(np.float32); b = kf["px1"].astype(np.float32)
        n0 = cv2.undistortPoints(a.reshape(-1, 1, 2), K, dist).reshape(-1, 2)
        n1 = cv2.undistortPoints(b.reshape(-1, 1, 2), K, dist).reshape(-1, 2)
        Pa, Pb = ds["poses"][kf["frame0"]], ds["poses"][kf["frame1"]]
        x, st = ref.iterative_LS_triangulation(n0, Pa, n1, Pb)
        inl = np.where(st == 1)[0]
        xr, sr = ref.iterative_LS_triangulation(n0[inl], Pa, n1[inl], Pb) if len(inl) else (np.zeros((0, 3), np.float32), np.zeros(0, int))
        px0.append(a); px1.append(b); P0.append(Pa); P1.append(Pb); steps.append(kf["step"])
        x1.append(x); s1.append(st); x2.append(xr); s2.append(sr)
        off.append(off[-1] + len(a)); off2.append(off2[-1] + len(inl))
    ref.set_triangl_output_dtype(float)
    np.savez_compressed(os.path.join(GOLDEN, "slam_replay_svo.npz"), K=K, dist=dist, steps=np.array(steps),
                        offsets=np.array(off), px0=np.concatenate(px0), px1=np.concatenate(px1), P0=np.array(P0),
                        P1=np.array(P1), x_first=np.concatenate(x1), status_first=np.concatenate(s1),
                        offsets_second=np.array(off2), x_second=np.concatenate(x2), status_second=np.concatenate(s2))
    print("wrote slam_replay_svo.npz: %d keyframes, %d correspondences (batch sizes %d..%d, median %d)" % (
        len(steps), off[-1], min(np.diff(off)), max(np.diff(off)), int(np.median(np.diff(off)))))


def formats_fixture():
    """5. formats.npz : the reference's PCD / TUM readers and writers (dataset_tools.py:71-272) on files it ships."""
    import tempfile
    import types
    path = os.path.join(REFERENCE_ROOT, "Work/python_libs/dataset_tools.py")
    src = open(path).read().split("\n")[:269]
    src = "\n".join(src).replace("colors != None", "colors is not None").replace("colors == None", "colors is None")
    mod = types.ModuleType("dataset_tools_reference")
    exec(compile(src, path, "exec"), mod.__dict__)
    base = os.path.join(REFERENCE_ROOT, "Work/SLAM/datasets/SVO/sin2_tex2_h1_v8_d")
    pcd_text = open(os.path.join(base, "map_out-slam2.pcd")).read()
    pcd_plain_text = open(os.path.join(base, "init_points.pcd")).read()
    traj_text = open(os.path.join(base, "traj_out.cam0-slam2.txt")).read()
    pts, cols, alpha = mod.load_3D_points_from_pcd_file(os.path.join(base, "map_out-slam2.pcd"), use_alpha=True)
    pts3, cols3, _ = mod.load_3D_points_from_pcd_file(os.path.join(base, "map_out-slam2.pcd"))
    ptsp, colsp, alphap = mod.load_3D_points_from_pcd_file(os.path.join(base, "init_points.pcd"))
    ts, locs, quats = mod.load_cam_trajectory_TUM(os.path.join(base, "traj_out.cam0-slam2.txt"))
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "o")
        mod.save_3D_points_to_pcd_file(f, pts, cols);  saved_bgra = open(f).read()
        mod.save_3D_points_to_pcd_file(f, pts, cols3); saved_bgr = open(f).read()
        mod.save_3D_points_to_pcd_file(f, ptsp);       saved_plain = open(f).read()
        mod.save_cam_trajectory_TUM(f, (ts[:25], locs[:25], quats[:25])); saved_traj = open(f).read()
    assert colsp is None and not alphap
    np.savez_compressed(os.path.join(GOLDEN, "formats.npz"), pcd_text=pcd_text, pcd_plain_text=pcd_plain_text,
                        traj_text=traj_text, points=pts, colors_bgra=cols, found_alpha=alpha, colors_bgr=cols3,
                        points_plain=ptsp, timestps=ts, locations=locs, quaternions=quats, saved_bgra=saved_bgra,
                        saved_bgr=saved_bgr, saved_plain=saved_plain, saved_traj=saved_traj)
    print("wrote formats.npz: %d coloured points, %d plain points, %d poses" % (len(pts), len(ptsp), len(ts)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "formats":
        formats_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "undistort":
        cv2_undistort_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "slam":
        slam_replay_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "vector_stat":
        vector_stat_fixture()
    else:
        vector_stat_fixture()
        per_point_fixtures()
        golden_cells()
        cv2_undistort_fixture()
        slam_replay_fixture()
        formats_fixture()
