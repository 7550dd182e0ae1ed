"""
SLAM keyframe replay: rebuild the per-keyframe triangulation batches of a recorded slam2 run from its log files and
push them through the GPU solvers with the call pattern of the keyframe / map-extension step.

The reference writes, next to every dataset it has processed (writer: Work/SLAM/application/own/slam2.py:743-865,
formats in the file headers):

    BA_info.calibrations.cam0.txt                     "fx fy shear u0 v0 k1 k2 p1 p2"
    BA_info.measurements.points2D.cam0-<name>.txt     "x y"            one block per frame, blank line = next frame
    BA_info.measurements.point2D3DAssocs.cam0-<name>.txt  "frameIdx point2DIdx point3DIdx", blank line = next step
    BA_info.measurements.point3DAddedIdxs-<name>.txt  "point3DIdx"     blank line = next step
    traj_out.cam0-<name>.txt                          TUM "timestamp tx ty tz qx qy qz qw" (camera-to-world pose)

A step that added 3-D points is a keyframe; every added point has exactly two observations logged in that step, one in the
base keyframe and one in the current frame.  That is the (imgp0, imgp1) pair slam2.py:544-546 collects, and together with
the two poses it reconstructs the arguments of the triangulation calls at slam2.py:551-555 and :583-585 without images.

The solvers are the GPU ones of `triangulation`; there is no CPU path here.
"""
import os
import time

import numpy as np


def _blocks(path, cast):
    """Blank-line separated blocks of whitespace separated rows ('#' lines are comments)."""
    out = [[]]
    with open(path) as f:
        for ln in f:
            if ln.startswith("#"):
                continue
            ln = ln.strip()
            if not ln:
                out.append([])
                continue
            out[-1].append([cast(v) for v in ln.split()])
    if out and not out[-1]:
        out.pop()                                   # the writer ends every file with an empty line
    return out


def P_from_pose_TUM(q, l):
    """4x4 world-to-camera matrix of a TUM pose (qx qy qz qw, location): inverse of [R(q) | l]
    (the reference's transforms.P_from_pose_TUM, Work/python_libs/transforms.py:252-269)."""
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    P = np.eye(4)
    P[0:3, 0:3] = R.T
    P[0:3, 3] = -R.T.dot(np.asarray(l, dtype=np.float64))
    return P


def load_dataset(base_dir, name="slam2", cam=0):
    """Returns dict(K, dist, poses[frame] -> 4x4 P, keyframes = [dict(step, frame0, frame1, idx3d, px0, px1)])."""
    j = lambda fn: os.path.join(base_dir, fn)                                                        # noqa: E731
    calib = _blocks(j("BA_info.calibrations.cam%d.txt" % cam), float)[0][0]
    fx, fy, shear, u0, v0 = calib[0:5]
    K = np.array([[fx, shear, u0], [0., fy, v0], [0., 0., 1.]])
    dist = np.array(list(calib[5:9]) + [0.])
    pts2d = [np.array(b, dtype=np.float64).reshape(-1, 2) for b in _blocks(j("BA_info.measurements.points2D.cam%d-%s.txt" % (cam, name)), float)]
    assocs = _blocks(j("BA_info.measurements.point2D3DAssocs.cam%d-%s.txt" % (cam, name)), int)
    added = _blocks(j("BA_info.measurements.point3DAddedIdxs-%s.txt" % name), int)
    poses = []
    with open(j("traj_out.cam%d-%s.txt" % (cam, name))) as f:
        for ln in f:
            if ln.startswith("#") or not ln.strip():
                continue
            v = [float(t) for t in ln.split()]
            poses.append(P_from_pose_TUM(v[4:8], v[1:4]))
    keyframes = []
    for step, add in enumerate(added):
        ids = [r[0] for r in add]
        if not ids or step >= len(assocs):
            continue
        want = set(ids)
        obs = {}
        for frame, i2d, i3d in assocs[step]:
            if i3d in want:
                obs.setdefault(i3d, []).append((frame, i2d))
        frames = sorted({f for o in obs.values() for (f, _) in o})
        if len(frames) != 2:
            continue                                # the initial map (step 0) comes from the chessboard / init file
        f0, f1 = frames
        keep = [i for i in ids if len(obs.get(i, [])) == 2]
        if not keep or f1 >= len(poses) or f1 >= len(pts2d):
            continue
        sel = [dict(obs[i]) for i in keep]
        px0 = np.array([pts2d[f0][s[f0]] for s in sel]); px1 = np.array([pts2d[f1][s[f1]] for s in sel])
        keyframes.append({"step": step, "frame0": f0, "frame1": f1, "idx3d": np.array(keep), "px0": px0, "px1": px1})
    return {"K": K, "dist": dist, "poses": poses, "keyframes": keyframes}


def replay(dataset, fused=True, in_dtype=np.float32, out_dtype=np.float32, repeat=1):
    """
    The keyframe step of slam2.py:541-600 for every recorded keyframe:
        undistort both observations, iterative_LS, keep status == 1, [solvePnP: not on this path], iterative_LS again on
        the inliers, keep status >= 0.
    fused=True uses the pixel-input entry point (undistortion in registers); False the reference's 3-call sequence.
    Returns a list of dict(step, x, status, inliers, x_final, kept, seconds).
    """
    import triangulation as tri
    K, dist = dataset["K"], dataset["dist"]
    old = tri.output_dtype
    tri.set_triangl_output_dtype(out_dtype)
    out = []
    try:
        for kf in dataset["keyframes"]:
            P0, P1 = dataset["poses"][kf["frame0"]], dataset["poses"][kf["frame1"]]
            px0 = kf["px0"].astype(in_dtype); px1 = kf["px1"].astype(in_dtype)
            best = None
            for _ in range(repeat):
                t0 = time.perf_counter()
                if fused:
                    x, st = tri.iterative_LS_triangulation_px(px0, P0, px1, P1, K, dist)
                    inl = np.where(st == 1)[0]
                    x2, st2 = tri.iterative_LS_triangulation_px(px0[inl], P0, px1[inl], P1, K, dist)
                else:
                    n0 = tri.undistort_points(px0, K, dist); n1 = tri.undistort_points(px1, K, dist)
                    x, st = tri.iterative_LS_triangulation(n0, P0, n1, P1)
                    inl = np.where(st == 1)[0]
                    x2, st2 = tri.iterative_LS_triangulation(n0[inl], P0, n1[inl], P1)
                kept = np.where(st2 >= 0)[0]
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out.append({"step": kf["step"], "x": x, "status": st, "inliers": inl, "x_final": x2[kept], "kept": inl[kept],
                        "seconds": best})
    finally:
        tri.set_triangl_output_dtype(old)
    return out
