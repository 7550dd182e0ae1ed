"""
File formats either side of the triangulation path: ASCII .pcd point clouds (map
export of the triangulated points) and TUM camera trajectories (the poses the camera matrices come from).

Mirrors the reference's `Work/python_libs/dataset_tools.py` for this path -- same function names, argument meaning,
return conventions and error behaviour:
    load_cam_trajectory_TUM / save_cam_trajectory_TUM        dataset_tools.py:71-115
    load_3D_points_from_pcd_file / save_3D_points_to_pcd_file dataset_tools.py:118-272
    convert_cam_poses_to_cam_trajectory_TUM                   dataset_tools.py:275-294 (+ transforms.py:271-283)
Host-side text I/O only (nothing here is data-parallel device work); parsing and colour packing are vectorised NumPy
instead of per-line Python loops, so a 10 M-point map writes in seconds.
"""
import io

import numpy as np

_PCD_HEADER = ("# .PCD v.7 - Point Cloud Data file format\n"
               "VERSION .7\n"
               "FIELDS x y z{rgb}\n"
               "SIZE 4 4 4{four}\n"
               "TYPE F F F{f}\n"
               "COUNT 1 1 1{one}\n"
               "WIDTH {n}\n"
               "HEIGHT 1\n"
               "VIEWPOINT 0 0 0 1 0 0 0\n"
               "POINTS {n}\n"
               "DATA ascii\n")


def _trajectory_arrays(rows, normalize_quaternions):
    """(timestps (N,), locations (N,3), quaternions (N,4)) float64; empty arrays of those shapes for no rows."""
    rows = np.asarray(rows, dtype=float).reshape(-1, 8)
    quaternions = rows[:, 4:8].copy()
    if normalize_quaternions and len(rows):
        quaternions /= np.linalg.norm(quaternions, axis=1, keepdims=True)
    return rows[:, 0].copy(), rows[:, 1:4].copy(), quaternions


def load_cam_trajectory_TUM(filename):
    """
    Camera trajectory in the TUM RGB-D format ("timestamp tx ty tz qx qy qz qw", '#' comments, blank lines; commas and
    tabs are accepted as separators).  Returns float64 arrays "timestps" (N,), "locations" (N,3) and unit-normalised
    "quaternions" (N,4).  A line that does not hold exactly 8 numbers raises ValueError, like the reference's unpacking.
    """
    rows = []
    with open(filename, 'r') as f:
        for line in f.read().replace(',', ' ').replace('\t', ' ').split('\n'):
            line = line.strip()
            if not line or line[0] == '#':
                continue
            values = [float(w) for w in line.split(' ')]       # consecutive blanks give '' -> ValueError, as upstream
            if len(values) != 8:
                raise ValueError("expected 8 values per pose, got %d: %r" % (len(values), line))
            rows.append(values)
    return _trajectory_arrays(rows, normalize_quaternions=True)


def save_cam_trajectory_TUM(filename, cam_trajectory):
    """Writes ("timestps", "locations", "quaternions") in the TUM format; numbers use Python's shortest round-trip repr."""
    timestps, locations, quaternions = cam_trajectory
    out = io.StringIO()
    out.write("# Format: timestamp tx ty tz qx qy qz qw\n")
    out.write("# Where translations and quaternions are defined in world coordinates (=> inverse of pose)\n")
    for t, l, q in zip(timestps, locations, quaternions):
        out.write(' '.join(str(v) for v in (t,) + tuple(l) + tuple(q)))
        out.write('\n')
    with open(filename, 'w') as f:
        f.write(out.getvalue())


def pose_TUM_from_P(P):
    """(quaternion (qx,qy,qz,qw), location) of the camera-to-world pose of a world-to-camera matrix P = [R | t]
    (transforms.py:271-283): location = -R^T t, quaternion of R^T.  Inverse of slam_replay.P_from_pose_TUM."""
    P = np.asarray(P, dtype=float)
    R = P[0:3, 0:3].T
    loc = -R @ P[0:3, 3]
    # Shepperd's method: pick the largest of (w, x, y, z) to divide by
    tr = np.trace(R)
    cand = np.array([tr, R[0, 0], R[1, 1], R[2, 2]])
    k = int(np.argmax(cand))
    if k == 0:
        w = 0.5 * np.sqrt(1.0 + tr)
        q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    else:
        i = k - 1; j = (i + 1) % 3; m = (i + 2) % 3
        s = 0.5 * np.sqrt(1.0 + R[i, i] - R[j, j] - R[m, m])
        q = np.empty(4)
        q[i] = s
        q[j] = (R[j, i] + R[i, j]) / (4 * s)
        q[m] = (R[m, i] + R[i, m]) / (4 * s)
        q[3] = (R[m, j] - R[j, m]) / (4 * s)
    if q[3] < 0:
        q = -q
    return q, loc


def convert_cam_poses_to_cam_trajectory_TUM(Ps, fps=30):
    """Camera matrices "Ps" (None entries are skipped) -> TUM trajectory; the first pose has timestamp 1 / fps."""
    rows = []
    for i, P in enumerate(Ps):
        if P is None:
            continue
        q, loc = pose_TUM_from_P(P)
        rows.append([float(1 + i) / fps] + list(loc) + list(q))
    return _trajectory_arrays(rows, normalize_quaternions=False)


def load_3D_points_from_pcd_file(filename, use_alpha=False):
    """
    ASCII .pcd -> (points float32 (N,3), colors uint8 (N,3|4) as (B,G,R[,A]) or None, found_alpha).
    Supported headers, as upstream: FIELDS "x y z" or "x y z rgb", HEIGHT 1, DATA ascii; anything else raises ValueError
    with the reference's messages; header entries are looked for in the order FIELDS, WIDTH, HEIGHT, DATA.
    """
    with open(filename, 'r') as f:
        lines = f.read().split('\n')
    num_points = 0
    use_colors = False
    wanted = ["FIELDS", "WIDTH", "HEIGHT", "DATA"]
    data_start = None
    for i, line in enumerate(lines):
        words = line.split(' ')
        if not wanted or words[0] != wanted[0]:
            continue
        entry = wanted.pop(0)
        if entry == "FIELDS":
            fields = words[1:]
            if fields == ['x', 'y', 'z']:
                use_colors = False
            elif fields == ['x', 'y', 'z', 'rgb']:
                use_colors = True
            else:
                raise ValueError("The following 'FIELDS' config in the .pcd-file is not supported: %s" % fields)
        elif entry == "WIDTH":
            num_points = int(words[1])
        elif entry == "HEIGHT":
            if int(words[1]) != 1:
                raise ValueError("Organized point clouds in the .pcd-file are not supported.")
        else:
            if words[1] != "ascii":
                raise ValueError("The following 'DATA' config in the .pcd-file is not supported: '%s'" % words[1])
            data_start = i + 1
            break
    if data_start is None:
        raise ValueError("The .pcd-file did not include all necessary header entries.")
    body = [l for l in lines[data_start: data_start + num_points]]
    if len(body) < num_points:
        raise ValueError("The .pcd-file did not include all advertised points. (%s instead of %s)" %
                         (len(body), num_points))
    if num_points == 0:
        return np.zeros((0, 3), dtype=np.float32), None, False
    ncol = 4 if use_colors else 3
    flat = np.array(' '.join(body).split(), dtype=np.float64)
    if flat.size != num_points * ncol:
        raise ValueError("The .pcd-file did not include all advertised points. (%s values instead of %s)" %
                         (flat.size, num_points * ncol))
    table = flat.reshape(num_points, ncol).astype(np.float32)
    if not use_colors:
        return table, None, False
    # the packed colour is the float32 whose little-endian bytes are (B, G, R, A)
    colors = np.ascontiguousarray(table[:, 3]).view(np.uint8).reshape(num_points, 4)
    found_alpha = True
    if not use_alpha:
        colors = colors[:, 0:3]
    return np.ascontiguousarray(table[:, 0:3]), np.ascontiguousarray(colors), found_alpha


def save_3D_points_to_pcd_file(filename, points, colors=None):
    """
    (N,3) points (any float dtype, e.g. the `x` a triangulation call returned) -> ASCII .pcd, optionally with uint8
    colours (B,G,R) or (B,G,R,A) packed into one float32 per point.  Alpha's two low bits are forced to 0b01 (so the
    packed float is never NaN / Inf / denormal); without alpha the maximum 0xFD is stored.  Values are printed with
    "%.8e", enough to recover the float32 -- and therefore the colour -- exactly.
    """
    points = np.asarray(points).astype(np.float32).reshape(-1, 3)
    has = colors is not None
    header = _PCD_HEADER.format(rgb=" rgb" * has, four=" 4" * has, f=" F" * has, one=" 1" * has, n=len(points))
    if has:
        colors = np.asarray(colors)
        if colors.dtype != np.uint8:
            colors = colors.astype(np.uint8)
        bgra = np.empty((len(points), 4), dtype=np.uint8)
        bgra[:, 0:3] = colors[:, 0:3]
        if colors.shape[1] == 4:
            bgra[:, 3] = (colors[:, 3] & 0b11111100) | 0b01
        else:
            bgra[:, 3] = 0xFD
        packed = bgra.view(np.float32).reshape(len(points), 1)
        table = np.concatenate((points, packed), axis=1)
    else:
        table = points
    out = io.StringIO()
    out.write(header)
    if len(table):
        np.savetxt(out, table.astype(np.float64), fmt="%.8e", delimiter=' ')     # float32 values, printed exactly as upstream
    else:
        out.write("\n")
    with open(filename, 'w') as f:
        f.write(out.getvalue())
