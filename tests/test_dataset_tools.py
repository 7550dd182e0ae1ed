"""PCD / TUM file formats: our dataset_tools against what the reference's own
dataset_tools.py loads and writes for files it ships (tests/golden/formats.npz, made by oracle/make_golden.py formats)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-quadrotor-slam_b200"))
import dataset_tools as dt          # noqa: E402
import slam_replay                  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "formats.npz"))


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(str(text))
    return str(p)


def test_load_pcd_with_colours_equals_reference(gold, tmp_path):
    f = _write(tmp_path, "map.pcd", gold["pcd_text"])
    pts, cols, alpha = dt.load_3D_points_from_pcd_file(f, use_alpha=True)
    assert pts.dtype == np.float32 and cols.dtype == np.uint8
    assert np.array_equal(pts, gold["points"]) and np.array_equal(cols, gold["colors_bgra"]) and alpha == bool(gold["found_alpha"])
    pts3, cols3, alpha3 = dt.load_3D_points_from_pcd_file(f)
    assert np.array_equal(pts3, gold["points"]) and np.array_equal(cols3, gold["colors_bgr"]) and alpha3


def test_load_pcd_without_colours_equals_reference(gold, tmp_path):
    f = _write(tmp_path, "init.pcd", gold["pcd_plain_text"])
    pts, cols, alpha = dt.load_3D_points_from_pcd_file(f)
    assert np.array_equal(pts, gold["points_plain"]) and cols is None and alpha is False


def test_save_pcd_is_byte_identical_to_reference(gold, tmp_path):
    f = str(tmp_path / "o.pcd")
    dt.save_3D_points_to_pcd_file(f, gold["points"], gold["colors_bgra"])
    assert open(f).read() == str(gold["saved_bgra"])
    dt.save_3D_points_to_pcd_file(f, gold["points"], gold["colors_bgr"])
    assert open(f).read() == str(gold["saved_bgr"])
    dt.save_3D_points_to_pcd_file(f, gold["points_plain"])
    assert open(f).read() == str(gold["saved_plain"])
    # a float64 triangulation result is written as its float32 rounding
    dt.save_3D_points_to_pcd_file(f, gold["points_plain"].astype(np.float64))
    assert open(f).read() == str(gold["saved_plain"])


def test_pcd_round_trip_and_alpha_bits(tmp_path):
    rng = np.random.RandomState(5)
    pts = (rng.randn(1000, 3) * 10).astype(np.float32)
    cols = rng.randint(0, 256, (1000, 4)).astype(np.uint8)
    f = str(tmp_path / "rt.pcd")
    dt.save_3D_points_to_pcd_file(f, pts, cols)
    p2, c2, alpha = dt.load_3D_points_from_pcd_file(f, use_alpha=True)
    assert np.array_equal(p2, pts) and alpha
    assert np.array_equal(c2[:, 0:3], cols[:, 0:3])
    assert np.array_equal(c2[:, 3], (cols[:, 3] & 0xFC) | 1)          # two low bits forced to 0b01
    dt.save_3D_points_to_pcd_file(f, pts, cols[:, 0:3])
    _, c3, _ = dt.load_3D_points_from_pcd_file(f, use_alpha=True)
    assert np.all(c3[:, 3] == 0xFD)
    dt.save_3D_points_to_pcd_file(f, np.zeros((0, 3)))
    p0, c0, a0 = dt.load_3D_points_from_pcd_file(f)
    assert p0.shape == (0, 3) and p0.dtype == np.float32 and c0 is None and a0 is False


@pytest.mark.parametrize("bad, msg", [
    ("FIELDS x y z normal_x\nWIDTH 1\nHEIGHT 1\nDATA ascii\n0 0 0 0\n", "'FIELDS' config"),
    ("FIELDS x y z\nWIDTH 1\nHEIGHT 2\nDATA ascii\n0 0 0\n", "Organized point clouds"),
    ("FIELDS x y z\nWIDTH 1\nHEIGHT 1\nDATA binary\n", "'DATA' config"),
    ("FIELDS x y z\nWIDTH 1\nHEIGHT 1\n", "necessary header entries"),
    ("FIELDS x y z\nWIDTH 3\nHEIGHT 1\nDATA ascii\n0 0 0", "advertised points"),
])
def test_pcd_unsupported_headers_raise_like_reference(tmp_path, bad, msg):
    f = _write(tmp_path, "bad.pcd", bad)
    with pytest.raises(ValueError, match=msg):
        dt.load_3D_points_from_pcd_file(f)


def test_load_trajectory_equals_reference(gold, tmp_path):
    f = _write(tmp_path, "traj.txt", gold["traj_text"])
    ts, locs, quats = dt.load_cam_trajectory_TUM(f)
    assert np.array_equal(ts, gold["timestps"]) and np.array_equal(locs, gold["locations"])
    assert np.array_equal(quats, gold["quaternions"])
    assert np.allclose(np.linalg.norm(quats, axis=1), 1.0, atol=1e-15)
    # commas / tabs as separators, comments, blank lines
    f2 = _write(tmp_path, "t2.txt", "# c\n\n1.0,0\t0 0 0,0 0 2\n")
    ts2, l2, q2 = dt.load_cam_trajectory_TUM(f2)
    assert ts2.tolist() == [1.0] and q2.tolist() == [[0, 0, 0, 1.0]]
    e = dt.load_cam_trajectory_TUM(_write(tmp_path, "e.txt", "# nothing\n"))
    assert e[0].shape == (0,) and e[1].shape == (0, 3) and e[2].shape == (0, 4)
    with pytest.raises(ValueError):
        dt.load_cam_trajectory_TUM(_write(tmp_path, "b.txt", "1 2 3\n"))


def test_save_trajectory_is_byte_identical_to_reference(gold, tmp_path):
    f = str(tmp_path / "o.txt")
    dt.save_cam_trajectory_TUM(f, (gold["timestps"][:25], gold["locations"][:25], gold["quaternions"][:25]))
    assert open(f).read() == str(gold["saved_traj"])


def test_poses_to_trajectory_inverts_P_from_pose_TUM(gold):
    quats, locs = gold["quaternions"], gold["locations"]
    Ps = [slam_replay.P_from_pose_TUM(q, l) for q, l in zip(quats, locs)]
    Ps[3] = None                                                     # skipped, its timestamp slot stays empty
    ts, l2, q2 = dt.convert_cam_poses_to_cam_trajectory_TUM(Ps, fps=50)
    keep = np.arange(len(quats)) != 3
    assert np.allclose(ts, (1 + np.arange(len(quats)))[keep] / 50.0)
    assert np.allclose(l2, locs[keep], atol=1e-12)
    sign = np.sign(np.sum(q2 * quats[keep], axis=1))[:, None]        # q and -q are the same rotation
    assert np.allclose(q2 * sign, quats[keep], atol=1e-9)
