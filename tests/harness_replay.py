"""
Replay of one cell of the reference's accuracy harness (Work/triangulation_comparison/
triangulation_comparison.py: test_1and2 :403-515, test_3 :517-627) with a pluggable solver backend.
Used by the CPU oracle tests and by the GPU parity tests, so both are checked against the reference's
own golden result files (tests/golden/golden_cells.json, extracted by oracle/make_golden.py).
"""
import numpy as np

import synthetic_rig as rig

ROBUSTNESS_THRESH = 1.0     # triangulation_comparison.py:372-373


def error_rms(error_vectors):                                   # :205-217
    errors = np.sum(np.concatenate(error_vectors) ** 2, axis=1)
    return np.sqrt(np.mean(errors)), np.sqrt(np.median(errors)), errors


def robustness_stat(errors, statuses):                          # :242-260
    statuses = np.concatenate(statuses)
    positives_est = statuses > 0
    fp = np.logical_and((errors <= ROBUSTNESS_THRESH) == False, positives_est)      # noqa: E712
    fn = np.logical_and(errors <= ROBUSTNESS_THRESH, positives_est == False)        # noqa: E712
    return np.mean(fp), np.mean(fn)


def vector_stat(error_vectors):                                 # :219-240
    """Mean vector and (population) covariance matrix over the trials, per point.  error_vectors: (num_trials, N, d)."""
    ev = np.asarray(error_vectors, dtype=np.float64)
    means = ev.mean(axis=0)
    dev = ev - means[None]
    covars = np.einsum('tni,tnj->nij', dev, dev) / ev.shape[0]
    return means, covars


def replay_cell(solvers, pose, points_3D=None, num_trials=100, rseed=rig.RSEED, sigma=0.8, discretized=True,
                k1=0.3, offset=40., return_vectors=False, on_trial=None):
    """
    solvers: list of callables (u1, P1, u2, P2) -> (x, status).  pose = (sideways, towards, angle) of cam 2.
    Returns a dict of the six summary statistics, each a list over solvers.
    """
    if points_3D is None:
        points_3D = rig.finite_3D_points(4)
    cam1, cam2 = rig.Camera(), rig.Camera()
    cam1.camera_pose(offset)
    cam2.camera_pose(offset, *pose)
    for cam in (cam1, cam2):
        cam.camera_intrinsics((640, 480), k1)
        cam.project_points(points_3D)
    errs3D = [[] for _ in solvers]; errs2D = [[] for _ in solvers]; statuses = [[] for _ in solvers]
    np.random.seed(rseed)                                        # reset_random() :355-363
    for _ in range(num_trials):
        cam1.apply_noise(sigma, discretized)
        cam2.apply_noise(sigma, discretized)
        u1 = cam1.normalized_points()
        u2 = cam2.normalized_points()
        for ti, solver in enumerate(solvers):
            x, status = solver(u1, cam1.P, u2, cam2.P)
            if on_trial is not None:
                on_trial(ti, x, status)                                              # e.g. device-side accumulation
            x = np.asarray(x, dtype=np.float64)
            errs3D[ti].append(x - points_3D[:, 0:3])                                 # :179-188
            xh = np.concatenate([x, np.ones((len(x), 1))], axis=1)
            with np.errstate(all='ignore'):
                errs2D[ti].append(cam1.project_points(xh, False) - cam1.points_2D_exact)   # :190-203
                errs2D[ti].append(cam2.project_points(xh, False) - cam2.points_2D_exact)
            statuses[ti].append(np.asarray(status))
    out = {k: [] for k in ("err3D_mean_summary", "err3D_median_summary", "err2D_mean_summary",
                           "err2D_median_summary", "false_pos_summary", "false_neg_summary")}
    for ti in range(len(solvers)):
        with np.errstate(all='ignore'):
            m3, md3, errors = error_rms(errs3D[ti])
            m2, md2, _ = error_rms(errs2D[ti])
            fp, fn = robustness_stat(errors, statuses[ti])
        for k, v in zip(out, (m3, md3, m2, md2, fp, fn)):
            out[k].append(float(v))
    if return_vectors:
        out["error_vectors_3D"] = [np.array(e) for e in errs3D]                     # errors_partitioned, :483
    return out
