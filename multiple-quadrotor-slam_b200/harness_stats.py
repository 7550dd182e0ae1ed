"""
Device-side evaluation statistics of the reference's accuracy harness
(Work/triangulation_comparison/triangulation_comparison.py):

    error_vectors_3D   :179-188      error_vectors_2D   :190-203      error_rms  :205-217      robustness_stat  :242-260
    vector_stat        :219-240

`CellStatistics` accumulates one cell of test_1and2 / test_3 (:440-480: num_trials noisy repetitions of one camera pose)
without leaving the GPU: every trial's solver output stays in HBM, the squared 3-D and 2-D errors of all trials are
written into two device buffers, and the six summary numbers come from device reductions -- sums for the root-MEAN-square,
an exact radix-selection median (np.median semantics) for the root-MEDIAN-square, and counts for the false-positive /
false-negative ratios.  With keep_vectors=True the solver results of all trials stay in HBM as one (trials, N, 3) array
and vector_stat() returns what the reference computes for the last pose of a trajectory (:483-487): the per-point mean
vector and covariance matrix of the 3-D error vectors over the trials.
"""
import numpy as np

import triangl_cuda as _tc

robustness_thresh_max = 1.0          # triangulation_comparison.py:372-373
robustness_thresh_min = 1.0


def error_rms_3D(points_3D_exact, points_3D_calc):
    """error_rms(error_vectors_3D(exact, calc)): (root mean, root median, squared errors).  `points_3D_exact` is the
    harness' (N,4) homogeneous cloud or (N,3); arrays may be host or device (both in the same memory space)."""
    errors, stats = _tc.eval_errors_3d(points_3D_calc, points_3D_exact)
    n = len(points_3D_calc)
    return float(np.sqrt(stats[0] / n)), float(np.sqrt(_tc.median(errors))), errors


def vector_stat(points_3D_exact, points_3D_calc_trials):
    """vector_stat(errors_partitioned) with errors_partitioned[t] = error_vectors_3D(exact, calc[t]): per-point means (N,3)
    and covariance matrices (N,3,3) over the trials; host arrays or device buffers (both in the same memory space)."""
    return _tc.vector_stat(points_3D_calc_trials, points_3D_exact)


def robustness_stat_3D(points_3D_exact, points_3D_calc, statuses):
    """robustness_stat(errors, statuses) with errors = squared 3-D errors: (false positive ratio, false negative ratio)."""
    _, stats = _tc.eval_errors_3d(points_3D_calc, points_3D_exact, statuses, robustness_thresh_max, robustness_thresh_min,
                                  want_errors=False)
    n = float(len(points_3D_calc))
    return stats[2] / n, stats[3] / n


class CellStatistics:
    """One (trajectory, pose) cell: call add_trial() once per noisy repetition, then summary()."""

    def __init__(self, points_3D, cams, num_trials, keep_vectors=False):
        """points_3D: (N,4) exact homogeneous cloud; cams: two dicts(K, dist, rvec, tvec, points_2D_exact).
        keep_vectors: keep every trial's points on the device for vector_stat() (the harness' last-pose statistics)."""
        self.n = len(points_3D)
        self.trials = num_trials
        self.exact = _tc.to_device(np.ascontiguousarray(points_3D, dtype=np.float64))
        self.cams = []
        for c in cams:
            self.cams.append(dict(K=np.asarray(c["K"], dtype=np.float64), dist=c["dist"], rvec=c["rvec"], tvec=c["tvec"],
                                  exact2d=_tc.to_device(np.ascontiguousarray(c["points_2D_exact"], dtype=np.float64))))
        self.err3d = _tc.DeviceArray((num_trials * self.n,), np.float64)
        self.err2d = _tc.DeviceArray((num_trials * 2 * self.n,), np.float64)
        self.x_trials = _tc.DeviceArray((num_trials, self.n, 3), np.float64) if keep_vectors else None
        self.t = 0
        self.sum3 = self.sum2 = self.nan3 = self.nan2 = self.fp = self.fn = 0.0

    def add_trial(self, x, status):
        """x (N,3), status (N,): DEVICE buffers straight from a solver call."""
        assert self.t < self.trials
        lib = _tc.lib()
        stats = np.zeros(4)
        s_i32 = int(np.dtype(status.dtype).itemsize == 4)
        x32 = int(np.dtype(x.dtype) == np.float32)
        e3 = self.err3d.view(self.t * self.n, (self.n,))
        _tc.check(lib.trgl_eval_errors_3d(_tc._ptr(x), _tc._ptr(self.exact), 4, _tc._ptr(status), s_i32,
                                          robustness_thresh_max, robustness_thresh_min, _tc._ptr(e3), _tc._dp(stats),
                                          self.n, x32, _tc.MEM_DEVICE, None))
        self.sum3 += stats[0]; self.nan3 += stats[1]; self.fp += stats[2]; self.fn += stats[3]
        if self.x_trials is not None:
            if x32:
                raise ValueError("keep_vectors needs float64 solver output")
            _tc.check(lib.trgl_memcpy_d2d(self.x_trials.ptr + self.t * self.n * 24, _tc._ptr(x), self.n * 24, None))
        for ci, c in enumerate(self.cams):                       # errs2D[ti] += error_vectors_2D(...)  (cam1 then cam2)
            _, proj = _tc.reproj_error(x, c["exact2d"], c["K"], c["dist"], c["rvec"], c["tvec"], want_proj=True)
            e2 = self.err2d.view((2 * self.t + ci) * self.n, (self.n,))
            _tc.check(lib.trgl_eval_errors_2d(_tc._ptr(proj), _tc._ptr(c["exact2d"]), _tc._ptr(e2), _tc._dp(stats), self.n,
                                              0, _tc.MEM_DEVICE, None))
            self.sum2 += stats[0]; self.nan2 += stats[1]
        self.t += 1

    def vector_stat(self):
        """(means (N,3), covars (N,3,3)) of the 3-D error vectors over the trials, as host arrays (:483-487)."""
        assert self.t == self.trials and self.x_trials is not None
        means, covars = _tc.vector_stat(self.x_trials, self.exact)
        _tc.synchronize()
        return means.to_host(), covars.to_host()

    def summary(self):
        """(err3D_mean, err3D_median, err2D_mean, err2D_median, false_pos, false_neg) as in :466-480."""
        assert self.t == self.trials
        n3 = float(self.trials * self.n)
        return (float(np.sqrt(self.sum3 / n3)), float(np.sqrt(_tc.median(self.err3d))),
                float(np.sqrt(self.sum2 / (2 * n3))), float(np.sqrt(_tc.median(self.err2d))),
                self.fp / n3, self.fn / n3)
