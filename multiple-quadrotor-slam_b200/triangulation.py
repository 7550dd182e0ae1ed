"""
Drop-in replacement of the reference's Work/python_libs/triangulation.py, backed by libtriangl_cuda (B200, sm_100a).

Same module name, callables, keyword names/defaults and return conventions as the reference:

    linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16) -> (x, status)   triangulation.py:6-25
    linear_LS_triangulation(u1, P1, u2, P2)                                -> (x, status)   :31-94 / triangulation.c:65-83
    iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5)            -> (x, status)   :100-195 / triangulation.c:104-161
    polynomial_triangulation(u1, P1, u2, P2)                               -> (x, status)   :198-232
    set_triangl_output_dtype(dtype)                                                         :259-267

`u1`,`u2` are (N,2) normalised image coordinates, `P1`,`P2` 3x4 or 4x4 camera matrices; `x` is a new (N,3) array of
`output_dtype`; `status` is bool (eigen / LS / polynomial) or int32 in {1,0,-1,-2,-3} (iterative, the C extension's
convention, triangulation_c/__init__.py:81).  Callers put this directory on sys.path and `import triangulation`,
exactly as slam2.py:11-19, triangulation_comparison.py:12-15 and calibrate.py:22 do.

Beyond the reference (all optional, defaults reproduce the reference):
  * set_triangl_compute_dtype(np.float32) selects the FP32-arithmetic kernels for float32 inputs/outputs;
  * set_triangl_semantics(iterative='c'|'py', eigen_rows=4|6) selects the iterative_LS control flow
    (C extension vs pure-Python fallback) and the OpenCV-4 / OpenCV-2.4 DLT system (F4);
  * device-resident inputs (triangl_cuda.DeviceArray or CUDA torch tensors) are accepted and give
    device-resident outputs;
  * resident(u1, u2) returns handles for the u1 / u2 arguments that upload the observations ONCE: the comparison harness
    runs all four solvers on the same observation set (triangulation_comparison.py:466-469); with the handles the first
    call leaves them in HBM and the other three read them there (results are host arrays as always);
  * undistort_points(imgp, cameraMatrix, distCoeffs) is cv2.undistortPoints on the GPU (bit-identical to cv2 4.13), and
    every solver has a `*_px` twin taking PIXEL coordinates plus the intrinsics, i.e. the three reference lines
        imgpnrm0 = cv2.undistortPoints(np.array([imgp0]), cameraMatrix, distCoeffs)[0]          (slam2.py:551)
        imgpnrm1 = cv2.undistortPoints(np.array([imgp1]), cameraMatrix, distCoeffs)[0]          (slam2.py:552)
        objp, status = iterative_LS_triangulation(imgpnrm0, P0, imgpnrm1, P1)                   (slam2.py:553-555)
    become   objp, status = iterative_LS_triangulation_px(imgp0, P0, imgp1, P1, cameraMatrix, distCoeffs)
    with the undistortion done in registers in front of the solve (same bits as the three-line version).
There is no CPU fallback: without the built library or a GPU every call raises.
"""
import numpy as np

import triangl_cuda as _tc

output_dtype = float
compute_dtype = np.float64
iterative_semantics = 'c'
eigen_rows = 4


def set_triangl_output_dtype(output_dtype_):
    """
    Set the datatype of the triangulated 3D point positions.
    (Default is set to "float")
    """
    global output_dtype
    output_dtype = output_dtype_


def set_triangl_compute_dtype(compute_dtype_):
    """Arithmetic type of the kernels: float64 (reference arithmetic, default) or float32 ("FP32 mode":
    only used when the inputs are float32 and the output dtype is float32)."""
    global compute_dtype
    if np.dtype(compute_dtype_) not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("compute dtype must be float32 or float64")
    compute_dtype = np.dtype(compute_dtype_).type


def set_triangl_semantics(iterative=None, eigen_rows_=None):
    global iterative_semantics, eigen_rows
    if iterative is not None:
        if iterative not in ('c', 'py'):
            raise ValueError("iterative semantics must be 'c' or 'py'")
        iterative_semantics = iterative
    if eigen_rows_ is not None:
        if eigen_rows_ not in (4, 6):
            raise ValueError("eigen_rows must be 4 or 6")
        eigen_rows = eigen_rows_


def resident(u1, u2):
    """
    "Upload once, solve many": handles to pass as (u1, u2) to any of the solvers of this module.  The first solver call
    uploads the observations and keeps them in HBM, every later call on the same handles reads them there; the results
    are host arrays exactly as with plain arrays.  (Four solvers on one observation set -- the harness' loop,
    triangulation_comparison.py:466-469 -- upload 32 instead of 128 bytes per correspondence.)
    """
    return _tc.resident(u1, u2)


def _kernel_out_dtype():
    """Storage type written by the kernel; anything but float32 is produced as float64 and cast on the host."""
    return np.float32 if np.dtype(output_dtype) == np.float32 else np.float64


def _finish(x, status, status_dtype=None):
    if isinstance(x, np.ndarray):
        if np.finfo(x.dtype) != np.finfo(output_dtype):          # triangulation.py:242-243
            x = x.astype(output_dtype)
        if status_dtype is not None and status.dtype != status_dtype:
            status = status.astype(status_dtype)
    return x, status


def linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16):
    """
    Linear Eigenvalue based (using SVD) triangulation.
    The status-vector is based on the assumption that all 3D points have finite coordinates.
    """
    x, status = _tc.linear_eigen(u1, P1, u2, P2, max_coordinate_value, eigen_rows, _kernel_out_dtype(), compute_dtype)
    return _finish(x, status)


def linear_LS_triangulation(u1, P1, u2, P2):
    """
    Linear Least Squares based triangulation.
    The status-vector will be True for all points.
    """
    x, status = _tc.linear_ls(u1, P1, u2, P2, _kernel_out_dtype(), compute_dtype)
    return _finish(x, status)


def iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5):
    """
    Iterative (Linear) Least Squares based triangulation.
    From "Triangulation", Hartley, R.I. and Sturm, P., Computer vision and image understanding, 1997.

    Additionally returns a status-vector to indicate outliers:
        1: inlier, and in front of both cameras
        0: outlier, but in front of both cameras
        -1: only in front of second camera
        -2: only in front of first camera
        -3: not in front of any camera
    Outliers are selected based on non-convergence of depth, and on negativity of depths (=> behind camera(s)).
    """
    sem = _tc.ITER_PY if iterative_semantics == 'py' else _tc.ITER_C
    x, status = _tc.iterative_ls(u1, P1, u2, P2, tolerance, sem, _kernel_out_dtype(), compute_dtype)
    # the C extension returns int32 (triangulation_c/__init__.py:81), the Python fallback a platform int (:125)
    return _finish(x, status, np.int64 if iterative_semantics == 'py' else None)


def polynomial_triangulation(u1, P1, u2, P2):
    """
    Polynomial (Optimal) triangulation.
    Uses Linear-Eigen for final triangulation.
    The status-vector is based on the assumption that all 3D points have finite coordinates.
    """
    x, status, all_nan = _tc.polynomial(u1, P1, u2, P2, None, 1.e16, eigen_rows, _kernel_out_dtype(), compute_dtype)
    if all_nan and len(status) >= 8:
        # every corrected point is NaN (F == 0): the reference re-estimates F from the matches with the normalised
        # 8-point algorithm and corrects again (triangulation.py:227-229)
        F = _tc.fundamental_8point(u1, u2)
        x, status, _ = _tc.polynomial(u1, P1, u2, P2, F, 1.e16, eigen_rows, _kernel_out_dtype(), compute_dtype,
                                      check_all_nan=False)
    return _finish(x, status)


def P_from_R_and_t(R, t):
    """4x4 P = [R | t; 0 0 0 1]  (the reference's transforms.P_from_R_and_t, transforms.py:156-168)."""
    P = np.eye(4)
    P[0:3, 0:3] = R
    P[0:3, 3:4] = np.asarray(t, dtype=np.float64).reshape(3, 1)
    return P


def P_from_rvec_and_tvec(rvec, tvec):
    """4x4 camera matrix from OpenCV's (rvec, tvec) (transforms.py:249 = P_from_R_and_t(cv2.Rodrigues(rvec)[0], tvec)),
    the form in which slam2.py:554-555 hands the two poses to the solvers.  Rodrigues' formula on the host."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.sqrt(r.dot(r)))
    if th < np.finfo(np.float64).eps:
        R = np.eye(3)
    else:
        k = r / th
        Kx = np.array([[0., -k[2], k[1]], [k[2], 0., -k[0]], [-k[1], k[0], 0.]])
        R = np.cos(th) * np.eye(3) + (1. - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx
    return P_from_R_and_t(R, tvec)


def undistort_points(imgp, cameraMatrix, distCoeffs=None):
    """cv2.undistortPoints(imgp, cameraMatrix, distCoeffs) without R / P: (N,2) normalised image coordinates in the
    dtype of `imgp` (any (...,2) shape is accepted, so the cv2-4.x (N,1,2) vs cv2-2.x (1,N,2) pitfall of the
    reference idiom `cv2.undistortPoints(np.array([imgp]), K, d)[0]` -- SURVEY.md F7 -- does not arise)."""
    return _tc.undistort_points(imgp, cameraMatrix, distCoeffs)


def _intr(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2):
    return _tc.Intrinsics(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2)


def linear_eigen_triangulation_px(imgp1, P1, imgp2, P2, cameraMatrix, distCoeffs=None, max_coordinate_value=1.e16,
                                  cameraMatrix2=None, distCoeffs2=None):
    """linear_eigen_triangulation on pixel coordinates (undistortion fused in front of the solve)."""
    x, status = _tc.linear_eigen(imgp1, P1, imgp2, P2, max_coordinate_value, eigen_rows, _kernel_out_dtype(),
                                 compute_dtype, pixel=_intr(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2))
    return _finish(x, status)


def multiview_LS_triangulation(us, Ps, valid=None, min_views=2):
    """
    Linear Least Squares triangulation from m >= 2 views (not in the reference, whose calls are all two-view; same
    conventions): "us" (m, N, 2) normalised observations, "Ps" m camera matrices (3x4 or 4x4), "valid" (m, N) which
    view observes which point (default: all).  Returns x (N, 3) of the output dtype and a bool status-vector that is True
    where at least "min_views" views observe the point.  With two views and no mask: linear_LS_triangulation.
    """
    x, status = _tc.multiview_ls(us, Ps, valid, min_views, _kernel_out_dtype())
    return _finish(x, status)


def linear_LS_triangulation_px(imgp1, P1, imgp2, P2, cameraMatrix, distCoeffs=None, cameraMatrix2=None,
                               distCoeffs2=None):
    """linear_LS_triangulation on pixel coordinates (undistortion fused in front of the solve)."""
    x, status = _tc.linear_ls(imgp1, P1, imgp2, P2, _kernel_out_dtype(), compute_dtype,
                              pixel=_intr(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2))
    return _finish(x, status)


def iterative_LS_triangulation_px(imgp1, P1, imgp2, P2, cameraMatrix, distCoeffs=None, tolerance=3.e-5,
                                  cameraMatrix2=None, distCoeffs2=None):
    """iterative_LS_triangulation on pixel coordinates: the SLAM keyframe call pattern slam2.py:551-555 in one call."""
    sem = _tc.ITER_PY if iterative_semantics == 'py' else _tc.ITER_C
    x, status = _tc.iterative_ls(imgp1, P1, imgp2, P2, tolerance, sem, _kernel_out_dtype(), compute_dtype,
                                 pixel=_intr(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2))
    return _finish(x, status, np.int64 if iterative_semantics == 'py' else None)


def polynomial_triangulation_px(imgp1, P1, imgp2, P2, cameraMatrix, distCoeffs=None, cameraMatrix2=None,
                                distCoeffs2=None):
    """polynomial_triangulation on pixel coordinates (undistortion fused in front of the Hartley-Sturm correction)."""
    intr = _intr(cameraMatrix, distCoeffs, cameraMatrix2, distCoeffs2)
    x, status, all_nan = _tc.polynomial(imgp1, P1, imgp2, P2, None, 1.e16, eigen_rows, _kernel_out_dtype(),
                                        compute_dtype, pixel=intr)
    if all_nan and len(status) >= 8:
        # the rare 8-point fallback needs the normalised points as arrays: undistort, then the plain path
        u1 = _tc.undistort_points(imgp1, intr.K1, intr.d1); u2 = _tc.undistort_points(imgp2, intr.K2, intr.d2)
        return polynomial_triangulation(u1, P1, u2, P2)
    return _finish(x, status)


def triangulate_and_evaluate(solver, u1, P1, u2, P2, min_status=0, max_sq_err=np.inf, **kwargs):
    """
    The call pattern of the SLAM keyframe step (slam2.py:553-563) and of the comparison harness
    (triangulation_comparison.py:466-480) in one go: solve, then re-project into both cameras and build the
    good-point mask `status > min_status and errors <= max_sq_err and in front of both cameras`.
    Returns x, status, good, (rms_err1, rms_err2) over the good points.
    """
    x, status = solver(u1, P1, u2, P2, **kwargs)
    _, _, good, sums = _tc.pair_reproj(x, u1, P1, u2, P2, status, min_status, max_sq_err, want_errors=False)
    ngood = max(sums[2], 1.0)
    return x, status, good, (float(np.sqrt(sums[0] / ngood)), float(np.sqrt(sums[1] / ngood)))
