"""
GPU mirror of the reprojection-error helpers of the reference's Work/python_libs/calibration_tools.py:

    reprojection_error(objp, imgp, cameraMatrix, distCoeffs, rvec, tvec)      -> (rms, imgp_reproj)   :116-124
    reprojection_error_ext(objp, imgp, cameraMatrix, distCoeffs, rvecs, tvecs) -> (mean_abs, rms)      :89-113

cv2.projectPoints (Rodrigues, rigid transform, perspective divide, k1 k2 p1 p2 k3 distortion, K) and the squared /
absolute residual reductions run in one fused kernel of libtriangl_cuda.  The rest of the reference module (chessboard
grids, intrinsics files, image undistortion) is file / image I/O and stays with the reference.
"""
import numpy as np

import triangl_cuda as _tc


def reprojection_error(objp, imgp, cameraMatrix, distCoeffs, rvec, tvec):
    """
    Minimalist version of "reprojection_error_ext()",
    only returns the RMS error of one image.
    """
    sums, proj = _tc.reproj_error(objp, imgp, cameraMatrix, distCoeffs, rvec, tvec, want_proj=True)
    n = float(len(proj))
    if isinstance(proj, np.ndarray):
        proj = proj.reshape(-1, 1, 2)          # cv2.projectPoints' output shape
    return np.sqrt((sums[0] + sums[1]) / n), proj


def reprojection_error_ext(objp, imgp, cameraMatrix, distCoeffs, rvecs, tvecs):
    """
    Returns the mean absolute error, and the RMS error of the reprojection
    of 3D points "objp" on the images from a camera
    with intrinsics ("cameraMatrix", "distCoeffs") and poses ("rvecs", "tvecs").
    The original 2D points should be given by "imgp".
    """
    mean_error = np.zeros(2)
    square_error = np.zeros(2)
    n_images = len(imgp)
    for i in range(n_images):
        sums, _ = _tc.reproj_error(objp[i], imgp[i], cameraMatrix, distCoeffs, rvecs[i], tvecs[i], want_proj=False)
        mean_error += sums[3:5] / len(imgp[i])
        square_error += sums[0:2] / len(imgp[i])
    return float(np.linalg.norm(mean_error / n_images)), float(np.sqrt(square_error.sum() / n_images))
