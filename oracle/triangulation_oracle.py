"""
CPU oracle for the batched two-view triangulation hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement (vectorised over points) of the reference's algorithm.  It is the
*checker* the CUDA path is compared against; nothing under `multiple-quadrotor-slam_b200/` may import
it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
use it.

Parity pinning (see DESIGN.md "Oracle"):
  * linear_LS / iterative_LS(C semantics) / linear_eigen(rows=6): reproduce sampled cells of the
    reference's own golden result files Work/triangulation_comparison/test_1and2.mat and test_3.mat to
    >= 9 digits (tests/test_oracle_golden.py, fixtures in tests/golden/golden_cells.json).
  * all four solvers: agree with the reference's pure-Python bodies (Work/python_libs/triangulation.py
    lines 1-233, exec'd in the build container by oracle/make_golden.py) on the committed per-point
    fixtures tests/golden/ref_py_*.npz.
  * polynomial: pinned against cv2 4.13 `correctMatches` (the only executable statement of that
    third-party algorithm; OpenCV-2 behaviour is NOT reproducible -> "parity unpinned vs OpenCV 2").

Reference statements followed (paths relative to /root/reference):
  build_Ab                  Work/python_libs/triangulation_c/triangulation.c:24-42
  lstsq_minnorm             cvSolve(..., DECOMP_SVD) call sites triangulation.c:81,130 / triangulation.py:92,151
  linear_LS_triangulation   triangulation.c:65-83, triangulation.py:31-94, triangulation_c/__init__.py:18-47
  iterative_LS_triangulation triangulation.c:104-161 (semantics='c'), triangulation.py:100-195 (semantics='py')
  linear_eigen_triangulation triangulation.py:6-25 (cv2.triangulatePoints; rows=4 is OpenCV 4, rows=6 OpenCV 2.4)
  polynomial_triangulation  triangulation.py:198-232 (cv2.correctMatches = Hartley-Sturm, H&Z alg. 12.1)
  reprojection_error        Work/python_libs/calibration_tools.py:116-124 (cv2.projectPoints)
  undistort_points          cv2.undistortPoints call sites Work/SLAM/application/own/slam2.py:551-552,
                            Work/triangulation_comparison/triangulation_comparison.py:164-173 (third-party: OpenCV;
                            pinned bit-exactly to cv2 4.13 output, tests/golden/cv2_undistort.npz)
"""
import numpy as np

DBL_EPSILON = np.finfo(np.float64).eps

output_dtype = float


def set_triangl_output_dtype(output_dtype_):
    """triangulation.py:261-267"""
    global output_dtype
    output_dtype = output_dtype_


def _as_f64(u):
    # triangulation_c/__init__.py:32-33 : inputs are up-cast to float64, all arithmetic is float64
    return np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(-1, 2))


def build_Ab(u1, P1, u2, P2):
    """A (N,4,3), b (N,4): triangulation.c:30-40 (rows cam1.x, cam1.y, cam2.x, cam2.y)."""
    u1 = _as_f64(u1); u2 = _as_f64(u2)
    P1 = np.asarray(P1, dtype=np.float64); P2 = np.asarray(P2, dtype=np.float64)
    n = len(u1)
    M = np.empty((n, 4, 4))
    M[:, 0, :] = u1[:, 0:1] * P1[2, :] - P1[0, :]
    M[:, 1, :] = u1[:, 1:2] * P1[2, :] - P1[1, :]
    M[:, 2, :] = u2[:, 0:1] * P2[2, :] - P2[0, :]
    M[:, 3, :] = u2[:, 1:2] * P2[2, :] - P2[1, :]
    return M[:, :, 0:3].copy(), -M[:, :, 3]


def lstsq_minnorm(A, b):
    """
    Minimum-norm least squares, OpenCV SVD back-substitution rule: singular values
    w_i <= 2*DBL_EPSILON*sum(w) are dropped (cv::SVBackSubst).  A (N,4,3), b (N,4) -> (N,3).
    Non-finite systems give NaN (cvSolve's failure is ignored by the reference: triangulation.c:81).
    """
    n = len(A)
    x = np.full((n, 3), np.nan)
    ok = np.isfinite(A).all(axis=(1, 2)) & np.isfinite(b).all(axis=1)
    if ok.any():
        U, w, Vt = np.linalg.svd(A[ok], full_matrices=False)
        thr = 2 * DBL_EPSILON * w.sum(axis=1, keepdims=True)
        with np.errstate(divide='ignore'):
            winv = np.where(w > thr, 1.0 / w, 0.0)
        utb = np.einsum('nij,ni->nj', U, b[ok])
        x[ok] = np.einsum('nji,nj->ni', Vt, winv * utb)
    return x


def linear_LS_triangulation(u1, P1, u2, P2):
    A, b = build_Ab(u1, P1, u2, P2)
    x = lstsq_minnorm(A, b)
    return x.astype(output_dtype), np.ones(len(x), dtype=bool)


def multiview_LS_triangulation(us, Ps, valid=None, min_views=2):
    """
    m-view generalisation of linear_LS_triangulation (no reference statement -- the definition is
    the stacked system of triangulation.c:30-40 for every observing view, solved like cvSolve(DECOMP_SVD) at :81).
    us (m,N,2), Ps m matrices, valid (m,N) or None.  Rows of non-observing views are zero rows: they change neither the
    singular values nor the minimum-norm solution.
    """
    us = np.asarray(us, dtype=np.float64)
    m, n = us.shape[0], us.shape[1]
    A = np.zeros((n, 2 * m, 3)); b = np.zeros((n, 2 * m))
    seen = np.ones((m, n), dtype=bool) if valid is None else np.asarray(valid).astype(bool)
    for v in range(m):
        P = np.asarray(Ps[v], dtype=np.float64)
        r0 = us[v][:, 0:1] * P[2, :] - P[0, :]
        r1 = us[v][:, 1:2] * P[2, :] - P[1, :]
        sel = seen[v]
        A[sel, 2 * v, :] = r0[sel, 0:3];     b[sel, 2 * v] = -r0[sel, 3]
        A[sel, 2 * v + 1, :] = r1[sel, 0:3]; b[sel, 2 * v + 1] = -r1[sel, 3]
    x = lstsq_minnorm(A, b)
    return x.astype(output_dtype), seen.sum(axis=0) >= min_views


def iterative_LS_core(u1, P1, u2, P2, tolerance=3.e-5, semantics='c'):
    """
    Returns x (N,3) f64, status (N,) int, n_solves (N,) and the convergence margin
    min over the decisive test of | |dd| - tol | (for knife-edge classification).
    """
    P1 = np.asarray(P1, dtype=np.float64); P2 = np.asarray(P2, dtype=np.float64)
    A, b = build_Ab(u1, P1, u2, P2)
    n = len(A)
    x = np.empty((n, 3))
    d1 = np.ones(n); d2 = np.ones(n)
    d1n = np.ones(n); d2n = np.ones(n)
    it = np.full(n, 10 if semantics == 'c' else 9)      # loop variable after the for (F2 in SURVEY.md)
    margin = np.full(n, np.inf)
    active = np.arange(n)
    for i in range(10):
        if len(active) == 0:
            break
        xa = lstsq_minnorm(A[active], b[active])
        x[active] = xa
        a1 = xa @ P1[2, 0:3] + P1[2, 3]                 # triangulation.c:133-134
        a2 = xa @ P2[2, 0:3] + P2[2, 3]
        d1n[active] = a1; d2n[active] = a2
        e1 = np.abs(a1 - d1[active]); e2 = np.abs(a2 - d2[active])
        with np.errstate(invalid='ignore'):
            conv = (e1 <= tolerance) & (e2 <= tolerance)
            margin[active] = np.minimum(margin[active],
                                        np.minimum(np.abs(e1 - tolerance), np.abs(e2 - tolerance)))
            brk = conv | (a1 == 0) | (a2 == 0) if semantics == 'c' else conv
        it[active[brk]] = i
        cont = active[~brk]
        with np.errstate(divide='ignore', invalid='ignore'):
            s1 = 1.0 / d1n[cont]; s2 = 1.0 / d2n[cont]   # triangulation.c:143-146 (cumulative)
            A[cont, 0:2, :] *= s1[:, None, None]; A[cont, 2:4, :] *= s2[:, None, None]
            b[cont, 0:2] *= s1[:, None];          b[cont, 2:4] *= s2[:, None]
        d1[cont] = d1n[cont]; d2[cont] = d2n[cont]
        active = cont
    with np.errstate(invalid='ignore'):
        status = ((it < 10) & (d1n > 0) & (d2n > 0)).astype(np.int32)   # triangulation.c:154-159
        status -= (d1n <= 0)
        status -= 2 * (d2n <= 0)
    n_solves = np.minimum(it + 1, 10)
    return x, status, n_solves, margin


def iterative_LS_triangulation(u1, P1, u2, P2, tolerance=3.e-5, semantics='c'):
    x, status, _, _ = iterative_LS_core(u1, P1, u2, P2, tolerance, semantics)
    if semantics == 'py':
        status = status.astype(np.int64)
    return x.astype(output_dtype), status


def eigen_homogeneous(u1, P1, u2, P2, rows=4):
    """
    Unit right-singular vector of the smallest singular value of the per-point DLT matrix.
    rows=4: OpenCV >= 3 (x*P[2]-P[0], y*P[2]-P[1] per view); rows=6: OpenCV 2.4 adds x*P[1]-y*P[0].
    """
    u1 = _as_f64(u1); u2 = _as_f64(u2)
    P1 = np.asarray(P1, dtype=np.float64)[0:3, 0:4]; P2 = np.asarray(P2, dtype=np.float64)[0:3, 0:4]
    n = len(u1)
    per = rows // 2
    M = np.empty((n, rows, 4))
    for j, (u, P) in enumerate(((u1, P1), (u2, P2))):
        M[:, per * j + 0, :] = u[:, 0:1] * P[2, :] - P[0, :]
        M[:, per * j + 1, :] = u[:, 1:2] * P[2, :] - P[1, :]
        if per == 3:
            M[:, per * j + 2, :] = u[:, 0:1] * P[1, :] - u[:, 1:2] * P[0, :]
    X = np.full((n, 4), np.nan)
    ok = np.isfinite(M).all(axis=(1, 2))
    if ok.any():
        X[ok] = np.linalg.svd(M[ok])[2][:, 3, :]
    return X


def linear_eigen_triangulation(u1, P1, u2, P2, max_coordinate_value=1.e16, rows=4):
    X = eigen_homogeneous(u1, P1, u2, P2, rows)
    with np.errstate(divide='ignore', invalid='ignore'):
        x = X[:, 0:3] / X[:, 3:4]                                     # triangulation.py:22
        x_status = np.max(np.abs(x), axis=1) <= max_coordinate_value   # triangulation.py:23 (NaN -> False)
    return x.astype(output_dtype), x_status


def fundamental_from_P(P1, P2):
    """triangulation.py:211-216 : F of the canonical pair, F = [t]x R with P_canon = P2 * inv(P1)."""
    P1_full = np.eye(4); P1_full[0:3, :] = np.asarray(P1, dtype=np.float64)[0:3, :]
    P2_full = np.eye(4); P2_full[0:3, :] = np.asarray(P2, dtype=np.float64)[0:3, :]
    P_canon = P2_full.dot(np.linalg.inv(P1_full))
    return np.cross(P_canon[0:3, 3], P_canon[0:3, 0:3], axisb=0).T


def hartley_sturm_coeffs(a, b, c, d, f1, f2):
    """k0..k6 of g(t) = t((at+b)^2+f2^2(ct+d)^2)^2 - (ad-bc)(1+f1^2 t^2)^2 (at+b)(ct+d)  (H&Z 12.7)."""
    f1s = f1 * f1; f2s = f2 * f2
    # q(t) = (at+b)^2 + f2^2 (ct+d)^2 = q2 t^2 + q1 t + q0
    q2 = a * a + f2s * c * c; q1 = 2 * (a * b + f2s * c * d); q0 = b * b + f2s * d * d
    # t*q^2
    tq = [0 * a, q0 * q0, 2 * q0 * q1, q1 * q1 + 2 * q0 * q2, 2 * q1 * q2, q2 * q2, 0 * a]
    # (ad-bc) (1 + f1^2 t^2)^2 (at+b)(ct+d)
    e = a * d - b * c
    r2 = a * c; r1 = a * d + b * c; r0 = b * d                 # (at+b)(ct+d)
    w0 = 1.0; w2 = 2 * f1s; w4 = f1s * f1s                     # (1+f1^2 t^2)^2
    h = [w0 * r0, w0 * r1, w0 * r2 + w2 * r0, w2 * r1, w2 * r2 + w4 * r0, w4 * r1, w4 * r2]
    return [tq[k] - e * h[k] for k in range(7)]


def _durand_kerner(coeffs, max_iters=100):
    """
    Vectorised restatement of cv::solvePoly as called by correctMatches (maxIters=100):
    leading coefficients with |c| <= DBL_EPSILON are dropped, start at (1+i)^k, Gauss-Seidel
    Durand-Kerner sweeps.  coeffs: (N,7) ascending powers.  Returns (N,6) complex roots, NaN padded.
    """
    n_pts = len(coeffs)
    roots_out = np.full((n_pts, 6), np.nan + 0j, dtype=np.complex128)
    deg = np.full(n_pts, 6)
    for n in range(6, 1, -1):
        drop = (deg == n) & ~(np.abs(coeffs[:, n]) > DBL_EPSILON)
        deg[drop] = n - 1
    for n in range(1, 7):
        sel = np.nonzero(deg == n)[0]
        if len(sel) == 0:
            continue
        c = coeffs[sel, :n + 1].astype(np.complex128)
        r = np.empty((len(sel), n), dtype=np.complex128)
        p = 1 + 0j
        for i in range(n):
            r[:, i] = p
            p = p * (1 + 1j)
        live = np.ones(len(sel), dtype=bool)
        with np.errstate(all='ignore'):
            for _ in range(max_iters):
                maxdiff = np.zeros(len(sel))
                for i in range(n):
                    pi = r[:, i]
                    num = c[:, n].copy(); den = c[:, n].copy()
                    for j in range(n):
                        num = num * pi + c[:, n - j - 1]
                        if j != i:
                            diff = pi - r[:, j]
                            den = np.where(diff != 0, den * diff, den)
                    step = num / den
                    r[:, i] = np.where(live, pi - step, pi)
                    maxdiff = np.maximum(maxdiff, np.abs(step))
                live &= ~(maxdiff <= 0)
                if not live.any():
                    break
        roots_out[sel, :n] = r
    return roots_out


def correct_matches(F, u1, u2, return_t=False):
    """
    Restatement of cv2.correctMatches(F, u1, u2) (Hartley-Sturm), SURVEY.md Appendix A.10.
    Per-point epipoles come from the 3x3 SVD of the translated F, as in OpenCV.
    """
    F = np.asarray(F, dtype=np.float64)
    u1 = _as_f64(u1); u2 = _as_f64(u2)
    n = len(u1)
    T1i = np.tile(np.eye(3), (n, 1, 1)); T1i[:, 0, 2] = u1[:, 0]; T1i[:, 1, 2] = u1[:, 1]
    T2i = np.tile(np.eye(3), (n, 1, 1)); T2i[:, 0, 2] = u2[:, 0]; T2i[:, 1, 2] = u2[:, 1]
    TFT = np.einsum('nji,jk,nkl->nil', T2i, F, T1i)
    ok = np.isfinite(TFT).all(axis=(1, 2))
    e1 = np.full((n, 3), np.nan); e2 = np.full((n, 3), np.nan)
    if ok.any():
        e1[ok] = np.linalg.svd(TFT[ok])[2][:, 2, :]
        e2[ok] = np.linalg.svd(np.transpose(TFT[ok], (0, 2, 1)))[2][:, 2, :]
    with np.errstate(all='ignore'):
        for e in (e1, e2):
            e /= np.sqrt(e[:, 0:1] ** 2 + e[:, 1:2] ** 2)
            e[e[:, 2] < 0] *= -1
        R1 = np.zeros((n, 3, 3)); R2 = np.zeros((n, 3, 3))
        for R, e in ((R1, e1), (R2, e2)):
            R[:, 0, 0] = e[:, 0]; R[:, 0, 1] = e[:, 1]; R[:, 1, 0] = -e[:, 1]; R[:, 1, 1] = e[:, 0]; R[:, 2, 2] = 1
        RTFTR = np.einsum('nij,njk,nlk->nil', R2, TFT, R1)
        f1 = e1[:, 2]; f2 = e2[:, 2]
        a = RTFTR[:, 1, 1]; b = RTFTR[:, 1, 2]; c = RTFTR[:, 2, 1]; d = RTFTR[:, 2, 2]
        k = np.stack(hartley_sturm_coeffs(a, b, c, d, f1, f2), axis=1)
        bad = ~np.isfinite(k).all(axis=1)
        k[bad] = 1.0
        roots = _durand_kerner(k)

        def cost(t):
            return t * t / (1 + f1 * f1 * t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f2 * f2 * (c * t + d) ** 2)

        s_val = 1.0 / (f1 * f1) + c * c / (a * a + f2 * f2 * c * c)      # s(t = inf)
        t_min = np.full(n, np.finfo(np.float64).max)
        for ti in range(6):
            t = roots[:, ti].real
            s = cost(t)
            better = s < s_val
            s_val = np.where(better, s, s_val)
            t_min = np.where(better, t, t_min)
        t_min[bad] = np.nan
        t = t_min
        h1 = np.stack([t * t * f1, t, t * t * f1 * f1 + 1], axis=1)
        ct_d = c * t + d; at_b = a * t + b
        h2 = np.stack([f2 * ct_d ** 2, -at_b * ct_d, f2 * f2 * ct_d ** 2 + at_b ** 2], axis=1)
        h1 = h1 / h1[:, 2:3]; h2 = h2 / h2[:, 2:3]
        n1 = np.einsum('nij,nkj,nk->ni', T1i, R1, h1)
        n2 = np.einsum('nij,nkj,nk->ni', T2i, R2, h2)
    if return_t == 'system':      # + the per-point quantities the cost s(t) and the polynomial g(t) are made of
        return n1[:, 0:2], n2[:, 0:2], t_min, k, (a, b, c, d, f1, f2)
    if return_t:
        return n1[:, 0:2], n2[:, 0:2], t_min, k
    return n1[:, 0:2], n2[:, 0:2]


def find_fundamental_8point(u1, u2):
    """Normalised 8-point F (cv2.findFundamentalMat(..., FM_8POINT)), used only by the all-NaN fallback."""
    u1 = _as_f64(u1); u2 = _as_f64(u2)
    def norm(u):
        m = u.mean(axis=0)
        s = np.sqrt(2.0) / np.mean(np.sqrt(((u - m) ** 2).sum(axis=1)))
        T = np.array([[s, 0, -s * m[0]], [0, s, -s * m[1]], [0, 0, 1]])
        return (u - m) * s, T
    p1, T1 = norm(u1); p2, T2 = norm(u2)
    A = np.stack([p2[:, 0] * p1[:, 0], p2[:, 0] * p1[:, 1], p2[:, 0], p2[:, 1] * p1[:, 0], p2[:, 1] * p1[:, 1],
                  p2[:, 1], p1[:, 0], p1[:, 1], np.ones(len(p1))], axis=1)
    w, V = np.linalg.eigh(A.T @ A)
    F0 = V[:, 0].reshape(3, 3)
    U, s, Vt = np.linalg.svd(F0)
    F0 = U @ np.diag([s[0], s[1], 0]) @ Vt
    F = T2.T @ F0 @ T1
    return F / F[2, 2] if abs(F[2, 2]) > np.finfo(float).eps else F


def polynomial_triangulation(u1, P1, u2, P2, rows=4):
    F = fundamental_from_P(P1, P2)
    u1n, u2n = correct_matches(F, u1, u2)
    if np.isnan(u1n).all() or np.isnan(u2n).all():                     # triangulation.py:227-229
        F = find_fundamental_8point(u1, u2)
        u1n, u2n = correct_matches(F, u1, u2)
    return linear_eigen_triangulation(u1n, P1, u2n, P2, rows=rows)      # triangulation.py:232


def rodrigues(rvec):
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = np.linalg.norm(r)
    if th < DBL_EPSILON:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K


def project_points(objp, rvec, tvec, K, dist):
    """cv2.projectPoints restated; dist = (k1,k2,p1,p2[,k3])."""
    X = np.asarray(objp, dtype=np.float64).reshape(-1, 3)
    R = rodrigues(rvec); t = np.asarray(tvec, dtype=np.float64).reshape(3)
    K = np.asarray(K, dtype=np.float64)
    dist = np.zeros(5) if dist is None else np.asarray(dist, dtype=np.float64).reshape(-1)
    dd = np.zeros(5); dd[:min(5, len(dist))] = dist[:5]
    k1, k2, p1, p2, k3 = dd
    Xc = X @ R.T + t
    with np.errstate(all='ignore'):
        x = Xc[:, 0] / Xc[:, 2]; y = Xc[:, 1] / Xc[:, 2]
        r2 = x * x + y * y
        rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
        xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], axis=1)


def reprojection_error(objp, imgp, cameraMatrix, distCoeffs, rvec, tvec):
    """calibration_tools.py:116-124 : (rms over N points, reprojected points (N,1,2))."""
    proj = project_points(objp, rvec, tvec, cameraMatrix, distCoeffs)
    imgp = np.asarray(imgp, dtype=np.float64).reshape(-1, 2)
    return np.sqrt(((proj - imgp) ** 2).sum() / float(len(imgp))), proj.reshape(-1, 1, 2)


def reprojection_error_ext(objp, imgp, cameraMatrix, distCoeffs, rvecs, tvecs):
    """calibration_tools.py:89-113"""
    mean_error = np.zeros(2); square_error = np.zeros(2)
    n_images = len(imgp)
    for i in range(n_images):
        proj = project_points(objp[i], rvecs[i], tvecs[i], cameraMatrix, distCoeffs)
        error = proj - np.asarray(imgp[i], dtype=np.float64).reshape(-1, 2)
        mean_error += np.abs(error).sum(axis=0) / len(imgp[i])
        square_error += (error ** 2).sum(axis=0) / len(imgp[i])
    return np.linalg.norm(mean_error / n_images), np.sqrt(square_error.sum() / n_images)


def undistort_points(src, cameraMatrix, distCoeffs=None, iters=5):
    """
    cv2.undistortPoints(src, K, dist) for the (k1,k2,p1,p2[,k3]) model with the default criteria (5 fixed-point
    iterations, no R / P), restating OpenCV's cvUndistortPointsInternal operation by operation: float64 arithmetic in
    the same evaluation order (x0 = (u-cx)*(1/fx); icdist = 1/(1+((k3 r2+k2) r2+k1) r2); a negative icdist restores the
    undistorted start value and stops).  Output (N,2) in the dtype of src (float32 stays float32), bit-identical to
    cv2 4.13 on the committed fixture.
    """
    src = np.asarray(src)
    out_dtype = np.float32 if src.dtype == np.float32 else np.float64
    p = src.reshape(-1, 2).astype(np.float64)
    K = np.asarray(cameraMatrix, dtype=np.float64)
    ifx, ify, cx, cy = 1.0 / K[0, 0], 1.0 / K[1, 1], K[0, 2], K[1, 2]
    x0 = (p[:, 0] - cx) * ifx; y0 = (p[:, 1] - cy) * ify
    x = x0.copy(); y = y0.copy()
    if distCoeffs is not None:
        dd = np.zeros(5); d_in = np.asarray(distCoeffs, dtype=np.float64).ravel(); dd[:min(5, len(d_in))] = d_in[:5]
        k1, k2, p1, p2, k3 = dd
        done = np.zeros(len(x), dtype=bool)
        with np.errstate(all='ignore'):
            for _ in range(iters):
                r2 = x * x + y * y
                icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
                neg = (icdist < 0) & ~done
                dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
                dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
                xn = np.where(neg, x0, (x0 - dX) * icdist); yn = np.where(neg, y0, (y0 - dY) * icdist)
                x = np.where(done, x, xn); y = np.where(done, y, yn)
                done |= neg
    return np.stack([x, y], axis=1).astype(out_dtype)


SOLVERS = {
    'linear_eigen': linear_eigen_triangulation,
    'linear_LS': linear_LS_triangulation,
    'iterative_LS': iterative_LS_triangulation,
    'polynomial': polynomial_triangulation,
}
