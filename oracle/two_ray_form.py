"""
NumPy restatement of the TWO-RAY CLOSED FORM the CUDA kernels use for iterative_LS and for the triangulation step of
polynomial (multiple-quadrotor-slam_b200/csrc/trgl_kernels.cuh, "two-ray closed form").  Test infrastructure only: it
lets the CPU suite check, without a GPU, that the closed form IS the reference's re-weighting loop
(Work/python_libs/triangulation_c/triangulation.c:104-161) -- same status vector, same points -- by comparing it with
oracle/triangulation_oracle.py, which follows the reference line by line.

Derivation.  Rows (a0, a1) of view 1 scaled by w1 and (c0, c1) of view 2 scaled by w2 (triangulation.c:30-40,143-146).
Each view's two planes meet in its viewing ray C_k + t n_k, n1 = a0 x a1, C_k the camera centre.  M_k = A_k^T A_k has rank 2
and adj(M_k) = n_k n_k^T, so by Cauchy-Binet the normal-equation solution of the weighted system is

    x(kappa) = (X1 + kappa X2) / (1 + kappa),   kappa = (w2/w1)^2 B/A,
    X1 = C1 + (t1/A) n1,  A = (c0.n1)^2 + (c1.n1)^2,  t1 = -sum_j (c_j.C1 - b_j)(c_j.n1)      (and 1 <-> 2 for X2, B, t2)

and the depths P_k[2,:].[x;1] are the same combination of the depths of X1 and X2.
"""
import numpy as np


def setup(u1, P1, u2, P2):
    """Per-correspondence constants: X1, X2 (n,3), kappa0 = B/A, depths d11 d12 d21 d22 of X1 / X2 in both views, and the
    kappa^2 bound tr^3 / (4 (A + B)) of the unweighted system."""
    u1 = np.asarray(u1, dtype=np.float64); u2 = np.asarray(u2, dtype=np.float64)
    P1 = np.asarray(P1, dtype=np.float64)[0:3]; P2 = np.asarray(P2, dtype=np.float64)[0:3]
    C1 = -np.linalg.solve(P1[:, 0:3], P1[:, 3]); C2 = -np.linalg.solve(P2[:, 0:3], P2[:, 3])
    E2 = P2[:, 0:3] @ C1 + P2[:, 3]                      # P2 [C1; 1]
    E1 = P1[:, 0:3] @ C2 + P1[:, 3]
    a0 = u1[:, 0:1] * P1[2, 0:3] - P1[0, 0:3]; a1 = u1[:, 1:2] * P1[2, 0:3] - P1[1, 0:3]
    c0 = u2[:, 0:1] * P2[2, 0:3] - P2[0, 0:3]; c1 = u2[:, 1:2] * P2[2, 0:3] - P2[1, 0:3]
    n1 = np.cross(a0, a1); n2 = np.cross(c0, c1)
    s0 = (c0 * n1).sum(1); s1 = (c1 * n1).sum(1)
    r0 = (a0 * n2).sum(1); r1 = (a1 * n2).sum(1)
    A = s0 * s0 + s1 * s1; B = r0 * r0 + r1 * r1
    e0 = u2[:, 0] * E2[2] - E2[0]; e1 = u2[:, 1] * E2[2] - E2[1]
    f0 = u1[:, 0] * E1[2] - E1[0]; f1 = u1[:, 1] * E1[2] - E1[1]
    t1 = -(e0 * s0 + e1 * s1); t2 = -(f0 * r0 + f1 * r1)
    with np.errstate(all='ignore'):
        X1 = C1 + (t1 / A)[:, None] * n1; X2 = C2 + (t2 / B)[:, None] * n2
        kap = B / A
        tr = (a0 ** 2).sum(1) + (a1 ** 2).sum(1) + (c0 ** 2).sum(1) + (c1 ** 2).sum(1)
        k2 = tr ** 3 / (4 * (A + B))
        res1 = ((e0 * e0 + e1 * e1) * A - t1 * t1) / ((e0 * e0 + e1 * e1) * A)      # squared residual of ray 1 in view 2's planes
        res2 = ((f0 * f0 + f1 * f1) * B - t2 * t2) / ((f0 * f0 + f1 * f1) * B)
    d = lambda P, X: X @ P[2, 0:3] + P[2, 3]
    return dict(X1=X1, X2=X2, kap=kap, d11=d(P1, X1), d12=d(P1, X2), d21=d(P2, X1), d22=d(P2, X2), k2=k2,
                res=np.maximum(res1, res2))


def linear_ls(u1, P1, u2, P2):
    """kappa = B/A: the unweighted least-squares point (what linear_LS returns; used after the Hartley-Sturm correction)."""
    s = setup(u1, P1, u2, P2)
    return (s["X1"] + s["kap"][:, None] * s["X2"]) / (1 + s["kap"])[:, None], s


def iterative_ls(u1, P1, u2, P2, tolerance=3.e-5, semantics='c'):
    """The reference's loop (triangulation.c:125-159) on the scalar kappa.  Returns x, status, n_solves."""
    s = setup(u1, P1, u2, P2)
    n = len(s["kap"])
    kap = s["kap"].copy(); d1 = np.ones(n); d2 = np.ones(n)
    d1n = np.ones(n); d2n = np.ones(n); kap_used = kap.copy()
    it = np.full(n, 10 if semantics == 'c' else 9)
    done = np.zeros(n, dtype=bool)
    for k in range(10):
        act = ~done
        with np.errstate(all='ignore'):
            inv = 1.0 / (1.0 + kap)
            x1 = (s["d11"] + kap * s["d12"]) * inv; x2 = (s["d21"] + kap * s["d22"]) * inv
            d1n[act] = x1[act]; d2n[act] = x2[act]; kap_used[act] = kap[act]
            conv = (np.abs(x1 - d1) <= tolerance) & (np.abs(x2 - d2) <= tolerance)
            brk = act & (conv | ((x1 == 0) | (x2 == 0) if semantics == 'c' else False))
            it[brk] = k; done |= brk
            cont = ~done
            r = x1 / x2
            kap = np.where(cont, kap * r * r, kap); d1 = np.where(cont, x1, d1); d2 = np.where(cont, x2, d2)
    x = (s["X1"] + kap_used[:, None] * s["X2"]) / (1 + kap_used)[:, None]
    with np.errstate(invalid='ignore'):
        status = ((it < 10) & (d1n > 0) & (d2n > 0)).astype(np.int32) - (d1n <= 0) - 2 * (d2n <= 0)
    return x, status, np.minimum(it + 1, 10), s
