"""One-off GPU probes (round 2): (1) e2e step breakdown, (2) linear_eigen forward-rig offenders, (3) fp32-mode error data."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "multiple-quadrotor-slam_b200"), os.path.join(ROOT, "harness")):
    sys.path.insert(0, p)
import synthetic_rig as rig          # noqa: E402
import triangl_cuda as tc            # noqa: E402
import triangulation as tri          # noqa: E402
from oracle import oracle_c          # noqa: E402

OUT = sys.argv[1]
os.makedirs(OUT, exist_ok=True)
SOLVERS = ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"]

# ---- (1) e2e breakdown -------------------------------------------------------------------------------------------
n = 10_000_000
u1, P1, u2, P2, _ = rig.bench_batch(n, "rotating", 0)
p1, p2 = tc.pinned_copy(u1), tc.pinned_copy(u2)
for rep in range(3):
    t0 = time.perf_counter()
    h1, h2 = tri.resident(p1, p2)
    t = [time.perf_counter()]
    for name in SOLVERS:
        x, st = getattr(tri, name + "_triangulation")(h1, P1, h2, P2)
        t.append(time.perf_counter())
    print("resident rep %d: handle %.2f ms, calls %s ms, total %.2f" % (rep, 1e3 * (t[0] - t0), ["%.2f" % (1e3 * (t[i + 1] - t[i])) for i in range(4)], 1e3 * (t[-1] - t0)), flush=True)
for rep in range(2):
    t = [time.perf_counter()]
    for name in SOLVERS:
        x, st = getattr(tri, name + "_triangulation")(p1, P1, p2, P2)
        t.append(time.perf_counter())
    print("plain rep %d: calls %s ms" % (rep, ["%.2f" % (1e3 * (t[i + 1] - t[i])) for i in range(4)]), flush=True)
# raw D2H of one result-sized buffer, pinned, in 12.5 MB pieces on one stream vs one piece
d = tc.DeviceArray((n, 3), np.float64); h = tc.pinned_empty((n, 3), np.float64)
for pieces in (1, 20):
    tc.synchronize(); t0 = time.perf_counter()
    step = d.nbytes // pieces
    for k in range(pieces):
        tc.check(tc.lib().trgl_memcpy_d2h(h.ctypes.data + k * step, d.ptr + k * step, step, None))
    tc.synchronize()
    print("D2H 240 MB in %d pieces: %.2f ms" % (pieces, 1e3 * (time.perf_counter() - t0)), flush=True)
del d, h, p1, p2, u1, u2, x, st

# ---- (2) linear_eigen forward-rig offenders ---------------------------------------------------------------------
n = 10_000_000
u1, P1, u2, P2, _ = rig.make_correspondences(n, "forward", sigma=0.8, seed=rig.RSEED + 17)
x, st = tc.linear_eigen(tc.to_device(u1), P1, tc.to_device(u2), P2)
tc.synchronize()
x = x.to_host(); st = st.to_host()
xo, so, amp = oracle_c.linear_eigen_triangulation(u1, P1, u2, P2, return_amp=True)
with np.errstate(all="ignore"):
    rel = np.max(np.abs(x - xo), axis=1) / np.max(np.abs(xo), axis=1)
bad = np.where(~(rel <= 1e-9))[0]
print("eigen forward: %d points over 1e-9; amp quantiles of offenders %s" % (len(bad), np.quantile(amp[bad], [0, .5, 1]) if len(bad) else None))
np.savez(os.path.join(OUT, "eigen_offenders.npz"), idx=bad, u1=u1[bad], u2=u2[bad], x=x[bad], xo=xo[bad], amp=amp[bad], rel=rel[bad], P1=P1, P2=P2)
# the same with polynomial
x, st, _ = tc.polynomial(tc.to_device(u1), P1, tc.to_device(u2), P2)
tc.synchronize()
x = x.to_host()
xo, so, amp = oracle_c.polynomial_triangulation(u1, P1, u2, P2, return_amp=True)
with np.errstate(all="ignore"):
    rel = np.max(np.abs(x - xo), axis=1) / np.max(np.abs(xo), axis=1)
bad = np.where(~(rel <= 1e-9) & ~(np.isnan(x).all(1) & np.isnan(xo).all(1)))[0]
print("polynomial forward: %d points over 1e-9" % len(bad))
np.savez(os.path.join(OUT, "poly_offenders.npz"), idx=bad, u1=u1[bad], u2=u2[bad], x=x[bad], xo=xo[bad], amp=amp[bad], rel=rel[bad], P1=P1, P2=P2)
for name in ("linear_LS", "iterative_LS"):
    xg, sg = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
    if name == "linear_LS":
        xo, so = oracle_c.linear_LS_triangulation(u1, P1, u2, P2); mg = np.full(n, 1.0)
    else:
        xo, so, mg = oracle_c.iterative_LS_triangulation(u1, P1, u2, P2, return_margin=True)
    cond = oracle_c.ls_condition(u1, P1, u2, P2)
    rel = np.max(np.abs(xg - xo), axis=1) / np.max(np.abs(xo), axis=1)
    bad = ~(rel <= 1e-9)
    print("%s forward: over 1e-9 %d, of which cond*1.1e-14 >= 1e-9: %d, knife %d, status mismatches %d (knife %d)" % (
        name, bad.sum(), (bad & (cond * 1.1e-14 >= 1e-9)).sum(), (bad & (mg < 1e-9)).sum(), (np.asarray(sg) != so).sum(), ((np.asarray(sg) != so) & (mg < 1e-9)).sum()), flush=True)
del u1, u2, x, xo

# ---- (3) fp32 mode data ----------------------------------------------------------------------------------------
tri.set_triangl_output_dtype(np.float32); tri.set_triangl_compute_dtype(np.float32)
res = {}
for sigma in (0.8,):
    for rname in ("forward", "general", "rotating"):
        u1, P1, u2, P2, _ = rig.make_correspondences(30011, rname, sigma, dtype=np.float32)
        w1, w2 = u1.astype(np.float64), u2.astype(np.float64)
        for name in ("linear_LS", "polynomial"):
            x, st = getattr(tri, name + "_triangulation")(u1, P1, u2, P2)
            if name == "linear_LS":
                xo, so = oracle_c.linear_LS_triangulation(w1, P1, w2, P2); c = oracle_c.ls_condition(w1, P1, w2, P2)
            else:
                # oracle with cv2's dtype convention: the corrected match is rounded to float32 before the triangulation
                from oracle import triangulation_oracle as orc
                F = orc.fundamental_from_P(P1, P2)
                c1, c2 = oracle_c.correct_matches(F, w1, w2)
                c1 = c1.astype(np.float32).astype(np.float64); c2 = c2.astype(np.float32).astype(np.float64)
                xo, so, c = oracle_c.linear_eigen_triangulation(c1, P1, c2, P2, return_amp=True)
            with np.errstate(all="ignore"):
                rel = np.max(np.abs(x - xo), axis=1) / np.max(np.abs(xo), axis=1)
            print("fp32 %s %s: rel quantiles %s max %.3g, over 1e-4: %d; cond/amp quantiles %s; status mism %d" % (
                name, rname, np.nanquantile(rel, [.5, .99, .999]), np.nanmax(rel), (rel > 1e-4).sum(), np.nanquantile(c, [.5, .99]), (np.asarray(st) != so).sum()), flush=True)
            res["%s_%s_rel" % (name, rname)] = rel; res["%s_%s_c" % (name, rname)] = c
np.savez(os.path.join(OUT, "fp32_data.npz"), **res)
