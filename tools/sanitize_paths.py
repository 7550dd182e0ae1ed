"""Drive the device-mode paths once each at a size compute-sanitizer gets through in seconds: hot kernels + follow-up kernels
with non-empty deferred lists (forward-motion rig, heavy noise), the fused evaluation epilogue, list overflow, FP32 mode,
pixel inputs, masked multi-view, vector statistics.
   compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_paths.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multiple-quadrotor-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "harness"))
import synthetic_rig as rig          # noqa: E402
import triangl_cuda as tc            # noqa: E402

tc.require_device()
n = 150_003
for rig_name, sigma in (("forward", 20.0), ("rotating", 0.8)):
    u1, P1, u2, P2, _ = rig.make_correspondences(n, rig_name, sigma)
    for dt in (np.float64, np.float32):
        d1, d2 = tc.to_device(u1.astype(dt)), tc.to_device(u2.astype(dt))
        x = tc.DeviceArray((n, 3), dt); sb = tc.DeviceArray((n,), np.uint8); si = tc.DeviceArray((n,), np.int32)
        for with_eval in (False, True):
            for cap in (None, 500):
                old = tc.set_deferred_capacity(cap) if cap else None
                for solver in ("linear_LS", "iterative_LS", "linear_eigen", "polynomial"):
                    fe = tc.FusedEval(n, dt, 0, (2.0 / 480) ** 2, want_errors=True, want_good=True) if with_eval else None
                    kw = dict(out_dtype=dt, compute_dtype=np.float64 if solver != "linear_LS" else dt, x=x, evaluate=fe)
                    if solver == "linear_LS":
                        tc.linear_ls(d1, P1, d2, P2, status=sb, **kw)
                    elif solver == "iterative_LS":
                        tc.iterative_ls(d1, P1, d2, P2, status=si, **kw)
                    elif solver == "linear_eigen":
                        tc.linear_eigen(d1, P1, d2, P2, status=sb, **kw)
                    else:
                        tc.polynomial(d1, P1, d2, P2, status=sb, check_all_nan=True, **kw)
                        tc.polynomial(d1, P1, d2, P2, status=sb, check_all_nan=False, **kw) if fe is None else None
                if cap:
                    tc.set_deferred_capacity(old)
        tc.synchronize()
    print("solvers ok", rig_name, "deferred so far", tc.deferred_total())
us, Ps, X, valid = rig.make_multiview(60_001, 12, 0.8, p_visible=0.6)
for m in (3, 8, 12):
    xm, sm = tc.multiview_ls(tc.to_device(us[:m]), Ps[:m], tc.to_device(valid[:m].astype(np.uint8)))
    xm, sm = tc.multiview_ls(tc.to_device(us[:m]), Ps[:m], None)
tc.synchronize()
print("multiview ok")
K = np.array([[480.0, 0, 320], [0, 480, 240], [0, 0, 1]]); dist = np.array([0.3, 0.01, 1e-3, -1e-3, 0.0])
u1, P1, u2, P2, _ = rig.make_correspondences(40_001, "rotating", 0.8)
px1 = u1 * 480 + np.array([320.0, 240.0]); px2 = u2 * 480 + np.array([320.0, 240.0])
for f in (tc.linear_ls, tc.iterative_ls, tc.linear_eigen, tc.polynomial):
    f(tc.to_device(px1), P1, tc.to_device(px2), P2, pixel=tc.Intrinsics(K, dist))
tc.synchronize()
print("pixel inputs ok; kernels launched", tc.launch_count())
