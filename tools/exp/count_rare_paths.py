"""How often polynomial's rare paths run, and what the follow-up kernel costs against the number of deferred points."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multiple-quadrotor-slam_b200')); sys.path.insert(0, os.path.join(ROOT, 'harness'))
import synthetic_rig as rig          # noqa: E402
import triangl_cuda as tc            # noqa: E402

for name, sig, n in (("forward", 0.8, 10_000_000), ("forward", 0.8, 1_000_000), ("forward", 8.0, 10_000_000), ("rotating", 0.8, 10_000_000)):
    base = min(n, 2_000_000)
    u1, P1, u2, P2, _ = rig.make_correspondences(base, name, sig)
    reps = n // base
    d1 = tc.to_device(np.tile(u1, (reps, 1))); d2 = tc.to_device(np.tile(u2, (reps, 1)))
    x = tc.DeviceArray((n, 3), np.float64); st = tc.DeviceArray((n,), np.uint8)
    tc.rare_path_counters(reset=True); d0 = tc.deferred_total()
    tc.polynomial(d1, P1, d2, P2, x=x, status=st, check_all_nan=False)
    cnt = tc.rare_path_counters(); deferred = tc.deferred_total() - d0
    e0, e1 = tc.Event(), tc.Event()
    ms = []
    for _ in range(10):
        e0.record(); tc.polynomial(d1, P1, d2, P2, x=x, status=st, check_all_nan=False); e1.record()
        ms.append(e0.elapsed_ms(e1))
    print("%s %.1f n %d deferred %d ms median %.4f" % (name, sig, n, deferred, float(np.median(ms))), cnt)
