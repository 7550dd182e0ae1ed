"""Device-resident timing of the m-view linear LS kernel: points/s and algorithmic HBM GB/s.
   python tools/sweep_multiview.py [--points N] [--views 2,4,8] [--visible 1.0]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multiple-quadrotor-slam_b200")); sys.path.insert(0, os.path.join(ROOT, "harness"))
import synthetic_rig as rig          # noqa: E402
import triangl_cuda as tc            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=10_000_000)
ap.add_argument("--views", default="2,4,8")
ap.add_argument("--visible", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
n = args.points
base = min(n, 1_000_000)
reps = -(-n // base)
for m in [int(v) for v in args.views.split(",")]:
    us, Ps, X, valid = rig.make_multiview(base, m, 0.8, p_visible=args.visible)
    du = tc.to_device(np.ascontiguousarray(np.tile(us, (1, reps, 1))[:, :n]))
    dv = tc.to_device(np.ascontiguousarray(np.tile(valid.astype(np.uint8), (1, reps))[:, :n])) if args.visible < 1.0 else None
    x = tc.DeviceArray((n, 3), np.float64); st = tc.DeviceArray((n,), np.uint8)
    d0 = tc.deferred_total()
    for _ in range(3):
        tc.multiview_ls(du, Ps, dv, x=x, status=st)
    tc.synchronize()
    deferred = (tc.deferred_total() - d0) / 3.0
    e = [tc.Event() for _ in range(args.iters + 1)]
    for i in range(args.iters):
        e[i].record(); tc.multiview_ls(du, Ps, dv, x=x, status=st)
    e[args.iters].record(); tc.synchronize()
    ms = np.array([e[i].elapsed_ms(e[i + 1]) for i in range(args.iters)])
    bpp = 16 * m * args.visible + (m if dv is not None else 0) + 25      # observations actually read + masks + x + status
    gbs = bpp * n / (np.median(ms) * 1e-3) / 1e9
    print(json.dumps({"solver": "multiview_LS", "views": m, "visible": args.visible, "n": n, "deferred_per_call": deferred, "ms_median": float(np.median(ms)),
                      "pts_per_s": n / (np.median(ms) * 1e-3), "alg_bytes_per_point": bpp, "alg_GBs": gbs,
                      "frac_of_6550": gbs / 6550.4}))
    del du, dv, x, st
