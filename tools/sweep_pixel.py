"""Fused pixel-input solvers vs (undistort kernel x2 + plain solver), device-resident, CUDA events.
   python tools/sweep_pixel.py [--points N] [--iters K]
Algorithmic bytes per correspondence, FP64: fused 57 B (60 B iterative: 32 B pixels in, 24 B x + status out);
unfused 2 x (16 in + 16 out) + 57 = 121 B (124 B)."""
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
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--rig", default="rotating")
args = ap.parse_args()
n = args.points
base = min(n, 2_000_000)
K = np.array([[480., 0, 320], [0, 480., 240], [0, 0, 1]]); dist = np.array([-0.28, 0.07, 2e-4, -1e-4, 0.01])
u1b, P1, u2b, P2, _ = rig.make_correspondences(base, args.rig, 0.8)


def to_px(u):       # forward distortion model + K (NumPy, no oracle import in tools)
    x, y = u[:, 0], u[:, 1]
    r2 = x * x + y * y
    k1, k2, p1, p2, k3 = dist
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x); yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([480. * xd + 320., 480. * yd + 240.], 1)


reps = -(-n // base)
d1 = tc.to_device(np.tile(to_px(u1b), (reps, 1))[:n]); d2 = tc.to_device(np.tile(to_px(u2b), (reps, 1))[:n])
n1 = tc.DeviceArray((n, 2), np.float64); n2 = tc.DeviceArray((n, 2), np.float64)
x = tc.DeviceArray((n, 3), np.float64); sb = tc.DeviceArray((n,), np.uint8); si = tc.DeviceArray((n,), np.int32)
intr = tc.Intrinsics(K, dist)
peak = 6550.4


def timed(fn):
    for _ in range(3):
        fn()
    tc.synchronize()
    e = [tc.Event() for _ in range(args.iters + 1)]
    for i in range(args.iters):
        e[i].record(); fn()
    e[args.iters].record(); tc.synchronize()
    return float(np.median([e[i].elapsed_ms(e[i + 1]) for i in range(args.iters)]))


def solver(name, a, b, pixel):
    if name == "linear_LS":
        tc.linear_ls(a, P1, b, P2, x=x, status=sb, pixel=pixel)
    elif name == "iterative_LS":
        tc.iterative_ls(a, P1, b, P2, x=x, status=si, pixel=pixel)
    elif name == "linear_eigen":
        tc.linear_eigen(a, P1, b, P2, x=x, status=sb, pixel=pixel)
    else:
        tc.polynomial(a, P1, b, P2, x=x, status=sb, check_all_nan=False, pixel=pixel)


ms_u = timed(lambda: tc.undistort_points(d1, K, dist, dst=n1))
print(json.dumps({"kernel": "undistort_points", "n": n, "ms": ms_u, "alg_bytes_per_point": 32,
                  "alg_GBs": 32 * n / ms_u / 1e6, "frac_of_measured_peak": 32 * n / ms_u / 1e6 / peak}))
for name in ("linear_LS", "iterative_LS", "linear_eigen", "polynomial"):
    def unfused():
        tc.undistort_points(d1, K, dist, dst=n1); tc.undistort_points(d2, K, dist, dst=n2); solver(name, n1, n2, None)
    ms_f = timed(lambda: solver(name, d1, d2, intr))
    ms_s = timed(unfused)
    bpp = 60 if name == "iterative_LS" else 57
    print(json.dumps({"solver": name + "_px", "n": n, "fused_ms": ms_f, "unfused_ms": ms_s, "speedup": ms_s / ms_f,
                      "fused_pts_per_s": n / ms_f * 1e3, "fused_alg_GBs": bpp * n / ms_f / 1e6,
                      "fused_frac_of_measured_peak": bpp * n / ms_f / 1e6 / peak}))
