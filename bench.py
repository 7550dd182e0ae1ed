#!/usr/bin/env python
"""
bench.py -- throughput of the batched two-view triangulation hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one pass of the hot path over one synthetic batch: the four solvers of the reference's
triangulation.py (linear_eigen, linear_LS, iterative_LS, polynomial) each followed by the fused two-view
reprojection-error / good-point mask, on P correspondences per GPU (default 10 M = BASELINE.json configs[1]).
Every solver call on P points counts P triangulated points, so a step is 4*P points.

  value  : device-timed, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
  e2e    : the same four solver calls through the drop-in Python API (`triangulation.*_triangulation`, i.e. the
           C ABI in host mode) with pinned HOST inputs; H2D and D2H copies are inside the timed region.
  roofline / per_solver : per-kernel CUDA-event durations measured inside the same timed steps.
  cpu_baseline : the CPU oracle timed on this box's host cores on a bounded sample (rank 0, N = 1 only).

`--impl reference` times the reference's CPU path instead (the oracle port: the reference is Python 2 + a
scipy.weave/OpenCV-2 extension that cannot be built), on all host cores, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multiple-quadrotor-slam_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "harness")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synthetic_rig as rig          # noqa: E402

SOLVERS = ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"]
# algorithmic bytes per point, FP64: u1 16 + u2 16 in, x 24 out, status 1 (bool) / 4 (int32)   
ALG_BYTES = {"linear_eigen": 57, "linear_LS": 57, "iterative_LS": 60, "polynomial": 57}
METRIC = "triangulated_points_per_sec"
UNIT = "points/s"


_REAL_STDOUT = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=10_000_000, help="correspondences per GPU")
    ap.add_argument("--rig", default="rotating", choices=list(rig.RIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="points of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true", help="N > 1: all-gather x over NCCL inside the timed step")
    ap.add_argument("--separate-eval", action="store_true",
                    help="run the two-view reprojection / good-mask evaluation as its own pass after every solver "
                         "(trgl_pair_reproj_async) instead of in the solver kernels' epilogue (trgl_set_fused_eval)")
    ap.add_argument("--fuse-ls-eval", action="store_true",
                    help="also fuse the evaluation into linear_LS.  Default: the three FP64-bound solvers carry it in their "
                         "epilogue (it hides behind the solve), the HBM-bound linear_LS is followed by the stand-alone "
                         "pass (fusing it there turns an HBM-bound kernel into an FP64-bound one for an 11 %% gain)")
    ap.add_argument("--p2p-gather", action="store_true",
                    help="N > 1: gather x/status by storing into every peer's buffer from inside the solver kernels "
                         "(CUDA IPC + NVLink, sharding.PeerGather) instead of a separate NCCL all-gather")
    ap.add_argument("--workload", default="solvers", choices=["solvers", "slam", "scene"],
                    help="solvers: the four solvers at --points per GPU (default, BASELINE configs[1]); "
                         "slam: keyframe map-extension latency at SLAM-sized batches (BASELINE configs[4]); "
                         "scene: 8 cameras pairwise (28 pairs), --points correspondences per GPU sharded by range over "
                         "the ranks, optional NCCL gather of x (BASELINE configs[3])")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_table():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm ---------------------------------------------------------------------------------------------------
def cpu_time_all_solvers(sample, rig_name, repeats=1):
    """Seconds the CPU oracle needs for the four solvers on `sample` points using all host cores."""
    from oracle import cpu_bench
    return cpu_bench.time_four_solvers(sample, rig_name, repeats)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import cpu_bench
    sample = args.cpu_sample or cpu_bench.default_sample()
    for _ in range(args.warmup):
        cpu_bench.time_four_solvers(min(sample, 20000), args.rig, 1)
    t = 0.0
    per = {s: 0.0 for s in SOLVERS}
    for _ in range(args.steps):
        dt, parts, cores, kind = cpu_bench.time_four_solvers(sample, args.rig, 1)
        t += dt
        for s in SOLVERS:
            per[s] += parts[s]
    value = 4.0 * sample * args.steps / t
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 2-camera rig (%s), all four solvers FP64; CPU arm runs a bounded sample of %d "
                               "points per step of the %d-point batch" % (args.rig, sample, args.points),
                   "points_per_gpu": args.points, "rig": args.rig},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d points x 4 solvers per step" % sample},
        "per_solver": {s: {"points_per_sec": sample * args.steps / per[s]} for s in SOLVERS},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---- GPU arm ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import triangl_cuda as tc
    import triangulation as tri

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tc.require_device()
    tc.check(tc.lib().trgl_set_device(local_rank))

    n = args.points
    # one seeded base batch, tiled to n (values repeat, which does not change per-point cost; inputs exceed L2)
    base_n = min(n, 2_000_000)
    u1b, P1, u2b, P2, _ = rig.make_correspondences(base_n, args.rig, sigma=0.8, seed=rig.RSEED + rank)
    reps = -(-n // base_n)
    u1 = np.tile(u1b, (reps, 1))[:n]; u2 = np.tile(u2b, (reps, 1))[:n]

    if world > 1:       # camera matrices come from rank 0 (the only input every shard shares)
        import torch
        cams = torch.from_numpy(np.concatenate([P1.ravel(), P2.ravel()])).cuda()
        dist.broadcast(cams, 0)
        c = cams.cpu().numpy()
        P1 = c[:12].reshape(3, 4).copy(); P2 = c[12:].reshape(3, 4).copy()

    d_u1, d_u2 = tc.to_device(u1), tc.to_device(u2)
    # one result buffer per solver: the four kernels are enqueued back to back (no host sync in between, so the
    # per-kernel CUDA events do not include Python launch latency), then the four fused reprojection passes run
    d_x = {s: tc.DeviceArray((n, 3), np.float64) for s in SOLVERS}
    d_st = {s: tc.DeviceArray((n,), np.int32 if s == "iterative_LS" else np.uint8) for s in SOLVERS}
    gather_buf = None
    peer = None
    if world > 1 and args.gather:
        import torch
        # x lives in torch tensors so NCCL can all-gather it; kernels and NCCL share the legacy default stream
        d_x = {s: torch.empty((n, 3), dtype=torch.float64, device="cuda") for s in SOLVERS}
        gather_buf = torch.empty((world * n, 3), dtype=torch.float64, device="cuda")
    elif world > 1 and args.p2p_gather:
        import sharding
        # every rank owns the full-size result of each solver; the solver kernels store into all of them over NVLink
        peer = {s: sharding.PeerGather(world * n, np.float64, np.int32 if s == "iterative_LS" else np.uint8) for s in SOLVERS}
        for s in SOLVERS:
            d_x[s], d_st[s] = peer[s].shard_outputs()

    ev = [tc.Event() for _ in range(9)]
    d_sums = tc.DeviceArray((4, 4), np.float64)      # per-solver reprojection sums, finished on the device

    fused = {s: tc.FusedEval(n, np.float64, 0, np.inf, want_errors=False, want_good=False, sums=d_sums.view(4 * si, (4,)))
             for si, s in enumerate(SOLVERS)
             if not args.separate_eval and (s != "linear_LS" or args.fuse_ls_eval)}

    def device_step(timed):
        k = 0
        sums_total = 0.0
        for name in SOLVERS:
            ev[k].record(); k += 1
            if peer is not None:
                peer[name].arm()
            fe = fused.get(name)                     # evaluation in the solver's epilogue: no second pass over x, u1, u2
            if name == "linear_eigen":
                tc.linear_eigen(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name], evaluate=fe)
            elif name == "linear_LS":
                tc.linear_ls(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name], evaluate=fe)
            elif name == "iterative_LS":
                tc.iterative_ls(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name], evaluate=fe)
            else:
                tc.polynomial(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name], check_all_nan=False, evaluate=fe)
            ev[k].record(); k += 1
        for si, name in enumerate(SOLVERS):
            if name in fused:
                continue
            # asynchronous variant: the grid-level sums are finished inside the kernel, nothing to wait for per solver
            tc.pair_reproj(d_x[name], d_u1, P1, d_u2, P2, d_st[name], 0, np.inf, want_errors=False, want_good=False,
                           sums_device=d_sums.view(4 * si, (4,)))
        if gather_buf is not None:
            for name in SOLVERS:
                dist.all_gather_into_tensor(gather_buf, d_x[name])
        ev[8].record()
        sums_total = float(d_sums.to_host()[:, 0:2].sum())      # the step's result comes back to the host (synchronises)
        if peer is not None:
            dist.barrier()                                        # every rank's stores have landed: gathered arrays valid
        if timed is not None:
            for i, name in enumerate(SOLVERS):
                timed[name].append(ev[2 * i].elapsed_ms(ev[2 * i + 1]))
        return sums_total

    def barrier():
        tc.synchronize()
        if dist is not None:
            dist.barrier()
        tc.synchronize()

    for _ in range(args.warmup):
        device_step(None)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_kernel = {s: [] for s in SOLVERS}
    launches0 = tc.launch_count()
    barrier()
    e0, e1 = tc.Event(), tc.Event()
    e0.record()
    for _ in range(args.steps):
        device_step(per_kernel)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_ms(e1)
    launches = tc.launch_count() - launches0

    # ---- end to end through the drop-in Python API, pinned host inputs, copies inside the timed region ----
    p_u1, p_u2 = tc.pinned_copy(u1), tc.pinned_copy(u2)
    del u1, u2

    def e2e_step():
        acc = 0.0
        for name in SOLVERS:
            x, st = getattr(tri, name + "_triangulation")(p_u1, P1, p_u2, P2)
            acc += float(x[-1, 2]) + float(st[-1])        # touch the result that came back over PCIe
        return acc

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    if dist is not None:
        import torch
        t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        e2e_s = e2e_ms / 1e3
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    traffic = traffic_table()
    value = 4.0 * n * world * args.steps / (elapsed_ms * 1e-3)
    per_solver = {}
    step_ms = elapsed_ms / args.steps
    for s in SOLVERS:
        ms = float(np.mean(per_kernel[s]))
        gbs = ALG_BYTES[s] * n / (ms * 1e-3) / 1e9
        per_solver[s] = {"kernel_ms": ms, "points_per_sec_per_gpu": n / (ms * 1e-3), "alg_bytes_per_point": ALG_BYTES[s],
                         "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "share_of_step": ms / step_ms,
                         "traffic_bytes_per_point": traffic.get(s),
                         # FP64 vector pipe utilisation of the same kernel under ncu (static, from profiles/): the bound
                         # of iterative_LS / linear_eigen / polynomial (north_star: "FP64 pipe utilisation")
                         "fp64_pipe_busy_ncu": (traffic.get("fp64_pipe_busy") or {}).get(s)}
    ls = per_solver["linear_LS"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "synthetic 2-camera rig (%s), %d points per GPU, all four solvers FP64 + two-view "
                               "reprojection error / good mask %s (BASELINE.json configs[1])"
                               % (args.rig, n, "as its own pass after each" if args.separate_eval else
                                  ("in each solver kernel's epilogue" if args.fuse_ls_eval else
                                   "in the epilogue of the three FP64-bound solver kernels, as its own pass after linear_LS")),
                   "points_per_gpu": n, "rig": args.rig, "sharding": "contiguous point ranges, no data-path collective"
                   + (", NCCL all-gather of x" if args.gather else "")
                   + (", x and status gathered by peer stores from inside the solver kernels (CUDA IPC / NVLink)"
                      if args.p2p_gather else ""),
                   "l2": "inputs (%.0f MB) exceed the 126 MB L2, no explicit flush" % (32.0 * n / 1e6)},
        "roofline": {"kernel": "k_linear_ls<f64>", "bound": "hbm", "achieved": ls["hbm_gbs"], "peak": hbm_peak,
                     "unit": "GB/s", "frac": ls["hbm_frac"],
                     "traffic": (traffic.get("linear_LS") * n) if traffic.get("linear_LS") else None,
                     "peak_source": peak_src, "alg_bytes_per_launch": ALG_BYTES["linear_LS"] * n,
                     "note": "HBM-bound solver named by the north-star target (k_linear_ls + its ~4 us follow-up kernel, "
                             "timed together); per_solver lists all four kernels (iterative_LS / linear_eigen / polynomial "
                             "are FP64-pipe bound).  peak is the measured 1:1 COPY bandwidth; this kernel reads 32 and "
                             "writes 25 bytes per point, so frac can reach ~1.03 at 100 M points (DESIGN.md section 6)"},
        "per_solver": per_solver,
        "e2e": {"value": 4.0 * n * world * e2e_steps / e2e_s, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": 4 * 32 * n, "d2h_bytes_per_step": (25 + 25 + 28 + 25) * n,
                "api": "triangulation.*_triangulation(u1, P1, u2, P2) with pinned host u1/u2 (C ABI host mode)"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench
        sample = args.cpu_sample or cpu_bench.default_sample()
        dt, parts, cores, kind = cpu_bench.time_four_solvers(sample, args.rig, 1)
        out["cpu_baseline"] = {"value": 4.0 * sample / dt, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": "%d points x 4 solvers, same rig and seed" % sample,
                               "per_solver_points_per_sec": {s: sample / parts[s] for s in SOLVERS}}
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


# ---- SLAM keyframe map-extension replay: latency at SLAM-sized batches (BASELINE.json configs[4]) ----------------
def run_slam(args):
    """
    One keyframe = the triangulation work of slam2.py:541-600 on B new correspondences: float32 pixel points, 4x4 P from
    (rvec, tvec), output dtype float32 (slam2.py:19):
        undistortPoints x2 -> iterative_LS -> keep status == 1 -> (solvePnP, not on this path) -> iterative_LS again on the
        inliers -> keep status >= 0.
    Timed end to end through the Python API with pageable host arrays (what SLAM holds), wall clock, median of K keyframes:
      dropin : undistort_points x2 + iterative_LS_triangulation x2   (the reference's call sequence, 4 library calls)
      fused  : iterative_LS_triangulation_px x2                      (undistortion in registers, 2 library calls)
      cpu    : cv2.undistortPoints x2 + the C oracle port of triangulation.c, 1 thread (the reference's shipped build
               has OpenMP disabled, triangulation_c/setup.py:12-13)
    """
    import triangl_cuda as tc
    import triangulation as tri
    tc.require_device()
    K = np.array([[525., 0, 319.5], [0, 525., 239.5], [0, 0, 1]])
    dist = np.array([-0.28, 0.07, 2e-4, -1e-4, 0.01])
    out = {"metric": "slam_keyframe_triangulation_latency", "unit": "us per keyframe (median)", "higher_is_better": False,
           "n_gpus": 1, "dtype": "f64 arithmetic, f32 storage", "data": "synthetic",
           "config": {"workload": "SLAM keyframe map-extension replay (BASELINE.json configs[4]): float32 pixels, 4x4 P, "
                                  "output float32, undistort x2 + iterative_LS x2 + status masks per keyframe",
                      "keyframes_per_size": args.steps, "warmup": args.warmup}, "sizes": {}}
    tri.set_triangl_output_dtype(np.float32)
    try:
        for B in (1000, 3000, 10000, 30000, 100000):
            u1, P1, u2, P2, _ = rig.make_correspondences(B, args.rig, sigma=0.8, seed=rig.RSEED + B)

            def to_px(u):
                x, y = u[:, 0], u[:, 1]
                r2 = x * x + y * y
                k1, k2, p1, p2, k3 = dist
                rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
                return np.stack([K[0, 0] * (x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)) + K[0, 2],
                                 K[1, 1] * (y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y) + K[1, 2]], 1).astype(np.float32)
            px1, px2 = to_px(u1), to_px(u2)
            P1f = np.eye(4); P1f[0:3] = P1
            P2f = np.eye(4); P2f[0:3] = P2

            def dropin():
                n1 = tri.undistort_points(px1, K, dist); n2 = tri.undistort_points(px2, K, dist)
                x, st = tri.iterative_LS_triangulation(n1, P1f, n2, P2f)
                inl = np.where(st == 1)[0]
                n1 = n1[inl]; n2 = n2[inl]
                x, st = tri.iterative_LS_triangulation(n1, P1f, n2, P2f)
                return x[np.where(st >= 0)[0]]

            def fused():
                x, st = tri.iterative_LS_triangulation_px(px1, P1f, px2, P2f, K, dist)
                inl = np.where(st == 1)[0]
                x, st = tri.iterative_LS_triangulation_px(px1[inl], P1f, px2[inl], P2f, K, dist)
                return x[np.where(st >= 0)[0]]

            def timed(fn, reps):
                for _ in range(args.warmup):
                    fn()
                ts = []
                for _ in range(reps):
                    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
                return 1e6 * float(np.median(ts)), 1e6 * float(np.percentile(ts, 95))
            l0 = tc.launch_count()
            d_med, d_p95 = timed(dropin, args.steps)
            f_med, f_p95 = timed(fused, args.steps)
            launches = tc.launch_count() - l0
            assert np.array_equal(dropin(), fused())
            row = {"dropin_us": d_med, "dropin_p95_us": d_p95, "fused_us": f_med, "fused_p95_us": f_p95,
                   "fused_points_per_sec": 2 * B / (f_med * 1e-6), "gpu_launches": launches}
            if not args.no_cpu_baseline:
                from oracle import oracle_c
                try:
                    import cv2
                    und = lambda p: cv2.undistortPoints(p.reshape(-1, 1, 2), K, dist).reshape(-1, 2)      # noqa: E731
                except ImportError:
                    from oracle import triangulation_oracle as orc
                    und = lambda p: orc.undistort_points(p, K, dist)                                       # noqa: E731
                oracle_c.set_num_threads(1)

                def cpu():
                    n1 = und(px1); n2 = und(px2)
                    x, st = oracle_c.iterative_LS_triangulation(n1, P1f[0:3], n2, P2f[0:3])
                    inl = np.where(st == 1)[0]
                    x, st = oracle_c.iterative_LS_triangulation(n1[inl], P1f[0:3], n2[inl], P2f[0:3])
                    return x[np.where(st >= 0)[0]].astype(np.float32)
                c_med, c_p95 = timed(cpu, max(3, min(args.steps, 2_000_000 // B)))
                row.update({"cpu_us": c_med, "cpu_threads": oracle_c.num_threads(), "speedup_fused_vs_cpu": c_med / f_med})
            out["sizes"][str(B)] = row
    finally:
        tri.set_triangl_output_dtype(float)
    emit(out)


# ---- multi-quadrotor scene: 8 cameras pairwise, correspondences sharded over the ranks (BASELINE.json configs[3]) ---
def run_scene(args, rank, world, local_rank):
    """
    8 poses on the trajectory-4/5 circle, all 28 camera pairs, an equal share of the correspondences per pair
    (sharding.pair_segments).  The concatenated correspondence array is sharded by contiguous range: rank r owns
    [r*P, (r+1)*P) of the world*P points and launches one solver call per pair segment that intersects its range, with
    that pair's camera matrices (every rank holds all 8 matrices, broadcast from rank 0).  With --gather every rank
    all-gathers x over NCCL/NVLink inside the timed step.  One step = the four solvers over the rank's range.
    """
    import sharding
    import triangl_cuda as tc
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tc.require_device()
    tc.check(tc.lib().trgl_set_device(local_rank))
    n = args.points
    total = n * world
    cams = sharding.broadcast_cameras(np.stack(rig.circle_cameras(8)), 0, "cuda" if world > 1 else None)
    segs = sharding.pair_segments(8, total)
    lo, hi = sharding.shard_range(total, rank, world)
    mine = sharding.intersect_segments(segs, lo, hi)
    # observations: one seeded cloud per pair, projected into both cameras of the pair (+ 0.8 px noise), tiled
    u1 = np.empty((n, 2)); u2 = np.empty((n, 2))
    for (i, j, off, cnt) in mine:
        base = min(cnt, 500_000)
        rng = np.random.RandomState(rig.RSEED + 100 * i + j)
        X = rig.ball_3D_points(base, 4., rng)
        obs = []
        for P in (cams[i], cams[j]):
            Xc = X.dot(P.T)
            obs.append(Xc[:, 0:2] / Xc[:, 2:3] + rng.normal(0, 0.8 / 480., (base, 2)))
        reps = -(-cnt // base)
        u1[off - lo:off - lo + cnt] = np.tile(obs[0], (reps, 1))[:cnt]
        u2[off - lo:off - lo + cnt] = np.tile(obs[1], (reps, 1))[:cnt]
    d_u1, d_u2 = tc.to_device(u1), tc.to_device(u2)
    del u1, u2
    gather_buf = None
    if world > 1 and args.gather:
        import torch
        d_x = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        gather_buf = torch.empty((world * n, 3), dtype=torch.float64, device="cuda")
    else:
        d_x = tc.DeviceArray((n, 3), np.float64)
    d_sb = tc.DeviceArray((n,), np.uint8); d_si = tc.DeviceArray((n,), np.int32)

    def rows(buf, a, cnt, cols):        # device sub-range handed to the C ABI
        if isinstance(buf, tc.DeviceArray):
            return buf.view(a * max(cols, 1), (cnt, cols) if cols else (cnt,))
        return buf[a:a + cnt]            # torch tensor

    def step():
        for name in SOLVERS:
            for (i, j, off, cnt) in mine:
                a = off - lo
                s1 = rows(d_u1, a, cnt, 2); s2 = rows(d_u2, a, cnt, 2); xs = rows(d_x, a, cnt, 3)
                if name == "linear_eigen":
                    tc.linear_eigen(s1, cams[i], s2, cams[j], x=xs, status=rows(d_sb, a, cnt, 0))
                elif name == "linear_LS":
                    tc.linear_ls(s1, cams[i], s2, cams[j], x=xs, status=rows(d_sb, a, cnt, 0))
                elif name == "iterative_LS":
                    tc.iterative_ls(s1, cams[i], s2, cams[j], x=xs, status=rows(d_si, a, cnt, 0))
                else:
                    tc.polynomial(s1, cams[i], s2, cams[j], x=xs, status=rows(d_sb, a, cnt, 0), check_all_nan=False)
            if gather_buf is not None:
                dist.all_gather_into_tensor(gather_buf, d_x)
        tc.synchronize()

    def barrier():
        tc.synchronize()
        if dist is not None:
            dist.barrier()
        tc.synchronize()
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = tc.launch_count()
    barrier()
    e0, e1 = tc.Event(), tc.Event()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_ms(e1)
    launches = tc.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])
    if rank == 0:
        emit({
            "metric": METRIC, "value": 4.0 * total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "multi-quadrotor scene: 8 cameras pairwise (28 pairs), %d correspondences per GPU "
                                   "(%d total) sharded by contiguous range, all four solvers FP64%s (BASELINE.json configs[3])"
                                   % (n, total, ", NCCL all-gather of x after each solver" if args.gather else ""),
                       "points_per_gpu": n, "pairs": 28, "segments_on_rank0": len(mine),
                       "gather_bytes_per_step": (4 * 24 * total) if args.gather else 0},
            "gpu_launches": launches, "clocks": clocks})
    if dist is not None:
        dist.destroy_process_group()


def main():
    # rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL / torch print an "NCCL version ..." banner
    # at communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to
    # the saved original stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.workload == "slam":
        return run_slam(args)
    if args.workload == "scene":
        return run_scene(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                         int(os.environ.get("LOCAL_RANK", "0")))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
