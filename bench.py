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
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import synthetic_rig as rig          # noqa: E402

SOLVERS = ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"]
# algorithmic bytes per point, FP64: u1 16 + u2 16 in, x 24 out, status 1 (bool) / 4 (int32)   
ALG_BYTES = {"linear_eigen": 57, "linear_LS": 57, "iterative_LS": 60, "polynomial": 57}
METRIC = "triangulated_points_per_sec"
UNIT = "points/s"


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
    print(json.dumps(out))


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
    if world > 1 and args.gather:
        import torch
        # x lives in torch tensors so NCCL can all-gather it; kernels and NCCL share the legacy default stream
        d_x = {s: torch.empty((n, 3), dtype=torch.float64, device="cuda") for s in SOLVERS}
        gather_buf = torch.empty((world * n, 3), dtype=torch.float64, device="cuda")

    ev = [tc.Event() for _ in range(9)]

    def device_step(timed):
        k = 0
        sums_total = 0.0
        for name in SOLVERS:
            ev[k].record(); k += 1
            if name == "linear_eigen":
                tc.linear_eigen(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name])
            elif name == "linear_LS":
                tc.linear_ls(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name])
            elif name == "iterative_LS":
                tc.iterative_ls(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name])
            else:
                tc.polynomial(d_u1, P1, d_u2, P2, x=d_x[name], status=d_st[name], check_all_nan=False)
            ev[k].record(); k += 1
        for name in SOLVERS:
            _, _, _, sums = tc.pair_reproj(d_x[name], d_u1, P1, d_u2, P2, d_st[name], 0, np.inf, want_errors=False,
                                           want_good=False)
            sums_total += sums[0] + sums[1]
            if gather_buf is not None:
                dist.all_gather_into_tensor(gather_buf, d_x[name])
        ev[8].record()
        tc.synchronize()
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
                         "traffic_bytes_per_point": traffic.get(s)}
    ls = per_solver["linear_LS"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "synthetic 2-camera rig (%s), %d points per GPU, all four solvers FP64 + fused two-view "
                               "reprojection error / good mask after each (BASELINE.json configs[1])" % (args.rig, n),
                   "points_per_gpu": n, "rig": args.rig, "sharding": "contiguous point ranges, no data-path collective"
                   + (", NCCL all-gather of x" if args.gather else ""),
                   "l2": "inputs (%.0f MB) exceed the 126 MB L2, no explicit flush" % (32.0 * n / 1e6)},
        "roofline": {"kernel": "k_linear_ls<f64>", "bound": "hbm", "achieved": ls["hbm_gbs"], "peak": hbm_peak,
                     "unit": "GB/s", "frac": ls["hbm_frac"],
                     "traffic": (traffic.get("linear_LS") * n) if traffic.get("linear_LS") else None,
                     "peak_source": peak_src, "alg_bytes_per_launch": ALG_BYTES["linear_LS"] * n,
                     "note": "HBM-bound solver named by the north-star target; per_solver lists all four kernels "
                             "(iterative_LS / linear_eigen / polynomial are FP64-pipe bound)"},
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
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
