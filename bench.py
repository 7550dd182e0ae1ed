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
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the `gather` sub-record (compute + result gather)")
    ap.add_argument("--gather-steps", type=int, default=5)
    ap.add_argument("--no-100m", action="store_true", help="skip roofline_100M / fp32_study (100 M points per GPU, rank 0)")
    ap.add_argument("--big-points", type=int, default=100_000_000)
    ap.add_argument("--big-iters", type=int, default=7)
    ap.add_argument("--no-scene", action="store_true", help="N > 1: skip the `scene` sub-record (8 cameras, 28 pairs)")
    ap.add_argument("--scene-points", type=int, default=100_000_000, help="correspondences per GPU of the scene sub-record")
    ap.add_argument("--scene-steps", type=int, default=3)
    ap.add_argument("--separate-eval", action="store_true",
                    help="run the two-view reprojection / good-mask evaluation as its own pass after every solver "
                         "(trgl_pair_reproj_async) instead of in the solver kernels' epilogue (trgl_set_fused_eval)")
    ap.add_argument("--separate-ls-eval", action="store_true",
                    help="run the evaluation of linear_LS as its own pass (trgl_pair_reproj_async) after the HBM-bound kernel "
                         "instead of in its epilogue.  Default: all four solvers carry the evaluation in their epilogue "
                         "(linear_LS + epilogue 0.157 ms vs 0.093 + 0.143 ms as two passes per 10 M points); the roofline "
                         "entry always times the plain HBM-bound k_linear_ls by itself")
    ap.add_argument("--workload", default="solvers", choices=["solvers", "slam", "scene"],
                    help="solvers: the four solvers at --points per GPU (default, BASELINE configs[1]); "
                         "slam: keyframe map-extension latency at SLAM-sized batches (BASELINE configs[4]); "
                         "scene: 8 cameras pairwise (28 pairs), --scene-points correspondences per GPU sharded by range "
                         "over the ranks, compute-only / NCCL gather / fused float32-map gather (BASELINE configs[3]; "
                         "also a sub-record of the default run at N > 1)")
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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self, tag):
        """Remember `tag` = now (perf_counter) to classify the samples afterwards."""
        self.marks = getattr(self, "marks", {})
        self.marks[tag] = time.perf_counter()

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
        marks = getattr(self, "marks", {})
        t_a, t_b = marks.get("load_start"), marks.get("load_end")
        under_load = []
        for ts, ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            if t_a is not None and t_b is not None and t_a <= ts <= t_b + 0.05:
                under_load.append(float(parts[0]))
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        use = under_load if under_load else sm
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "samples_under_device_load": len(under_load),
                "load_window": "the timed steps followed by the same step repeated back to back for >= 0.6 s (untimed), so "
                               "that the 50 ms sampler sees the clocks the timed region ran at",
                "reasons": sorted(reasons)}


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
STATUS_DTYPE = {"linear_eigen": np.uint8, "linear_LS": np.uint8, "iterative_LS": np.int32, "polynomial": np.uint8}
NVLINK_PEAK_GBS = 770.0      # measured per direction on this pool's B200s (profiles/r01e_scale_*)


def pin_to_gpu_numa(local_rank):
    """Bind this rank (and with it the first-touch placement of its pinned staging memory) to the NUMA node of its GPU."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0)), "pinned": False}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # 00000000:1b:00.0 -> 0000:1b:00.0
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        info["numa_nodes_on_box"] = nodes
        if cpus and info["numa_node"] is not None and info["numa_node"] >= 0 and nodes > 1:
            os.sched_setaffinity(0, cpus)
            info["pinned"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as exc:                                # noqa: BLE001  (sysfs layout differs in containers)
        info["error"] = repr(exc)[:120]
    return info


def tiled_to_device(tc, base, n):
    """(base_n, k) pinned host array -> (n, k) device array made of repeated async uploads of the base (no n-sized host copy)."""
    d = tc.DeviceArray((n,) + base.shape[1:], base.dtype)
    row = int(np.prod(base.shape[1:])) * base.dtype.itemsize
    off = 0
    while off < n:
        m = min(len(base), n - off)
        tc.check(tc.lib().trgl_memcpy_h2d(d.ptr + off * row, base.ctypes.data, m * row, None))
        off += m
    tc.synchronize()
    return d


def time_launches(tc, fn, warm, iters):
    """Median / min device time of `fn` (CUDA events on the launching stream, back to back)."""
    for _ in range(warm):
        fn()
    tc.synchronize()
    ev = [tc.Event() for _ in range(iters + 1)]
    for i in range(iters):
        ev[i].record(); fn()
    ev[iters].record(); tc.synchronize()
    ms = np.array([ev[i].elapsed_ms(ev[i + 1]) for i in range(iters)])
    return float(np.median(ms)), float(ms.min())


def run_ours(args, rank, world, local_rank):
    numa = pin_to_gpu_numa(local_rank)
    import triangl_cuda as tc
    import triangulation as tri

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tc.require_device()
    tc.check(tc.lib().trgl_set_device(local_rank))

    n = args.points
    # the benchmarked batch: harness/synthetic_rig.bench_batch (one seeded 2 M-point batch tiled to n; inputs exceed L2);
    # tests/test_gpu_parity.py::test_bench_input_parity checks exactly these arrays against the oracle on every point
    u1, P1, u2, P2, base_n = rig.bench_batch(n, args.rig, rank)

    if world > 1:       # camera matrices come from rank 0 (the only input every shard shares)
        cams = torch.from_numpy(np.concatenate([P1.ravel(), P2.ravel()])).cuda()
        dist.broadcast(cams, 0)
        c = cams.cpu().numpy()
        P1 = c[:12].reshape(3, 4).copy(); P2 = c[12:].reshape(3, 4).copy()

    d_u1, d_u2 = tc.to_device(u1), tc.to_device(u2)
    d_x = {s: tc.DeviceArray((n, 3), np.float64) for s in SOLVERS}
    d_st = {s: tc.DeviceArray((n,), STATUS_DTYPE[s]) for s in SOLVERS}
    ev_scratch = [tc.Event() for _ in range(9)]
    d_sums = tc.DeviceArray((4, 4), np.float64)      # per-solver reprojection sums, finished on the device
    thr = (2.0 / 480) ** 2                           # 2 px at f = 480: the harness' reprojection threshold scale

    # the good mask the north star names is WRITTEN in the timed region (1 B/point), by the solver's epilogue or by the pass
    fused = {s: tc.FusedEval(n, np.float64, 0, thr, want_errors=False, want_good=True, sums=d_sums.view(4 * si, (4,)))
             for si, s in enumerate(SOLVERS)
             if not args.separate_eval and (s != "linear_LS" or not args.separate_ls_eval)}
    d_good = {s: tc.DeviceArray((n,), np.bool_) for s in SOLVERS if s not in fused}

    def solve(name, u1_, u2_, x, st, evaluate=None):
        if name == "linear_eigen":
            tc.linear_eigen(u1_, P1, u2_, P2, x=x, status=st, evaluate=evaluate)
        elif name == "linear_LS":
            tc.linear_ls(u1_, P1, u2_, P2, x=x, status=st, evaluate=evaluate)
        elif name == "iterative_LS":
            tc.iterative_ls(u1_, P1, u2_, P2, x=x, status=st, evaluate=evaluate)
        else:
            tc.polynomial(u1_, P1, u2_, P2, x=x, status=st, check_all_nan=False, evaluate=evaluate)

    # Results of a step -- the four solvers' reprojection sums and polynomial's all-NaN words (the np.isnan(u_new).all()
    # test of triangulation.py:227) -- come back to the host through page-locked buffers, asynchronously: step k's are
    # read after step k+1 has been enqueued, so the host never lets the GPU run dry between steps (two buffers in flight).
    h_sums = [tc.pinned_empty((4, 4), np.float64) for _ in range(2)]
    h_flags = [tc.pinned_empty((2,), np.uint32) for _ in range(2)]
    done = [tc.Event(), tc.Event()]
    state = {"k": 0, "pending": [False, False], "checksum": 0.0}

    def collect(slot):
        if state["pending"][slot]:
            done[slot].synchronize()
            assert h_flags[slot][0] != 0 and h_flags[slot][1] != 0, "polynomial: every corrected point is NaN"
            state["checksum"] += float(h_sums[slot][:, 0:2].sum())
            state["pending"][slot] = False

    def device_step(evs, peer=None, gather_buf=None):
        """evs: a set of 9 events of this step's own (timed steps: read after the region), or None."""
        k = 0
        ev = evs if evs is not None else ev_scratch
        slot = state["k"] & 1
        for si, name in enumerate(SOLVERS):
            ev[k].record(); k += 1
            x, st = d_x[name], d_st[name]
            if peer is not None:
                pg = peer[name]
                x, st = pg.shard_outputs()
                pg.arm()
            fe = fused.get(name)                     # evaluation in the solver's epilogue: no second pass over x, u1, u2
            solve(name, d_u1, d_u2, x, st, fe)
            if name == "polynomial":
                tc.polynomial_flags_async(h_flags[slot])
            if fe is None:
                # stand-alone pass (asynchronous variant: the grid-level sums are finished inside the kernel)
                tc.pair_reproj(x, d_u1, P1, d_u2, P2, st, 0, thr, want_errors=False, want_good=d_good[name],
                               sums_device=d_sums.view(4 * si, (4,)))
            ev[k].record(); k += 1
            if gather_buf is not None:
                dist.all_gather_into_tensor(gather_buf[name][0], gather_buf[name][2])
                dist.all_gather_into_tensor(gather_buf[name][1], gather_buf[name][3])
        ev[8].record()
        d_sums.to_host(out=h_sums[slot], sync=False)
        done[slot].record()
        state["pending"][slot] = True
        state["k"] += 1
        collect(slot ^ 1)                                         # the PREVIOUS step's results (this one keeps the GPU busy)
        if peer is not None:
            collect(slot)
            dist.barrier()                                        # every rank's stores have landed: gathered arrays valid

    def barrier():
        tc.synchronize()
        if dist is not None:
            dist.barrier()
        tc.synchronize()

    def timed_steps(steps, warmup, per_kernel=None, **kw):
        for _ in range(warmup):
            device_step(None, **kw)
        ev_sets = [[tc.Event() for _ in range(9)] for _ in range(steps)] if per_kernel is not None else [None] * steps
        barrier()
        collect(0); collect(1)
        e0, e1 = tc.Event(), tc.Event()
        e0.record()
        for i in range(steps):
            device_step(ev_sets[i], **kw)
        collect(0); collect(1)                        # every step's results have been read on the host
        e1.record()
        barrier()
        ms = e0.elapsed_ms(e1)
        if per_kernel is not None:
            for evs in ev_sets:
                for i, name in enumerate(SOLVERS):
                    per_kernel[name].append(evs[2 * i].elapsed_ms(evs[2 * i + 1]))
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_kernel = {s: [] for s in SOLVERS}
    for _ in range(args.warmup):
        device_step(None)
    sampler.mark("load_start")
    launches0 = tc.launch_count()
    elapsed_ms = timed_steps(args.steps, 0, per_kernel)
    launches = tc.launch_count() - launches0
    # the timed region lasts a few tens of milliseconds: keep the same load up for the clock sampler (untimed)
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 0.6:
        device_step(None)
    sampler.mark("load_end")
    # the step times linear_LS together with its evaluation; the HBM-bound kernel alone is timed right here, same buffers
    ls_alone_ms, _ = time_launches(tc, lambda: tc.linear_ls(d_u1, P1, d_u2, P2, x=d_x["linear_LS"], status=d_st["linear_LS"]), 3, 20)

    # ---- result gather (N > 1): the only real exchange step of the path ------------------------------
    gather = None
    if world > 1 and not args.no_gather:
        import sharding
        gather = {"points_per_gpu": n, "steps": args.gather_steps, "nvlink_peak_gbs_per_direction": NVLINK_PEAK_GBS,
                  "compute_only_ms_per_step": elapsed_ms / args.steps, "variants": {}}
        total = world * n

        def record(tag, ms, egress_bpp, ingress_bpp, note):
            step_s = ms / args.gather_steps * 1e-3
            gather["variants"][tag] = {
                "ms_per_step": ms / args.gather_steps, "value": 4.0 * total / step_s, "unit": UNIT,
                "nvlink_egress_gbs_per_rank": egress_bpp * n / step_s / 1e9,
                "nvlink_ingress_gbs_at_receiver": ingress_bpp * n / step_s / 1e9,
                "frac_of_nvlink_peak": max(egress_bpp, ingress_bpp) * n / step_s / 1e9 / NVLINK_PEAK_GBS, "note": note}

        # (1) baseline: NCCL all-gather of x and status after each solver
        tx = {s: torch.empty((n, 3), dtype=torch.float64, device="cuda") for s in SOLVERS}
        ts = {s: torch.empty((n,), dtype=torch.int32 if s == "iterative_LS" else torch.uint8, device="cuda") for s in SOLVERS}
        gbuf = {s: (torch.empty((total, 3), dtype=torch.float64, device="cuda"),
                    torch.empty((total,), dtype=ts[s].dtype, device="cuda"), tx[s], ts[s]) for s in SOLVERS}
        keep_x, keep_st = dict(d_x), dict(d_st)
        d_x.update(tx); d_st.update(ts)
        ms = timed_steps(args.gather_steps, 2, gather_buf=gbuf)
        bpp = 3 * 24 + 25 + 28          # x + status of the four solvers
        record("nccl_allgather_f64", ms, (world - 1) * bpp, (world - 1) * bpp,
               "ncclAllGather of x (24 B) and status (1 | 4 B) after each solver, every rank receives the map")
        d_x.update(keep_x); d_st.update(keep_st)
        del gbuf, tx, ts
        # (2..5) the gather fused into the solver kernels' stores (sharding.PeerGather / trgl_set_result_mirrors[_f32])
        for tag, gdt, root, note in (
                ("fused_allgather_f64", None, None, "peer stores from inside the solver kernels, every rank receives the float64 map"),
                ("fused_allgather_f32map", np.float32, None,
                 "same, the gathered map in float32 (the reference's SLAM map dtype, slam2.py:19); own shard stays float64"),
                ("fused_gather_to_rank0_f64", None, 0, "only rank 0 receives the map: 1/(N-1) of the egress per rank"),
                ("fused_gather_to_rank0_f32map", np.float32, 0, "rank 0 receives the float32 map")):
            peer = {s: sharding.PeerGather(total, np.float64, STATUS_DTYPE[s], gather_dtype=gdt, root=root) for s in SOLVERS}
            ms = timed_steps(args.gather_steps, 2, peer=peer)
            eg = sum(peer[s].egress_bytes_per_point() for s in SOLVERS)
            receivers = 1 if root is not None else world
            ing = sum((world - 1) * (peer[s].xb + peer[s].sb) for s in SOLVERS)
            record(tag, ms, eg, ing, note)
            gather["variants"][tag]["receivers"] = receivers
            barrier()
            for s in SOLVERS:
                peer[s].close()
            del peer
        best = max(gather["variants"].items(), key=lambda kv: kv[1]["value"] if "allgather" in kv[0] else 0)
        gather["best_allgather"] = best[0]

    # ---- driver-run evidence at 100 M points / GPU: linear_LS roofline + the FP32 / FP64 study (BASELINE configs[2]) ------
    hbm_peak, peak_src = peaks()
    big = None
    if not args.no_100m and rank == 0:
        nb = args.big_points
        big = {"points": nb, "timing": "median of %d launches after 3 warm-up launches, CUDA events" % args.big_iters,
               "rig": args.rig, "entries": {}}
        for mode in ("f64", "f32io", "f32"):
            dt = np.float64 if mode == "f64" else np.float32
            b1 = tc.pinned_copy(np.ascontiguousarray(u1[:base_n].astype(dt))); b2 = tc.pinned_copy(np.ascontiguousarray(u2[:base_n].astype(dt)))
            if mode != "f32":
                g1 = tiled_to_device(tc, b1, nb); g2 = tiled_to_device(tc, b2, nb)
                gx = tc.DeviceArray((nb, 3), dt); gsb = tc.DeviceArray((nb,), np.uint8); gsi = tc.DeviceArray((nb,), np.int32)
            cdt = np.float32 if mode == "f32" else np.float64
            isz = np.dtype(dt).itemsize
            for name in (("linear_LS",) if mode == "f32" else ("linear_LS", "iterative_LS", "polynomial", "linear_eigen")):
                kw = dict(out_dtype=dt, compute_dtype=cdt, x=gx)
                if name == "linear_LS":
                    fn = lambda: tc.linear_ls(g1, P1, g2, P2, status=gsb, **kw)                           # noqa: E731
                elif name == "iterative_LS":
                    fn = lambda: tc.iterative_ls(g1, P1, g2, P2, status=gsi, **kw)                        # noqa: E731
                elif name == "linear_eigen":
                    fn = lambda: tc.linear_eigen(g1, P1, g2, P2, status=gsb, **kw)                        # noqa: E731
                else:
                    fn = lambda: tc.polynomial(g1, P1, g2, P2, status=gsb, check_all_nan=False, **kw)     # noqa: E731
                med, mn = time_launches(tc, fn, 3, args.big_iters)
                bpp = 7 * isz + (4 if name == "iterative_LS" else 1)
                gbs = bpp * nb / (med * 1e-3) / 1e9
                big["entries"].setdefault(name, {})[mode] = {
                    "ms": med, "ms_min": mn, "points_per_sec": nb / (med * 1e-3), "alg_bytes_per_point": bpp,
                    "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                    "arithmetic": "float32" if mode == "f32" else "float64 registers", "storage": "float64" if mode == "f64" else "float32"}
            if mode == "f32":
                del g1, g2, gx, gsb, gsi

    # ---- end to end through the drop-in Python API, pinned host inputs, copies inside the timed region ----
    p_u1, p_u2 = tc.pinned_copy(u1), tc.pinned_copy(u2)
    del u1, u2

    def e2e_plain():
        acc = 0.0
        for name in SOLVERS:
            x, st = getattr(tri, name + "_triangulation")(p_u1, P1, p_u2, P2)
            acc += float(x[-1, 2]) + float(st[-1])        # touch the result that came back over PCIe
        return acc

    def e2e_resident():
        # the harness' loop (triangulation_comparison.py:466-469): four solvers on ONE observation set, uploaded once
        h1, h2 = tri.resident(p_u1, p_u2)
        acc = 0.0
        for name in SOLVERS:
            x, st = getattr(tri, name + "_triangulation")(h1, P1, h2, P2)
            acc += float(x[-1, 2]) + float(st[-1])
        return acc

    def time_e2e(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        tc.synchronize()
        mine = time.perf_counter() - t0
        barrier()
        return mine, time.perf_counter() - t0

    e2e_steps = max(2, min(args.steps, 5))
    plain_mine, plain_s = time_e2e(e2e_plain, e2e_steps)
    res_mine, res_s = time_e2e(e2e_resident, e2e_steps)
    pcie = None
    if rank == 0:
        # the ceiling of the host path: plain pinned copies of the step's bytes, one direction at a time
        blk = tc.pinned_empty((1 << 28,), np.uint8); dblk = tc.DeviceArray((1 << 28,), np.uint8)
        pcie = {}
        for tag, call in (("h2d_gbs", lambda: tc.lib().trgl_memcpy_h2d(dblk.ptr, blk.ctypes.data, blk.nbytes, None)),
                          ("d2h_gbs", lambda: tc.lib().trgl_memcpy_d2h(blk.ctypes.data, dblk.ptr, blk.nbytes, None))):
            call(); tc.synchronize()
            t0 = time.perf_counter()
            for _ in range(4):
                call()
            tc.synchronize()
            pcie[tag] = 4 * blk.nbytes / (time.perf_counter() - t0) / 1e9
        del blk, dblk
    clocks = sampler.stop() if rank == 0 else None

    h2d_res, d2h = 32 * n, (25 + 25 + 28 + 25) * n
    per_rank = None
    if dist is not None:
        t = torch.tensor([plain_s * 1e3, res_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plain_s, res_s = float(t[0]) / 1e3, float(t[1]) / 1e3
        mine = torch.tensor([res_mine, plain_mine, float(numa["numa_node"] if numa["numa_node"] is not None else -1),
                             float(numa["cpus"])], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "numa_node": int(v[2]), "cpus": int(v[3]),
                     "resident_h2d_gbs": h2d_res * e2e_steps / float(v[0]) / 1e9, "resident_d2h_gbs": d2h * e2e_steps / float(v[0]) / 1e9,
                     "plain_h2d_gbs": 4 * h2d_res * e2e_steps / float(v[1]) / 1e9, "plain_d2h_gbs": d2h * e2e_steps / float(v[1]) / 1e9}
                    for r, v in enumerate(allr)]
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt[0])

    # ---- multi-quadrotor scene sub-record (N > 1, BASELINE configs[3]) ----------------------------------------------
    scene = None
    if world > 1 and not args.no_scene:
        del d_u1, d_u2, d_x, d_st, p_u1, p_u2, fused, d_good
        tc.pinned_cache_clear()
        scene = scene_record(args, rank, world, tc, dist, torch)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    traffic = traffic_table()
    value = 4.0 * n * world * args.steps / (elapsed_ms * 1e-3)
    per_solver = {}
    step_ms = elapsed_ms / args.steps
    static_src = "static: profiles/traffic.json (ncu --set full of the same kernel; not measured in this run)"
    # FP64 side of the roofline ("measure with an FMA microbenchmark in the same run"): the device's DFMA rate
    # with two and with three distinct register sources per instruction, live; the executed FP64 instructions per point are
    # the ncu counts of the same kernels (static)
    fp64 = None
    try:
        r2, r3 = tc.fp64_fma_rate(2, 8, 4), tc.fp64_fma_rate(3, 8, 4)
        fp64 = {"fma_rate_two_register_sources": r2, "fma_rate_three_register_sources": r3, "unit": "warp FMA instructions/s",
                "tflops_two_register_sources": r2 * 64 / 1e12, "tflops_three_register_sources": r3 * 64 / 1e12,
                "method": "trgl_fp64_fma_rate: 8 independent chains per thread, 32 warps per SM, difference of a 4000- and a "
                          "12000-iteration launch (CUDA events), run right after the timed steps; a DFMA with three distinct "
                          "64-bit register sources issues every 3 cycles instead of 2 (register-file banks), "
                          "csrc/trgl_probe.cuh"}
    except RuntimeError as exc:                      # the probe is diagnostics only
        fp64 = {"error": str(exc)}
    executed = traffic.get("fp64_executed") or {}
    for s in SOLVERS:
        ms = float(np.mean(per_kernel[s]))
        gbs = ALG_BYTES[s] * n / (ms * 1e-3) / 1e9
        ex = executed.get("linear_LS_eval" if (s == "linear_LS" and s in fused_names(args)) else s) or {}
        fp64_rec = None
        if ex and fp64 and "error" not in fp64:
            warp_instr = ex["fp64_instr_per_point"] * n / 32.0
            fp64_rec = {"fp64_instr_per_point_ncu": ex["fp64_instr_per_point"], "instr_per_point_ncu": ex["instr_per_point"],
                        "three_source_share_ncu": ex["three_source_share"],
                        "achieved_warp_instr_per_s": warp_instr / (ms * 1e-3),
                        "frac_of_measured_fma_rate": warp_instr / (ms * 1e-3) / fp64["fma_rate_two_register_sources"],
                        "frac_with_register_bank_cycles": warp_instr * ex["issue_cycles_over_minimum"] / (ms * 1e-3)
                        / fp64["fma_rate_two_register_sources"]}
        per_solver[s] = {"fp64": fp64_rec,"kernel_ms": ms, "points_per_sec_per_gpu": n / (ms * 1e-3), "alg_bytes_per_point": ALG_BYTES[s],
                         "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "share_of_step": ms / step_ms,
                         "includes": "solver kernel + follow-up kernel + evaluation (%s)" %
                                     ("epilogue" if s in fused_names(args) else "stand-alone pass"),
                         "traffic_bytes_per_point_ncu": traffic.get(s),
                         "fp64_pipe_busy_ncu": (traffic.get("fp64_pipe_busy") or {}).get(s), "ncu_fields": static_src}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "synthetic 2-camera rig (%s), %d points per GPU, all four solvers FP64 + two-view "
                               "reprojection error / good mask (written) %s; polynomial's all-NaN flag read once per step "
                               "(BASELINE.json configs[1])"
                               % (args.rig, n, "as its own pass after each" if args.separate_eval else
                                  ("in each solver kernel's epilogue" if not args.separate_ls_eval else
                                   "in the epilogue of the three FP64-bound solver kernels, as its own pass after linear_LS")),
                   "points_per_gpu": n, "rig": args.rig,
                   "sharding": "contiguous point ranges; `value` has no data-path collective, the result gather is in `gather`",
                   "l2": "inputs (%.0f MB) exceed the 126 MB L2, no explicit flush" % (32.0 * n / 1e6)},
        "per_solver": per_solver,
        "fp64_peak": fp64,
        "e2e": {"value": 4.0 * n * world * e2e_steps / res_s, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": h2d_res, "d2h_bytes_per_step": d2h,
                "api": "h1, h2 = triangulation.resident(u1, u2); triangulation.*_triangulation(h1, P1, h2, P2) x 4 -- pinned host "
                       "u1/u2 uploaded ONCE per step by the first solver call, host results (C ABI: trgl_set_input_retention + "
                       "TRGL_MEM_DEVICE_IN)",
                "per_call_upload": {"value": 4.0 * n * world * e2e_steps / plain_s, "h2d_bytes_per_step": 4 * h2d_res,
                                    "d2h_bytes_per_step": d2h,
                                    "api": "triangulation.*_triangulation(u1, P1, u2, P2) x 4 with pinned host u1/u2 (every call "
                                           "uploads them again: the reference-shaped call sequence unchanged)"},
                "pcie_copy_ceiling_rank0": pcie, "numa": numa, "per_rank": per_rank},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    # roofline of the dominant HBM-bound kernel: live CUDA-event time of this run; traffic is the ncu figure of the same kernel
    ls_ms, ls_src = ls_alone_ms, ("k_linear_ls<f64> + its ~4 us follow-up kernel, median of 20 launches back to back right after the "
                                  "timed steps (CUDA events); inside the step linear_LS runs with the evaluation epilogue "
                                  "(per_solver.linear_LS)")
    if ls_ms:
        gbs = ALG_BYTES["linear_LS"] * n / (ls_ms * 1e-3) / 1e9
        out["roofline"] = {"kernel": "k_linear_ls<f64>", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gbs / hbm_peak, "launch_ms": ls_ms, "timed": ls_src,
                           "traffic": (traffic.get("linear_LS") * n) if traffic.get("linear_LS") else None,
                           "traffic_source": static_src, "peak_source": peak_src,
                           "alg_bytes_per_launch": ALG_BYTES["linear_LS"] * n,
                           "note": "peak is the measured 1:1 COPY bandwidth; this kernel reads 32 and writes 25 bytes per point. "
                                   "At 10 M points ~10 % of the written lines are still dirty in the 126 MB L2 when the kernel ends, "
                                   "so roofline_100M is the figure to judge by (DESIGN.md section 6)"}
    if big is not None:
        e = big["entries"]["linear_LS"]["f64"]
        out["roofline_100M"] = {"kernel": "k_linear_ls<f64>", "bound": "hbm", "points": big["points"], "launch_ms": e["ms"],
                                "achieved": e["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": e["hbm_frac"],
                                "alg_bytes_per_launch": e["alg_bytes_per_point"] * big["points"], "peak_source": peak_src}
        out["fp32_study"] = big
    if gather is not None:
        out["gather"] = gather
    if scene is not None:
        out["scene"] = scene
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


def fused_names(args):
    return [s for s in SOLVERS if not args.separate_eval and (s != "linear_LS" or not args.separate_ls_eval)]


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
        # SLAM's own batches are 1-601 points (slam2.py:1080-1082 max_amount_keypoints = 300), BASELINE
        # configs[4] asks for 1 k - 100 k
        for B in (5, 50, 300, 600, 1000, 3000, 10000, 30000, 100000):
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
            tc.set_trace(1); tc.get_trace()
            f_med, f_p95 = timed(fused, args.steps)
            tr = tc.get_trace(); tc.set_trace(0)
            launches = tc.launch_count() - l0
            assert np.array_equal(dropin(), fused())
            row = {"dropin_us": d_med, "dropin_p95_us": d_p95, "fused_us": f_med, "fused_p95_us": f_p95,
                   "fused_points_per_sec": 2 * B / (f_med * 1e-6), "gpu_launches": launches,
                   # mean host microseconds per library call of the fused path, by phase (2 calls per keyframe)
                   "fused_call_breakdown_us": {k: (v / max(tr["calls"], 1.0)) for k, v in tr.items() if k != "calls"}}
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
        cross = [int(b) for b, r in out["sizes"].items() if "cpu_us" in r and r["fused_us"] < r["cpu_us"]]
        out["crossover_batch"] = {"first_size_where_gpu_beats_1_thread_cpu": min(cross) if cross else None,
                                  "sizes_measured": [int(b) for b in out["sizes"]]}
    finally:
        tri.set_triangl_output_dtype(float)
    emit(out)


# ---- multi-quadrotor scene: 8 cameras pairwise, correspondences sharded over the ranks (BASELINE.json configs[3]) ---
def scene_record(args, rank, world, tc, dist, torch):
    """
    8 poses on the trajectory-4/5 circle, all 28 camera pairs, an equal share of the correspondences per pair
    (sharding.pair_segments).  The concatenated correspondence array is sharded by contiguous range: rank r owns
    [r*P, (r+1)*P) of the world*P points and launches one solver call per pair segment that intersects its range, with
    that pair's camera matrices (every rank holds all 8 matrices, broadcast from rank 0).  One step = the four solvers
    over the rank's range.  Timed three ways: compute only; + NCCL all-gather of x (float64) after each solver; + the
    gather fused into the solver kernels' stores with the map in float32 (sharding.PeerGather).
    """
    import sharding
    n = args.scene_points
    total = n * world
    cams = sharding.broadcast_cameras(np.stack(rig.circle_cameras(8)), 0, "cuda" if world > 1 else None)
    segs = sharding.pair_segments(8, total)
    lo, hi = sharding.shard_range(total, rank, world)
    mine = sharding.intersect_segments(segs, lo, hi)
    # observations: one seeded cloud per pair, projected into both cameras of the pair (+ 0.8 px noise), tiled on the device
    d_u1 = tc.DeviceArray((n, 2), np.float64); d_u2 = tc.DeviceArray((n, 2), np.float64)
    for (i, j, off, cnt) in mine:
        base = min(cnt, 500_000)
        rng = np.random.RandomState(rig.RSEED + 100 * i + j)
        X = rig.ball_3D_points(base, 4., rng)
        for P, d in ((cams[i], d_u1), (cams[j], d_u2)):
            Xc = X.dot(P.T)
            obs = tc.pinned_copy(np.ascontiguousarray(Xc[:, 0:2] / Xc[:, 2:3] + rng.normal(0, 0.8 / 480., (base, 2))))
            done = 0
            while done < cnt:
                m = min(base, cnt - done)
                tc.check(tc.lib().trgl_memcpy_h2d(d.ptr + (off - lo + done) * 16, obs.ctypes.data, m * 16, None))
                done += m
            tc.synchronize()
    d_x = tc.DeviceArray((n, 3), np.float64)
    d_sb = tc.DeviceArray((n,), np.uint8); d_si = tc.DeviceArray((n,), np.int32)

    def step(peer=None, gbuf=None):
        for name in SOLVERS:
            st_all = d_si if name == "iterative_LS" else d_sb
            x_all = d_x
            if peer is not None:
                x_all, st_all = peer[name].shard_outputs()
            if gbuf is not None:
                x_all = gbuf[1]
            for (i, j, off, cnt) in mine:
                a = off - lo
                s1 = d_u1.view(2 * a, (cnt, 2)); s2 = d_u2.view(2 * a, (cnt, 2))
                xs = x_all.view(3 * a, (cnt, 3)) if isinstance(x_all, tc.DeviceArray) else x_all[a:a + cnt]
                ss = st_all.view(a, (cnt,))
                if peer is not None:
                    peer[name].arm(a)
                if name == "linear_eigen":
                    tc.linear_eigen(s1, cams[i], s2, cams[j], x=xs, status=ss)
                elif name == "linear_LS":
                    tc.linear_ls(s1, cams[i], s2, cams[j], x=xs, status=ss)
                elif name == "iterative_LS":
                    tc.iterative_ls(s1, cams[i], s2, cams[j], x=xs, status=ss)
                else:
                    tc.polynomial(s1, cams[i], s2, cams[j], x=xs, status=ss, check_all_nan=False)
            if gbuf is not None:
                dist.all_gather_into_tensor(gbuf[0], gbuf[1])
        tc.synchronize()
        if peer is not None:
            dist.barrier()

    def barrier():
        tc.synchronize()
        if dist is not None:
            dist.barrier()
        tc.synchronize()

    def timed(steps, **kw):
        step(**kw); step(**kw)
        barrier()
        e0, e1 = tc.Event(), tc.Event()
        l0 = tc.launch_count()
        e0.record()
        for _ in range(steps):
            step(**kw)
        e1.record()
        barrier()
        ms = e0.elapsed_ms(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms / steps, tc.launch_count() - l0

    steps = args.scene_steps
    rec = {"workload": "multi-quadrotor scene: 8 cameras pairwise (28 pairs), %d correspondences per GPU (%d total) sharded by "
                       "contiguous range, all four solvers FP64 (BASELINE.json configs[3])" % (n, total),
           "points_per_gpu": n, "total_points": total, "pairs": 28, "segments_on_rank0": len(mine), "steps": steps,
           "variants": {}}
    ms, launches = timed(steps)
    rec["variants"]["compute_only"] = {"ms_per_step": ms, "value": 4.0 * total / (ms * 1e-3), "unit": UNIT,
                                       "gpu_launches_rank0": launches}
    if dist is not None and world > 1:
        gx = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        gall = torch.empty((total, 3), dtype=torch.float64, device="cuda")
        ms, _ = timed(steps, gbuf=(gall, gx))
        rec["variants"]["nccl_allgather_x_f64"] = {
            "ms_per_step": ms, "value": 4.0 * total / (ms * 1e-3), "unit": UNIT, "gather_bytes_per_step": 4 * 24 * total,
            "nvlink_ingress_gbs_per_rank": 4 * 24 * (total - n) / (ms * 1e-3) / 1e9}
        del gx, gall
        peer = {"u8": sharding.PeerGather(total, np.float64, np.uint8, gather_dtype=np.float32),
                "i32": sharding.PeerGather(total, np.float64, np.int32, gather_dtype=np.float32)}
        peer = {s: peer["i32" if s == "iterative_LS" else "u8"] for s in SOLVERS}
        ms, _ = timed(steps, peer=peer)
        bpp = sum(peer[s].xb + peer[s].sb for s in SOLVERS)
        rec["variants"]["fused_allgather_f32map"] = {
            "ms_per_step": ms, "value": 4.0 * total / (ms * 1e-3), "unit": UNIT, "gather_bytes_per_step": bpp * total,
            "nvlink_ingress_gbs_per_rank": bpp * (total - n) / (ms * 1e-3) / 1e9,
            "frac_of_nvlink_peak": bpp * (total - n) / (ms * 1e-3) / 1e9 / NVLINK_PEAK_GBS,
            "note": "x (float32 map rows, slam2.py:19) and status stored into every rank's gathered arrays from inside the solver kernels"}
        barrier()
        for pg in set(peer.values()):
            pg.close()
    return rec


def run_scene(args, rank, world, local_rank):
    import triangl_cuda as tc
    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tc.require_device()
    tc.check(tc.lib().trgl_set_device(local_rank))
    args.scene_points = args.points if args.points != 10_000_000 else args.scene_points
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rec = scene_record(args, rank, world, tc, dist, torch)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        best = rec["variants"]["compute_only"]
        emit({"metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": world, "steps": rec["steps"], "warmup": 2,
              "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f64", "data": "synthetic", "config": {"workload": rec["workload"]}, "scene": rec,
              "gpu_launches": best["gpu_launches_rank0"] * world, "clocks": clocks})
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
