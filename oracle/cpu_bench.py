"""
CPU baseline leg of bench.py (test infrastructure): times the C restatement of the reference's native path
(oracle/c/triangl_oracle.c, OpenMP over points as in triangulation.c:70,109) on all host cores.
kind = "port": the reference's own C extension needs scipy.weave + OpenCV-2 C headers + Python 2 and cannot be built.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "harness")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synthetic_rig as rig              # noqa: E402
from oracle import oracle_c              # noqa: E402

SOLVERS = ["linear_eigen", "linear_LS", "iterative_LS", "polynomial"]


def host_cores():
    """Cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for nproc > 1: the CPU arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def use_all_cores():
    """omp_set_num_threads(all cores of the box) -- explicit, whatever OMP_NUM_THREADS the launcher exported."""
    oracle_c.set_num_threads(host_cores())
    return oracle_c.num_threads()


def default_sample():
    """About 10-30 s of CPU work in total on a typical host (4 solvers)."""
    return 2_000_000


def time_four_solvers(sample, rig_name, repeats=1):
    use_all_cores()
    u1, P1, u2, P2, _ = rig.make_correspondences(sample, rig_name, sigma=0.8, seed=rig.RSEED)
    parts = {}
    total = 0.0
    for name in SOLVERS:
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle_c.SOLVERS[name](u1, P1, u2, P2)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        parts[name] = best
        total += best
    return total, parts, oracle_c.num_threads(), "port"
