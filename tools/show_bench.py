"""Compact view of one bench.py JSON line (for gpurun logs)."""
import json
import sys

d = json.loads([ln for ln in open(sys.argv[1]) if ln.strip().startswith("{")][-1])


def short(v, depth=0):
    if isinstance(v, float):
        return float("%.4g" % v)
    if isinstance(v, dict):
        return {k: short(x, depth + 1) for k, x in v.items() if k not in ("note", "api", "workload", "includes", "ncu_fields", "timed", "peak_source", "traffic_source", "l2", "sharding")}
    if isinstance(v, list):
        return [short(x, depth + 1) for x in v]
    return v


print(json.dumps(short(d), indent=1)[:9000])
