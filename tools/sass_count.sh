#!/bin/bash
# Static instruction mix and register count of the FP64 hot kernels (no GPU needed).
#   tools/sass_count.sh [extra nvcc flags, e.g. -DPROBE_EVAL=false -DTRGL_EVAL_MINB=3]
set -e
OUT=${SASS_OUT:-/tmp/probe}
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xptxas -v "$@" \
    -cubin -o $OUT.cubin "$(dirname "$0")/exp/sass_probe.cu" 2> $OUT.ptxas
cuobjdump -sass $OUT.cubin > $OUT.sass
python3 - "$OUT" <<'PY'
import re, sys, collections
out = sys.argv[1]
regs = {}
name = None
for ln in open(out + ".ptxas"):
    m = re.search(r"Compiling entry function '(\w+)'", ln)
    if m: name = m.group(1)
    m = re.search(r"Used (\d+) registers", ln)
    if m and name: regs[name] = (int(m.group(1)), prev_spill)
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m: prev_spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
cur = None
mix = collections.defaultdict(collections.Counter)
for ln in open(out + ".sass"):
    m = re.search(r"Function : (\w+)", ln)
    if m: cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1).split(".")[0]
        mix[cur][op] += 1
for k, c in mix.items():
    short = re.sub(r"_ZN4trgl\d+", "", k)[:28]
    fp64 = sum(v for o, v in c.items() if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    total = sum(c.values())
    r = regs.get(k, (0, (0, 0, 0)))
    print("%-28s regs %3d stack %d spill %d/%d | total %4d FP64 %4d (DFMA %d DMUL %d DADD %d DSETP %d) MUFU %d LDS %d STS %d BRA %d SHFL %d" % (
        short, r[0], r[1][0], r[1][1], r[1][2], total, fp64, c["DFMA"], c["DMUL"], c["DADD"], c["DSETP"], c["MUFU"], c["LDS"], c["STS"], c["BRA"], c["SHFL"]))
PY
