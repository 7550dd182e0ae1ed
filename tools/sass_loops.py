"""Per-loop instruction counts of the kernels in a cuobjdump -sass listing (static view of the dynamic cost: the outermost
loop of a persistent kernel is one tile).   python tools/sass_loops.py /tmp/probe.sass [name-substring ...]"""
import collections
import re
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
path, pats = sys.argv[1], sys.argv[2:]
cur = None
funcs = collections.OrderedDict()
for ln in open(path):
    m = re.search(r"Function : (\w+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for k, ins in funcs.items():
    if pats and not any(p in k for p in pats):
        continue
    print("==", re.sub(r"_ZN4trgl\d+", "", k)[:60], "total", len(ins))
    loops = []
    for a, t in ins:
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2 and int(m2.group(1), 16) < a:
                loops.append((int(m2.group(1), 16), a))
    for lo, hi in sorted(set(loops), key=lambda r: r[0] - r[1])[:int(1e9)]:
        body = [x for x in ins if lo <= x[0] <= hi]
        if len(body) < 100:
            continue
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
        fp = sum(c[o] for o in FP64)
        print("  loop %#06x..%#06x: %4d instr, FP64 %4d, MUFU %2d, CALL %d, other %4d | %s" % (
            lo, hi, len(body), fp, c["MUFU"], c["CALL"], len(body) - fp, " ".join("%s %d" % kv for kv in c.most_common(10))))
