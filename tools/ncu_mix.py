"""Executed-instruction mix and stall samples per opcode of an exported `ncu --page source --csv` file.
   python tools/ncu_mix.py gpurun_out/<tag>/full_<kernel>.source.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci, ce, cs = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex = collections.Counter(); sm = collections.Counter()
tot = tots = 0
for r in rows[2:]:
    try:
        n = float(r[ce]); s = float(r[cs])
    except (ValueError, IndexError):
        continue
    txt = re.sub(r"^@!?U?P\d+\s+", "", r[ci].strip())
    op = txt.split()[0].split(".")[0] if txt else ""
    ex[op] += n; sm[op] += s; tot += n; tots += s
print("warp instructions executed %.0f, samples %.0f" % (tot, tots))
print("executed %:", " ".join("%s %.1f" % (k, 100 * v / tot) for k, v in ex.most_common(24)))
print("samples  %:", " ".join("%s %.1f" % (k, 100 * v / tots) for k, v in sm.most_common(16)))
