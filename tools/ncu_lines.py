"""Joins an ncu SASS source page (CSV: per-address instruction counts and stall samples) with nvdisasm line info,
to get per-CUDA-source-line instruction counts. Usage:
    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled-substring> [top]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
import os
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ia, ii, isamp, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
per_addr = []
for r in rows[hi + 1:]:
    if len(r) > ii and r[ii].isdigit():
        per_addr.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), int(r[ii]), int(r[isamp]) if r[isamp].isdigit() else 0, r[isrc]))
base = per_addr[0][0]
# nvdisasm with line info
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fun, line_of = None, {}
cur_line = None
in_fun = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        in_fun = mangled in m.group(1)
        continue
    if not in_fun:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
agg = defaultdict(lambda: [0, 0])
tot = sum(p[1] for p in per_addr)
tots = sum(p[2] for p in per_addr)
for addr, n, s, src in per_addr:
    key = line_of.get(addr - base, ("?", 0))
    agg[key][0] += n
    agg[key][1] += s
print("total warp-instructions %d, samples %d, SASS lines %d" % (tot, tots, len(per_addr)))
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%9d %5.1f%%  samples %5.1f%%  %s:%d" % (n, 100.0 * n / tot, 100.0 * s / max(tots, 1), key[0], key[1]))
