"""Top SASS instructions of a kernel by stall samples, with the stall reasons and the CUDA source line.
    python tools/ncu_stalls.py <report.ncu-rep> <kernel regex> <cubin> <mangled-substring> [top]"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "--launch-skip", os.environ.get("NCU_SKIP", "0"), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur_line, in_fun = {}, None, False
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
recs = []
for r in rows[hi + 1:]:
    if len(r) <= ix["# Samples"] or not r[ix["Instructions Executed"]].isdigit():
        continue
    addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    recs.append((addr, r))
base = recs[0][0]
tot = sum(int(r[ix["# Samples"]] or 0) for _, r in recs)
by_reason = defaultdict(int)
for _, r in recs:
    for c in stall_cols:
        v = r[ix[c]]
        if v.isdigit():
            by_reason[c] += int(v)
print("samples %d; by reason: %s" % (tot, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])[:8])))
for addr, r in sorted(recs, key=lambda ar: -int(ar[1][ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    reasons = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]].isdigit() and int(r[ix[c]]) > 0), reverse=True)[:3]
    ln = line_of.get(addr - base, ("?", 0))
    print("%5.1f%%  %-60s %s:%d  [%s]" % (100.0 * s / max(tot, 1), r[ix["Source"]][:60], ln[0], ln[1], ", ".join("%s %d" % (b, a) for a, b in reasons)))
