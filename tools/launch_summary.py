"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
n = max(len(v) for v in agg.values())
for k, v in agg.items():
    print("%-40s n=%3d avg=%8.1f us  share=%5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
print("total per substep (cold-cache, serialised): %.1f us" % (tot / n))
