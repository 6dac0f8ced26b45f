"""Summarises an `ncu --set full` report (all kernels of one substep) as a markdown table for profiles/.
    python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.md
    python tools/ncu_summary.py <report.ncu-rep> --json <workload> <particles> [profiles/ncu_traffic.json]
The second form also merges the DRAM traffic per launch of k_g2p / k_p2g into the JSON table that bench.py reads for
`roofline.traffic` (so that the number in the bench line always comes from a committed capture, never from a constant)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
json_args = sys.argv[sys.argv.index("--json") + 1:] if "--json" in sys.argv else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
idx = {n: i for i, n in enumerate(h)}


def g(r, name, scale=1.0, fmt="%.1f"):
    if name not in idx or r[idx[name]] in ("", "n/a"):
        return "-"
    try:
        return fmt % (float(r[idx[name]].replace(",", "")) * scale)
    except ValueError:
        return r[idx[name]]


stall = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
print("| kernel | time us | grid x block | regs | warp-instr (M) | issue active % | warps active % | DRAM read MB | DRAM write MB | DRAM active % | L2 hit % | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx["Kernel Name"]].replace("(DeviceData, int)", "").replace("(DeviceData)", "").replace("void ", "")[:40]
    top = sorted(((float(r[idx[n]]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                  for n in stall if r[idx[n]] not in ("", "n/a")), reverse=True)[:3]
    unit_t = rows[1][idx["gpu__time_duration.sum"]]
    t = float(r[idx["gpu__time_duration.sum"]]) * (1e-3 if unit_t in ("nsecond", "ns") else 1.0)
    print("| `%s` | %.1f | %s x %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        name, t, g(r, "launch__grid_size", fmt="%d"), g(r, "launch__block_size", fmt="%d"), g(r, "launch__registers_per_thread", fmt="%d"),
        g(r, "smsp__inst_executed.sum", 1e-6, "%.2f"), g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "dram__bytes_read.sum", 1.0, "%.2f"),
        g(r, "dram__bytes_write.sum", 1.0, "%.2f"), g(r, "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"),
        g(r, "lts__t_sector_hit_rate.pct"), ", ".join("%s %.2f" % (b, a) for a, b in top)))
print()
print("units as reported by ncu: time %s, dram bytes %s" % (rows[1][idx["gpu__time_duration.sum"]], rows[1][idx["dram__bytes_read.sum"]]))

if json_args:
    import json
    import os

    workload, particles = json_args[0], int(json_args[1])
    path = json_args[2] if len(json_args) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    unit = rows[1][idx["dram__bytes_read.sum"]]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    for r in rows[2:]:
        full = r[idx["Kernel Name"]]
        key = "k_g2p" if full.startswith("k_g2p<") or "k_g2p<" in full.split("(")[0] else None
        if key is None and "k_p2g<" in full.split("(")[0]:
            key = "k_p2g"  # one kernel for all blocks (collider side included) since round 2
        if key is None or "cdf" in full.split("(")[0]:
            continue
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale
        unit_t = rows[1][idx["gpu__time_duration.sum"]]
        t_us = float(r[idx["gpu__time_duration.sum"]]) * (1e-3 if unit_t in ("nsecond", "ns") else 1.0)
        table.setdefault(workload, {})[key] = {"particles": particles, "dram_bytes": rd + wr, "dram_read_bytes": rd,
                                                "dram_write_bytes": wr, "ncu_time_us": t_us, "kernel": full.split("(")[0],
                                                "report": os.path.basename(rep)}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    sys.stderr.write("updated %s\n" % path)
