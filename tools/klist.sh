#!/bin/bash
# per-kernel durations of one steady-state substep (ncu, serialised + cold: shares only): tools/klist.sh <config> [variant]
cfg=$1; v=${2:-main}
if [ "$v" = main ]; then lib=""; else lib="wgsparkl_b200/_variants/lib_$v.so"; fi
B200MPM_LIB=$lib B200MPM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 22 --csv --log-file gpurun_out/klist_$v.csv python tools/run_config.py $cfg 1 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/klist_$v.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:32],[]).append(float(r[vi]))
print("$v", {k:round(sum(v)/len(v)/1000 if max(v)>1000 else sum(v)/len(v),1) for k,v in agg.items()})
PY
