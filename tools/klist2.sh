#!/bin/bash
# per-kernel durations of one substep after <frames> frames: tools/klist2.sh <config> <frames>
cfg=$1; fr=$2
skip=$(( fr * 20 * 11 ))
B200MPM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c 22 --csv --log-file gpurun_out/klist2.csv python tools/run_config.py $cfg $(( fr + 1 )) > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/klist2.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:30],[]).append(float(r[vi]))
print("frame $fr", {k:round(sum(v)/len(v)/1000 if max(v)>1000 else sum(v)/len(v),1) for k,v in agg.items()})
PY
