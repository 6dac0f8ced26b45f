#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck, initcheck) over small scenes that reach every kernel:
#   1. coupled scene: sort, CPIC, both P2G instantiations, plastic G2P, dynamic bodies
#   2. two slabs on one GPU (LocalSlabs): emigrate / immigrate / halo pack + add / dead-tail drop
# Usage (GPU box): tools/sanitize.sh     Output summary: gpurun_out/sanitize.txt
mkdir -p gpurun_out
cat > gpurun_out/_san1.py <<'PY'
import sys
sys.path.insert(0, ".")
from wgsparkl_b200 import scenes
from wgsparkl_b200.pipeline import MpmData, MpmPipeline
s = scenes.mixed_coupled_3d(12, 12, 12, n_dynamic=2)
s["bodies"]["translation"][2:, 1] = 12.0
pipe = MpmPipeline(0, 3)
data = MpmData(pipe, s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
pipe.queue_step(data, 6)
pipe.sync()
p = data.read_particles()
print("ok coupled", len(p), data.status())
data.close(); pipe.close()
PY
cat > gpurun_out/_san2.py <<'PY'
import sys
sys.path.insert(0, ".")
from wgsparkl_b200 import scenes
from wgsparkl_b200.sharded import LocalSlabs
s = scenes.elastic_cube_3d(12, y_offset=-5.0)
s["particles"]["velocity"][:, 0] = 8.0
grp = LocalSlabs(s, 2)
before = grp.live_counts()
grp.step(30)
got = grp.gather_particles()
print("ok slabs", len(got), before, grp.live_counts())
grp.close()
PY
: > gpurun_out/sanitize.txt
for script in _san1 _san2; do
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $script $tool" | tee -a gpurun_out/sanitize.txt
  B200MPM_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python gpurun_out/$script.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|^ok " | head -12 | tee -a gpurun_out/sanitize.txt
done
done
