"""Timeline of one SHARDED substep inside the graph replay, per rank (see tools/timeline.py):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/timeline_sharded.py [frames_before]
Workload: bench.py's weak-scaling sand dam (2M particles per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.sharded import ShardedMpm  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
before = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scene = scenes.sand_dam_3d(50 * world, 200, 200, grid_capacity=65536)
stream = torch.cuda.Stream()
sh = ShardedMpm(scene, rank, world, local, stream=stream)
spf = scene["substeps_per_frame"]
for _ in range(before):
    sh.step(spf)
sh.sync()
sh.data.debug_timeline()  # on
sh.step(4)
sh.sync()
sh.data.debug_timeline()
acc = {}
reps = 10
for _ in range(reps):
    dist.barrier()
    sh.step(1)
    sh.sync()
    tl = sh.data.debug_timeline()
    t0 = min(v[0] for v in tl.values() if v)
    for k, v in tl.items():
        if v:
            a = acc.setdefault(k, [0.0, 0.0])
            a[0] += (v[0] - t0) / 1e3
            a[1] += (v[1] - t0) / 1e3
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d: one substep, mean of %d (us relative to the substep's first kernel)" % (rank, reps), flush=True)
        for k, (a, b) in sorted(acc.items(), key=lambda kv: kv[1][0]):
            print("  %-16s %8.1f .. %8.1f  (%.1f us)" % (k, a / reps, b / reps, (b - a) / reps), flush=True)
sh.data.debug_timeline(enable=False)
sh.close()
dist.destroy_process_group()
