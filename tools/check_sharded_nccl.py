"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): the NCCL slab-sharded run must reproduce the
single-GPU run. rank 0 runs the unsharded simulation too and compares the gathered particles.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_sharded_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402
from wgsparkl_b200.sharded import ShardedMpm  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ok = True
    for name, scene, n, native, p2p in (
        ("elastic, sliding in +x [nvlink p2p ]", scenes.elastic_cube_3d(16, y_offset=-5.0, nx=16 * world), 80, True, True),
        ("sand + solids + bodies  [nvlink p2p ]", scenes.mixed_coupled_3d(12 * world, 12, 12, n_dynamic=2), 40, True, True),
        ("elastic, sliding in +x [native nccl]", scenes.elastic_cube_3d(16, y_offset=-5.0, nx=16 * world), 80, True, False),
        ("elastic, sliding in +x [torch.dist ]", scenes.elastic_cube_3d(16, y_offset=-5.0, nx=16 * world), 80, False, False),
    ):
        if name.startswith("elastic"):
            scene["particles"]["velocity"][:, 0] = 6.0
            scene["particles"]["velocity"][:, 1] = -2.0
        else:
            scene["bodies"]["translation"][2:, 1] = 12.0
        sh = ShardedMpm(scene, rank, world, local, native=native, p2p=p2p)
        if rank == 0 and p2p:
            print('   p2p active:', getattr(sh, 'p2p', False), flush=True)
        n0 = sh.num_live()
        sh.step(n)
        sh.sync()
        n1 = sh.num_live()
        got = sh.gather_particles()
        counts = [None] * world
        dist.all_gather_object(counts, (n0, n1))
        if rank == 0:
            pipe = MpmPipeline(local, 3)
            data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
            pipe.queue_step(data, n)
            pipe.sync()
            ref = data.read_particles()
            errs = {f: parity.field_rel_err(got[f], ref[f]) for f in ("position", "velocity", "def_grad")}
            sand = "sand" in name
            good = errs["position"] <= (1e-5 if sand else 2e-6) and errs["velocity"] <= (5e-3 if sand else 1e-4) \
                and errs["def_grad"] <= 1e-5 and np.array_equal(got["cdf_affinity"], ref["cdf_affinity"])
            good = good and sum(c[0] for c in counts) == sum(c[1] for c in counts) == len(ref)
            print("%-38s world=%d live before/after %s  errors %s  -> %s" % (
                name, world, counts, {k: "%.2e" % v for k, v in errs.items()}, "PASS" if good else "FAIL"), flush=True)
            ok = ok and good
            data.close()
            pipe.close()
        sh.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
