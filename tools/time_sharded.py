"""Timing breakdown of the sharded substep (diagnostic). Single process: phased path vs full graph. Under torchrun:
per-substep time of the native sharded step, and of the bare NCCL exchanges of the same sizes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402
from wgsparkl_b200.sharded import LocalSlabs, ShardedMpm, exchange_with_neighbours  # noqa: E402


def timeit(fn, n, stream):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        fn(n)
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nside = 100
    if world == 1:
        scene = scenes.elastic_cube_3d(nside, y_offset=-5.0)
        stream = torch.cuda.Stream()
        pipe = MpmPipeline(0, 3)
        pipe.set_stream(stream.cuda_stream)
        data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
        pipe.queue_step(data, 60)
        print("full-substep graph      : %.1f us/substep" % timeit(lambda n: pipe.queue_step(data, n), 100, stream))
        grp = LocalSlabs(scene, 1)
        grp.step(60)
        print("phased (begin/end)      : %.1f us/substep" % timeit(lambda n: grp.step(n), 100, grp.stream))
        grp2 = LocalSlabs(scene, 2)
        grp2.step(60)
        print("2 local slabs, one GPU  : %.1f us/substep (both slabs serialised)" % timeit(lambda n: grp2.step(n), 100, grp2.stream))
        return
    import torch.distributed as dist

    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    scene = scenes.elastic_cube_3d(nside, y_offset=-5.0, nx=nside * world)
    for p2p in (False, True):
        sh = ShardedMpm(scene, rank, world, local, p2p=p2p)
        sh.step(60)
        dist.barrier()
        t = timeit(lambda n: sh.step(n), 200, sh.stream)
        if rank == 0:
            print("native sharded substep  : %.1f us/substep (world %d, 1M particles per GPU, p2p=%s)" % (t, world, getattr(sh, "p2p", False)))
        if not p2p:
            sh.close()
    # bare exchanges of the same sizes through torch.distributed (NCCL), 2 per substep
    def ex(n):
        for _ in range(n):
            exchange_with_neighbours(dist, sh.mig_send[0], sh.mig_send[1], sh.mig_recv[0], sh.mig_recv[1], rank, world)
            exchange_with_neighbours(dist, sh.halo_send[0], sh.halo_send[1], sh.halo_recv[0], sh.halo_recv[1], rank, world)
    ex(20)
    dist.barrier()
    t = timeit(ex, 200, sh.stream)
    if rank == 0:
        print("2 bare neighbour exchanges (%d + %d KB): %.1f us" % (sh.mig_send[0].numel() // 1024, sh.halo_send[0].numel() // 1024, t))
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
