"""Generates the golden fixtures in tests/golden/*.npz from the oracle (python tests/golden/make_golden.py).

There are no golden vectors in the reference and it cannot be run here, so these fixtures pin the
ORACLE (against drift between rounds) and give the GPU box fixed expected outputs; they do not pin the oracle to
the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from wgsparkl_b200 import scenes  # noqa: E402


def _elastic3d():
    s = scenes.elastic_cube_3d(8, y_offset=-6.0)
    s["particles"]["velocity"][:, 1] = -4.0
    return s


def _sand3d():
    return scenes.sand_column_3d(6, 10, 6, y_offset=-5.0)


def _elastic2d():
    s = scenes.elastic_block_2d(16)
    s["particles"]["position"][:, 1] -= 9.9
    return s


def _coupled3d():
    s = scenes.mixed_coupled_3d(8, 8, 8, n_dynamic=2)
    s["bodies"]["translation"][2:, 1] = 10.0
    return s


def _trimesh3d():
    return scenes.elastic_cube_on_trimesh_3d(10)  # mesh colliders: scene carries "rigid_particles"


CASES = {"elastic3d": (_elastic3d, 25), "sand3d": (_sand3d, 25), "elastic2d": (_elastic2d, 25), "coupled3d": (_coupled3d, 15),
         "trimesh3d": (_trimesh3d, 25)}


def canonical_blocks(blocks):
    order = np.lexsort((blocks["vid"][:, 2], blocks["vid"][:, 1], blocks["vid"][:, 0]))
    return blocks["vid"][order], blocks["num_particles"][order]


def run_case(oracle_mod, make_scene, substeps):
    s = make_scene()
    sim = oracle_mod.OracleSim(s["dim"], s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
    if s.get("rigid_particles") is not None:
        sim.set_rigid_particles(*s["rigid_particles"])
    sim.step(substeps)
    p = sim.read_particles()
    blocks, _ = sim.read_grid()
    vids, counts = canonical_blocks(blocks)
    out = {"position": p["position"], "velocity": p["velocity"], "def_grad": p["def_grad"],
           "cdf_affinity": p["cdf_affinity"], "plastic_hardening": p["plastic_hardening"],
           "block_vids": vids, "block_counts": counts}
    if len(s["bodies"]):
        out["body_translation"] = sim.read_body_poses()["translation"]
        out["body_linvel"] = sim.read_body_vels()["linear"]
    sim.close()
    return out


if __name__ == "__main__":
    from oracle import oracle

    only = sys.argv[1:]  # e.g. `python tests/golden/make_golden.py trimesh3d` adds one fixture, leaving the others alone
    for name, (mk, n) in CASES.items():
        if only and name not in only:
            continue
        out = run_case(oracle, mk, n)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
