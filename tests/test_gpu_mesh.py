"""Mesh colliders: trimesh (3D) / polyline (2D) CPIC through rigid particles - transform of the
sample points, block activation by sample points, p2g_cdf - CUDA path against the oracle, through the C ABI."""
import numpy as np
import pytest

import parity
from wgsparkl_b200 import scenes
from wgsparkl_b200.pipeline import MpmData

pytestmark = pytest.mark.gpu


def _both(scene, pipe, oracle_mod):
    data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    data.set_rigid_particles(*scene["rigid_particles"])
    sim = oracle_mod.OracleSim(scene["dim"], scene["params"], scene["particles"], scene["bodies"], scene["cell_width"],
                               scene["grid_capacity"])
    sim.set_rigid_particles(*scene["rigid_particles"])
    return data, sim


@pytest.mark.parametrize("dim", [3, 2])
def test_block_activation_and_node_cdf(pipe2, pipe3, oracle_mod, dim):
    """After the sort + grid_update_cdf + p2g_cdf stages: the sample points activate exactly the reference's extra
    blocks (sort.wgsl:38-86), and every node carries the same affinity / sign bits, closest collider and distance."""
    scene = scenes.elastic_cube_on_trimesh_3d(12) if dim == 3 else scenes.elastic_block_on_polyline_2d(40)
    pipe = pipe3 if dim == 3 else pipe2
    data, sim = _both(scene, pipe, oracle_mod)
    pipe.sort_only(data)
    pipe.sync()
    for st in range(4):  # update rigid particles, grid sort, grid_update_cdf, p2g_cdf
        sim.stage(st)
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    parity.assert_sort_equal(gb, data.read_sorted_ids(), ob, sim.read_sorted_ids())
    parity.assert_grid_close(gb, gn, ob, on, 0.0)
    coloured = on["cdf_affinities"] != 0
    assert coloured.sum() > 100 and ((on["cdf_affinities"] >> 16) != 0).sum() > 20, "meshes must colour nodes on both sides"
    plain_data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe.sort_only(plain_data)
    assert plain_data.status()[0] < data.status()[0], "sample points must activate blocks of their own"
    plain_data.close()
    data.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_substeps_on_mesh_colliders(pipe2, pipe3, oracle_mod, dim):
    """40 substeps of an elastic body on mesh colliders (3D: tilted slab + a kinematic bar turning into the cube)."""
    scene = scenes.elastic_cube_on_trimesh_3d(12) if dim == 3 else scenes.elastic_block_on_polyline_2d(40)
    pipe = pipe3 if dim == 3 else pipe2
    data, sim = _both(scene, pipe, oracle_mod)
    pipe.queue_step(data, 40)
    pipe.sync()
    sim.step(40)
    g, o = data.read_particles(), sim.read_particles()
    assert (o["cdf_affinity"] != 0).sum() > 40, "the scene must put particles next to the meshes"
    # integer work: bit-exact (k_p2g_cdf / k_transform_rigid round every product and sum on their own, like the oracle)
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    assert parity.field_rel_err(g["position"], o["position"]) <= 1e-5
    assert parity.field_rel_err(g["velocity"], o["velocity"]) <= 2e-3
    assert parity.field_rel_err(g["def_grad"], o["def_grad"]) <= 1e-4
    if dim == 3:
        assert parity.field_rel_err(data.read_body_poses()["translation"], sim.read_body_poses()["translation"]) <= 1e-6
    data.close()
