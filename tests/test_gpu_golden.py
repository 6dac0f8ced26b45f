"""CUDA path against the committed golden fixtures (tests/golden/*.npz, generated from the oracle by
tests/golden/make_golden.py): multi-substep runs of small seeded scenes, through the C ABI."""
import os

import numpy as np
import pytest

import parity
from golden.make_golden import CASES, canonical_blocks
from wgsparkl_b200.pipeline import MpmData, MpmPipeline

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_against_golden(name):
    make_scene, substeps = CASES[name]
    s = make_scene()
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    pipe = MpmPipeline(0, s["dim"])
    data = MpmData(pipe, s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
    if s.get("rigid_particles") is not None:
        data.set_rigid_particles(*s["rigid_particles"])
    pipe.queue_step(data, substeps)
    pipe.sync()
    g = data.read_particles()
    blocks, _ = data.read_grid()
    vids, counts = canonical_blocks(blocks)
    # 15-25 substeps from rest. Elastic scenes: 1e-4 on velocities. Scenes with E = 2e9 sand: every ulp of a
    # singular value is 2 mu eps ~ 200 Pa of stress, i.e. ~1e-3 m/s of velocity noise per substep on the
    # lightest-loaded particles; positions still agree to a few ulps.
    sand = name in ("sand3d", "coupled3d")
    assert parity.field_rel_err(g["position"], ref["position"]) <= (1e-5 if sand else 2e-6)
    assert parity.field_rel_err(g["velocity"], ref["velocity"]) <= (5e-3 if sand else 1e-4)
    assert parity.field_rel_err(g["def_grad"], ref["def_grad"]) <= 1e-5
    assert np.array_equal(g["cdf_affinity"], ref["cdf_affinity"])  # integer work: bit-exact, mesh colliders included
    assert parity.field_rel_err(g["plastic_hardening"], ref["plastic_hardening"]) <= 1e-4
    assert np.array_equal(vids, ref["block_vids"]) and np.array_equal(counts, ref["block_counts"])
    if "body_translation" in ref.files:
        assert parity.field_rel_err(data.read_body_poses()["translation"], ref["body_translation"]) <= 1e-5
        assert parity.field_rel_err(data.read_body_vels()["linear"], ref["body_linvel"]) <= 1e-3
    data.close()
    pipe.close()
