"""GPU parity of the "grid sort" pass (WgGrid::queue_sort, src/grid/grid.rs:30-207) against the oracle:
bit-exact after canonicalisation. Calls go through the C ABI."""
import numpy as np
import pytest

import parity
from wgsparkl_b200 import scenes
from wgsparkl_b200.pipeline import MpmData

pytestmark = pytest.mark.gpu


def _run_both(scene, pipe, oracle_mod):
    data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe.sort_only(data)
    pipe.sync()
    gb, gn = data.read_grid()
    gs = data.read_sorted_ids()
    osim = oracle_mod.OracleSim(scene["dim"], scene["params"], scene["particles"], scene["bodies"], scene["cell_width"],
                                scene["grid_capacity"])
    osim.stage(0)
    osim.stage(1)
    osim.stage(2)
    ob, on = osim.read_grid()
    os_ = osim.read_sorted_ids()
    return data, (gb, gn, gs), (ob, on, os_)


def test_reference_lattice_on_round_ties(pipe3, oracle_mod):
    """The reference's own gpu_grid_sort input (grid.rs:355-373): every coordinate is a round() tie."""
    scene = scenes.reference_test_lattice()
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe3, oracle_mod)
    parity.assert_sort_equal(gb, gs, ob, os_)
    assert len(gb) == 27
    data.close()


@pytest.mark.parametrize("n_side", [16, 40])
def test_jittered_cube(pipe3, oracle_mod, n_side):
    scene = scenes.elastic_cube_3d(n_side, y_offset=3.0)
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe3, oracle_mod)
    parity.assert_sort_equal(gb, gs, ob, os_)
    parity.assert_grid_close(gb, gn, ob, on, 0.0)  # momentum zero after reset; node cdf from collide()
    data.close()


def test_negative_coordinates_and_2d(pipe2, oracle_mod):
    scene = scenes.elastic_block_2d(40)
    scene["particles"]["position"][:, 0] -= 3.3  # straddle the origin: negative block ids
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe2, oracle_mod)
    parity.assert_sort_equal(gb, gs, ob, os_)
    parity.assert_grid_close(gb, gn, ob, on, 0.0)
    data.close()


def test_cell_sorted_within_block(pipe3):
    """B200-specific invariant: inside a block the sorted range is ordered by cell-in-block."""
    scene = scenes.elastic_cube_3d(24, y_offset=3.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe3.sort_only(data)
    gb, _ = data.read_grid()
    gs = data.read_sorted_ids()
    pos = scene["particles"]["position"]
    h = np.float32(scene["cell_width"])
    c = (np.rint(pos / h) - 1).astype(np.int64)
    cell = (c[:, 0] & 3) + 4 * (c[:, 1] & 3) + 16 * (c[:, 2] & 3)
    blk = c >> 2
    for b in range(len(gb)):
        f, n = int(gb["first_particle"][b]), int(gb["num_particles"][b])
        ids = gs[f : f + n]
        assert np.all(blk[ids] == gb["vid"][b]), "particle sorted into the wrong block"
        assert np.all(np.diff(cell[ids]) >= 0), "block range is not cell-sorted"
    data.close()


def test_empty_and_single_particle(pipe3, oracle_mod):
    scene = scenes.elastic_cube_3d(4, y_offset=3.0)
    scene["particles"] = scene["particles"][:1].copy()
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe3, oracle_mod)
    parity.assert_sort_equal(gb, gs, ob, os_)
    assert len(gb) == 8
    data.close()
    scene["particles"] = scene["particles"][:0].copy()
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe3.sort_only(data)
    pipe3.queue_step(data, 2)
    pipe3.sync()
    assert data.status() == (0, False)
    data.close()


def test_capacity_overflow_is_reported_not_fatal(pipe3):
    """grid.wgsl:126-128: the reference drops blocks silently; here stepping must survive and report it."""
    scene = scenes.elastic_cube_3d(24, y_offset=3.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], 16)
    pipe3.queue_step(data, 3)
    pipe3.sync()
    nb, overflow = data.status()
    assert overflow and nb <= 16
    out = data.read_particles()
    assert np.all(np.isfinite(out["position"]))
    data.close()


def test_prefix_sum_known_answer(pipe3, oracle_mod):
    """The reference's only numeric test (src/grid/prefix_sum.rs:180-230): LEN 15071, inputs ones / iota /
    random % 10000, expected = WgPrefixSum::eval_cpu."""
    LEN = 15071
    rng = np.random.default_rng(7)
    inputs = [np.ones(LEN, dtype=np.uint32), np.arange(LEN, dtype=np.uint32),
              (rng.integers(0, 2**32, LEN, dtype=np.uint64) % 10_000).astype(np.uint32)]
    for v in inputs + [np.ones(1, dtype=np.uint32), np.ones(2048, dtype=np.uint32), np.ones(2049 * 3, dtype=np.uint32)]:
        got = pipe3.prefix_sum(v)
        exp = np.concatenate([[0], np.cumsum(v.astype(np.uint64))[:-1]]).astype(np.uint32)
        assert np.array_equal(got, exp)
        assert np.array_equal(oracle_mod.prefix_sum(v), exp)


@pytest.mark.parametrize("cell_width", [0.1, 0.37, 1.0 / 3.0, 2.5])
def test_round_ties_and_their_neighbours(pipe3, oracle_mod, cell_width):
    """round(p / h) is evaluated without an IEEE division on the fast path (common.cuh round_div): particles
    exactly on, one ulp below and one ulp above half-integer quotients - for cell widths that are not powers of
    two, at small and large coordinates (inside the +-511-block range of pack_key3) - must land in the same blocks and cells as the oracle's true division."""
    rng = np.random.default_rng(7)
    h = np.float32(cell_width)
    ks = np.concatenate([rng.integers(-40, 40, 600), rng.integers(-1900, 1900, 600)]).astype(np.float32)
    base = ((ks + np.float32(0.5)) * h).astype(np.float32)
    variants = [base, np.nextafter(base, np.float32(np.inf)), np.nextafter(base, np.float32(-np.inf)),
                np.nextafter(np.nextafter(base, np.float32(np.inf)), np.float32(np.inf))]
    x = np.concatenate(variants).astype(np.float32)
    n = len(x)
    scene = scenes.elastic_cube_3d(4, y_offset=3.0)
    parts = np.repeat(scene["particles"][:1], n).copy()
    parts["position"][:, 0] = x
    parts["position"][:, 1] = rng.permutation(x)
    parts["position"][:, 2] = (rng.integers(0, 3, n).astype(np.float32) + np.float32(0.5)) * h
    scene["particles"] = parts
    scene["cell_width"] = float(h)
    scene["grid_capacity"] = 1 << 16
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe3, oracle_mod)
    assert not data.status()[1], "test scene must fit the grid capacity"
    parity.assert_sort_equal(gb, gs, ob, os_)
    data.close()


def test_nearly_full_hash_table(pipe3, oracle_mod):
    """Long linear-probing chains (grid.wgsl:121-164): the capacity is the smallest power of two that still holds
    the active blocks, so most insertions and every neighbour look-up walk over several occupied slots."""
    scene = scenes.elastic_cube_3d(20, y_offset=3.0)  # 10^3 cells: 27..64 touched blocks
    osim = oracle_mod.OracleSim(scene["dim"], scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], 4096)
    for st in (0, 1, 2):
        osim.stage(st)
    nb = osim.num_active_blocks()
    osim.close()
    cap = 1
    while cap < nb:
        cap <<= 1
    scene["grid_capacity"] = cap
    assert nb > 0.6 * cap, (nb, cap)
    data, (gb, gn, gs), (ob, on, os_) = _run_both(scene, pipe3, oracle_mod)
    assert data.status() == (nb, False)
    parity.assert_sort_equal(gb, gs, ob, os_)
    pipe3.queue_step(data, 3)  # and the substep kernels walk the same chains (neighbour table)
    assert np.all(np.isfinite(data.read_positions()))
    assert not data.status()[1]
    data.close()
