"""Helpers shared by the parity tests: canonicalisation of the sparse grid / sort output and
field comparisons between the CUDA path (through the C ABI) and the oracle.

Canonicalisation follows SURVEY §8c: header-id numbering, intra-block order and hash slots are
nondeterministic in the reference (grid.wgsl:327; sort.wgsl:126,133), so blocks are compared as a
set keyed by BlockVirtualId and particle ids as per-block sets.
"""
import numpy as np


def block_order(blocks):
    vid = blocks["vid"]
    return np.lexsort((vid[:, 2], vid[:, 1], vid[:, 0]))


def canonical_sort(blocks, sorted_ids):
    """-> (vids sorted, counts, list of sorted id arrays per block)"""
    order = block_order(blocks)
    vids = blocks["vid"][order]
    counts = blocks["num_particles"][order]
    sets = []
    for b in order:
        f, c = int(blocks["first_particle"][b]), int(blocks["num_particles"][b])
        sets.append(np.sort(sorted_ids[f : f + c]))
    return vids, counts, sets


def assert_sort_equal(gpu_blocks, gpu_sorted, ora_blocks, ora_sorted):
    gv, gc, gs = canonical_sort(gpu_blocks, gpu_sorted)
    ov, oc, os_ = canonical_sort(ora_blocks, ora_sorted)
    assert gv.shape == ov.shape, "active block count differs: %d vs %d" % (len(gv), len(ov))
    assert np.array_equal(gv, ov), "active block sets differ"
    assert np.array_equal(gc, oc), "per-block particle counts differ"
    for a, b in zip(gs, os_):
        assert np.array_equal(a, b), "per-block particle id sets differ"
    # every particle appears exactly once
    n = len(gpu_sorted)
    assert np.array_equal(np.sort(gpu_sorted), np.arange(n, dtype=gpu_sorted.dtype))


def canonical_nodes(blocks, nodes):
    order = block_order(blocks)
    return blocks["vid"][order], nodes[order]


def field_rel_err(a, b):
    """max |a - b| over the field, relative to the largest magnitude in the oracle field."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-30)
    return float(np.max(np.abs(a - b)) / scale)


PARTICLE_FIELDS = ("position", "velocity", "def_grad", "affine")


def particle_errors(gpu, ora, fields=PARTICLE_FIELDS):
    return {f: field_rel_err(gpu[f], ora[f]) for f in fields}


def assert_particles_close(gpu, ora, tol, fields=PARTICLE_FIELDS, tols=None):
    errs = particle_errors(gpu, ora, fields)
    for f, e in errs.items():
        t = (tols or {}).get(f, tol)
        assert e <= t, "field %s: relative error %.3e > %.1e (all: %s)" % (f, e, t, errs)
    return errs


def assert_grid_close(gpu_blocks, gpu_nodes, ora_blocks, ora_nodes, tol):
    gv, gn = canonical_nodes(gpu_blocks, gpu_nodes)
    ov, on = canonical_nodes(ora_blocks, ora_nodes)
    assert np.array_equal(gv, ov), "active block sets differ"
    e = field_rel_err(gn["momentum_velocity_mass"], on["momentum_velocity_mass"])
    assert e <= tol, "grid momentum/velocity/mass: relative error %.3e > %.1e" % (e, tol)
    assert np.array_equal(gn["cdf_affinities"], on["cdf_affinities"]), "node affinities differ"
    assert np.array_equal(gn["cdf_closest_id"], on["cdf_closest_id"]), "node closest_id differ"
    m = on["cdf_distance"] < 1e9
    if m.any():
        ed = field_rel_err(gn["cdf_distance"][m], on["cdf_distance"][m])
        assert ed <= 1e-5, "node cdf distance: relative error %.3e" % ed
    return e


def affine_abs_bound(parts, dim, cell_width, dt, ulps=4.0):
    """f32 conditioning bound of the APIC affine matrix (particle_update.wgsl:132):
    affine = grad_v * mass - stress * (V0 * inv_d * dt), and the Kirchhoff stress multiplies
    (sigma - 1) and (J - 1) by 2 mu and lambda (linear_elasticity.wgsl:32-35). A relative error of
    `ulps` f32 ulps in the singular values (exp/log in the Drucker-Prager return mapping are only
    specified to a few ulps in WGSL and in CUDA) therefore moves the affine term by at most
        ulps * eps * (2 mu + d lambda) * V0 * inv_d * dt      (absolute, per particle).
    The reference is subject to the same bound against its own exact-arithmetic statement."""
    eps = float(np.finfo(np.float32).eps)
    stiff = 2.0 * np.abs(parts["mu"].astype(np.float64)) + dim * np.abs(parts["lambda"].astype(np.float64))
    return ulps * eps * stiff * parts["init_volume"] * (4.0 / (cell_width * cell_width)) * dt


def assert_affine_close(gpu, ora, dim, cell_width, dt, rel=1e-5):
    err = np.abs(gpu["affine"].astype(np.float64) - ora["affine"]).max(axis=1)
    bound = affine_abs_bound(ora, dim, cell_width, dt) + rel * np.abs(ora["affine"]).max()
    worst = float((err / bound).max())
    assert worst <= 1.0, "affine: error is %.2fx the f32 conditioning bound" % worst
    return worst
