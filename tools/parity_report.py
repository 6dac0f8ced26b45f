"""Prints GPU-vs-oracle error tables for the parity scenes (diagnostic; run on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from oracle import oracle  # noqa: E402
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402


def affine_bound(parts, dim, h, dt, ulps=4.0):
    eps = np.finfo(np.float32).eps
    lam = np.maximum(np.abs(parts["lambda"]), np.abs(parts["dp_lambda"]) * 0)
    stiff = 2.0 * np.abs(parts["mu"]) + dim * np.abs(parts["lambda"])
    return ulps * eps * stiff * parts["init_volume"] * (4.0 / (h * h)) * dt


def report(name, scene, develop, n=1, tweak=None):
    dim = scene["dim"]
    pipe = MpmPipeline(0, dim)
    sim0 = oracle.OracleSim(dim, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    sim0.step(develop)
    parts = sim0.read_particles()
    bodies = scene["bodies"].copy()
    if len(bodies):
        poses, vels = sim0.read_body_poses(), sim0.read_body_vels()
        bodies["translation"], bodies["rotation"] = poses["translation"], poses["rotation"]
        bodies["linvel"], bodies["angvel"] = vels["linear"], vels["angular"]
    if tweak:
        tweak(parts)
    data = MpmData(pipe, scene["params"], parts, bodies, scene["cell_width"], scene["grid_capacity"])
    sim = oracle.OracleSim(dim, scene["params"], parts, bodies, scene["cell_width"], scene["grid_capacity"])
    pipe.queue_step(data, n)
    pipe.sync()
    sim.step(n)
    g, o = data.read_particles(), sim.read_particles()
    errs = parity.particle_errors(g, o)
    dt = float(scene["params"].dt)
    bound = affine_bound(o, dim, scene["cell_width"], dt)
    aerr = np.abs(g["affine"].astype(np.float64) - o["affine"]).max(axis=1)
    print("%-28s n=%d dev=%d  " % (name, n, develop) + "  ".join("%s=%.2e" % kv for kv in errs.items()))
    print("    max|affine|=%.3e  max abs err=%.3e  max(err/bound)=%.3f  aff mismatches=%d  sd<-.05: %d" % (
        np.abs(o["affine"]).max(), aerr.max(), (aerr / np.maximum(bound, 1e-30)).max(),
        int((g["cdf_affinity"] != o["cdf_affinity"]).sum()), int((o["cdf_signed_distance"] < -0.05 * scene["cell_width"]).sum())))
    for f in ("cdf_normal", "cdf_signed_distance", "cdf_rigid_vel", "plastic_det", "plastic_hardening", "plastic_log_vol_gain", "phase"):
        print("    %-22s %.3e" % (f, parity.field_rel_err(g[f], o[f])), end="")
    print()
    if len(bodies):
        gp, op = data.read_body_poses(), sim.read_body_poses()
        gv, ov = data.read_body_vels(), sim.read_body_vels()
        print("    body trans err %.2e rot err %.2e lin err %.2e (max %.3e) ang err %.2e (max %.3e)" % (
            np.abs(gp["translation"] - op["translation"]).max(), np.abs(gp["rotation"] - op["rotation"]).max(),
            np.abs(gv["linear"] - ov["linear"]).max(), np.abs(ov["linear"]).max(),
            np.abs(gv["angular"] - ov["angular"]).max(), np.abs(ov["angular"]).max()))
    data.close()
    pipe.close()


if __name__ == "__main__":
    report("reference lattice (quirk)", scenes.reference_test_lattice(), 0, n=3)
    s = scenes.elastic_cube_3d(16, y_offset=3.0, ground=False)
    s["particles"]["velocity"][:, 1] = -4.0
    report("elastic cube free", s, 60)
    s = scenes.elastic_cube_3d(16, y_offset=-6.0)
    s["particles"]["velocity"][:, 1] = -4.0

    def push(parts):
        low = (parts["cdf_affinity"] != 0) & (parts["position"][:, 1] < -2.8) & (parts["position"][:, 0] > 0.0)
        parts["position"][low, 1] -= 0.25

    report("elastic cube ground", s, 60, tweak=push)
    report("elastic cube ground 100", scenes.elastic_cube_3d(12, y_offset=-5.0), 0, n=100)
    report("sand column", scenes.sand_column_3d(12, 24, 12, y_offset=-5.0), 60)
    report("sand column 100", scenes.sand_column_3d(12, 24, 12, y_offset=-5.0), 0, n=100)
    s = scenes.elastic_block_2d(40)
    s["particles"]["position"][:, 1] -= 9.9
    report("2d elastic", s, 40)
    s = scenes.mixed_coupled_3d(12, 12, 12, n_dynamic=2)
    s["bodies"]["translation"][2:, 1] = 12.0
    report("mixed coupled", s, 30)
    report("mixed coupled x20", s, 30, n=20)
