"""Diagnostic: how far apart are (a) the CUDA path, (b) the oracle with reference rounding semantics and
(c) the oracle in exact-sigma mode, on the stiff-sand golden scene, as substeps accumulate?
    python tools/sand_noise.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from golden.make_golden import CASES  # noqa: E402
from oracle import oracle  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

for name in ("sand3d", "coupled3d"):
    mk, _ = CASES[name]
    for n in (1, 2, 5, 10, 25):
        s = mk()
        res = {}
        for mode in (0, 1):
            oracle.lib().oracle_set_exact_sigma(mode)
            sim = oracle.OracleSim(s["dim"], s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
            sim.step(n)
            res[mode] = sim.read_particles()
            sim.close()
        oracle.lib().oracle_set_exact_sigma(0)
        pipe = MpmPipeline(0, s["dim"])
        data = MpmData(pipe, s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
        pipe.queue_step(data, n)
        pipe.sync()
        g = data.read_particles()
        data.close()
        pipe.close()
        for f in ("position", "velocity", "def_grad"):
            print("%s n=%2d %-9s gpu-ref %.2e  gpu-exact %.2e  ref-exact %.2e" % (
                name, n, f, parity.field_rel_err(g[f], res[0][f]), parity.field_rel_err(g[f], res[1][f]),
                parity.field_rel_err(res[0][f], res[1][f])))
