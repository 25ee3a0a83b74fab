"""dust self-absorption on the Sphere2D grid: totals of the engine and of the reference over more batches and other seeds"""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import common, test_dust_gpu as t
import skirt_b200 as sk
e = sk.Engine(0)
for grid in ("grid " + common.SYM_GRIDS["sphere2d"], "grid " + common.SYM_GRIDS["cylinder2d"]):
    S, p = t._ref_pan(grid=grid)
    t._engine_for(e, S, p)
    Lv = S.prepare_dust(True)
    ref, gpu = [], []
    for b in range(48):
        S.reset(77100 + 1000 * b); S.run_dust(True, 1.0); ref.append(S.labs_dust().sum())
        Npp = S.packages_per_lambda()
        e.reset_labs_dust(); e.run_dust(1, Lv, Npp, seed=9070 + b); gpu.append(e.fetch_labs_dust().sum())
    ref, gpu = np.array(ref), np.array(gpu)
    z = (gpu.mean() - ref.mean()) / np.hypot(gpu.std(ddof=1) / np.sqrt(len(gpu)), ref.std(ddof=1) / np.sqrt(len(ref)))
    print(grid.split()[1], "gpu", gpu.mean(), "ref", ref.mean(), "rel", gpu.mean() / ref.mean() - 1, "z", z, "stuck", e.stuck_counts(), flush=True)
