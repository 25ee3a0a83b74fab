"""Barycentric trees (OctTreeDustGrid::barycentric, BinTreeDustGrid::directionMethod = Barycenter) built by the reference
itself (oracle/_ref): do the walkers -- the restated CPU oracle, and with --gpu the device walker through the C ABI -- follow
them bit for bit?  The product-side builders do not grow such trees; this checks the table-driven walkers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                                   # noqa: E402
from oracle import oracle_py, skirtref                          # noqa: E402


def main(gpu):
    engine = None
    if gpu:
        import skirt_b200 as sk
        engine = sk.Engine(0)
    ok = True
    for kind, search, lv in (("octtree", 0, 5), ("octtree", 1, 5), ("octtree", 2, 5), ("bintree", 0, 12), ("bintree", 1, 12)):
        spec = common.spec_grid(kind, search=search, maxlevel=lv).replace(" 0 50\n", " 1 50\n")
        S = skirtref.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v()).setup()
        t, med = S.grid_tables(), S.medium()
        box = t["box"].reshape(-1, 6); c0 = t["child0"]
        inner = np.flatnonzero(c0 >= 0)
        mid = 0.5 * (box[inner, :3] + box[inner, 3:]); split = box[c0[inner], 3:]
        offcentre = float(np.mean(np.abs(split - mid).max(axis=1) > 1e-9 * np.abs(box[0]).max()))
        r, k = common.rays(20000, common.C1_BOX, 31)
        ref = S.path_batch(r, k, ell=0, nthreads=os.cpu_count() or 1)
        o = oracle_py.Oracle(t, med)
        line = dict(kind=kind, search=search, cells=int(S.Ncells), offcentre_splits=round(offcentre, 3),
                    oracle=bool(common.paths_bit_identical(o.path_batch(r, k, ell=0), ref)))
        if engine is not None:
            engine.set_grid(t); engine.medium(med["rho"], med["kext"], med["ksca"], med["g"])
            line["gpu"] = bool(common.paths_bit_identical(engine.path_batch(r, k, ell=0), ref))
            line["gpu_whichcell"] = bool(np.array_equal(engine.whichcell(r[:5000]), S.whichcell(r[:5000])))
        ok &= all(v for kk, v in line.items() if kk in ("oracle", "gpu", "gpu_whichcell"))
        print(line, flush=True)
    print("ALL OK" if ok else "MISMATCH")


if __name__ == "__main__":
    main("--gpu" in sys.argv)
