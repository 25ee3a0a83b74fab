import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, test_polarization_gpu as T
import skirt_b200 as sk
from oracle import skirtref as sr
ins = T._instruments()
cfg = common.cfg_c1(n=20, packages=2e5, instruments=ins, tau=float(os.environ.get("TAU", "2.0")), threads=1, dustsamples=10)
mu = T.thomson_mueller(1)
S = sr.RefSim(common.ref_spec(cfg), luminosities=[[1.0]], mixes=common.mix_v(), mueller=[mu]).setup()
tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
e = sk.Engine(0); e.set_grid(tables); e.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
e.medium_polarization(*[v[None] for v in mu]); e.sources(cfg["sources"], L, 0.5); e.instruments(ins)
Npp = S.packages_per_lambda()
S.reset(1); S.run_stellar(); e.reset_results(); st = e.run_stellar(Npp, seed=5)
for i in range(3):
    nf = ins[i]["Nxp"] * ins[i]["Nyp"]
    row = []
    for c in (2, 5, 6, 7, 8):
        r = S.full_channel(i, c, nf); g = (e.fetch_frame_channel(i, c), e.fetch_sed_channel(i, c))
        row.append(f"ch{c}: ref frame {r[0].sum():+.4f} sed {r[1].sum():+.4f} | gpu frame {g[0].sum():+.4f} sed {g[1].sum():+.4f}")
    print(ins[i]["name"], *row, sep="\n   ")
