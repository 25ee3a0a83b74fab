"""diagnostic: per-cell absorbed luminosity on a Voronoi grid, engine vs reference runs; prints the worst cells"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import skirt_b200 as sk
from oracle import skirtref as sr
B = int(os.environ.get("B", "16"))
spec = common.spec_grid("voronoi", search=1, packages=1e5, threads=os.cpu_count() or 1, extra=("storeabs 1",))
S = sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), particles=common.voronoi_particles(3000)).setup()
tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
Npp = S.packages_per_lambda()
e = sk.Engine(0); e.set_grid(tables); e.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
e.sources([dict(geometry=1, p=[4000 * common.PC, 350 * common.PC, 0, 0, 0])], L, 0.5)
e.instruments([dict(kind=2, distance=1e7 * common.PC, inclination=float(np.radians(88)))])
ref, gpu = [], []
for b in range(B):
    S.reset(300 + 1000 * b); S.run_stellar(); ref.append(S.labs().ravel().copy())
    e.reset_results(); e.run_stellar(Npp, store_absorption=True, seed=40 + b); gpu.append(e.fetch_labs().ravel())
a, r = np.array(gpu), np.array(ref)
ma, mr = a.mean(0), r.mean(0); sa, sr_ = a.std(0, ddof=1) / np.sqrt(B), r.std(0, ddof=1) / np.sqrt(B)
ok = (sa > 0) & (sr_ > 0) & (sa < 0.15 * ma) & (sr_ < 0.15 * mr)
z = np.where(ok, (ma - mr) / np.sqrt(sa ** 2 + sr_ ** 2 + 1e-300), 0.0)
print("lib", os.environ.get("SKG_LIBRARY", "default"), "bins", ok.sum(), "beyond 3 sigma", int((np.abs(z) >= 3).sum()), "mean z", z[ok].mean(), "max", np.abs(z).max(), "total ratio", a.sum() / r.sum())
nbr = np.diff(tables["nbrStart"]); vol = S.volumes(); walls = np.array([np.any(tables["nbrIds"][tables["nbrStart"][m]:tables["nbrStart"][m + 1]] < 0) for m in range(len(vol))])
for m in np.argsort(-np.abs(z))[:8]:
    print(f"  cell {m}: z {z[m]:+.1f} gpu {ma[m]:.4g}+-{sa[m]:.2g} ref {mr[m]:.4g}+-{sr_[m]:.2g} ratio {ma[m]/mr[m]:.3f} nbrs {nbr[m]} wall {walls[m]} vol^(1/3) {vol[m]**(1/3)/common.PC:.0f} pc rho {medium['rho'][m,0]:.3g}")
