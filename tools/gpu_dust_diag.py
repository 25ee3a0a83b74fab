"""diagnostic: per-(cell, wavelength) absorbed dust luminosity of one self-absorption cycle, engine vs reference runs"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import test_dust_gpu as T
import skirt_b200 as sk
B = int(os.environ.get("B", "16"))
S, p = T._ref_pan(grid=None)
e = sk.Engine(0); T._engine_for(e, S, p)
Lv = S.prepare_dust(True)
ref, gpu = [], []
for b in range(B):
    S.reset(100 + 1000 * b); S.run_dust(True, 0.1); ref.append(S.labs_dust().copy())
    Npp = S.packages_per_lambda()
    e.reset_labs_dust(); e.run_dust(1, Lv, Npp, seed=70 + b); gpu.append(e.fetch_labs_dust())
a, r = np.array(gpu), np.array(ref)          # [B, Ncells, Nlambda]
print("totals ratio", a.sum() / r.sum(), "Npp", Npp)
for name, aa, rr in (("cell x lambda", a.reshape(B, -1), r.reshape(B, -1)), ("per lambda", a.sum(1), r.sum(1)), ("per cell", a.sum(2), r.sum(2))):
    ma, mr = aa.mean(0), rr.mean(0); sa, sr_ = aa.std(0, ddof=1) / np.sqrt(B), rr.std(0, ddof=1) / np.sqrt(B)
    for sig in (0.3, 0.15, 0.05):
        ok = (sa > 0) & (sr_ > 0) & (sa < sig * ma) & (sr_ < sig * mr)
        if ok.sum() < 3: continue
        z = (ma[ok] - mr[ok]) / np.sqrt(sa[ok] ** 2 + sr_[ok] ** 2)
        print(f"  {name:14s} signal<{sig}: bins {ok.sum():6d} beyond3 {np.mean(np.abs(z)>=3):.4f} mean z {z.mean():+.3f} std z {z.std():.3f} max {np.abs(z).max():.1f}  var ratio gpu/ref {np.median(sa[ok]/sr_[ok]):.3f}")
