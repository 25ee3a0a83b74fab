"""Generates skirt_b200/data/interstellar_dustmix.json: the reference's InterstellarDustMix
(InterstellarDustMix.cpp:21-58, data in dat/DustMix/InterstellarDustMix.dat) resampled by the reference's
own code (DustMix::addpopulation, DustMix.cpp:300-321) onto a 256-point logarithmic wavelength grid.
Needs /root/reference and oracle/_ref; run in the build container only."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import PC, box_line, C1_BOX
from oracle import skirtref as sr

N = 256
spec = "\n".join(["sim pan", "threads 1", "packages 10", f"loggrid 0.02e-6 3000e-6 {N}", box_line(C1_BOX),
                  "grid cartesian 2 2 2 lin lin lin", "dustsamples 1",
                  f"stellar expdisk {4000*PC!r} {350*PC!r} 0 0", f"dust 1.0 0.55e-6 expdisk {4000*PC!r} {140*PC!r} 0 0",
                  f"instrument sed s 1e23 0 0 0"]) + "\n"
S = sr.RefSim(spec, luminosities=[np.ones(N)], mixes=[(np.ones(N), np.ones(N), np.zeros(N))]).setup()
lam, _ = S.wavelengths()
kabs, ksca, g = S.interstellar_mix()
out = dict(source="SKIRT v7.3 InterstellarDustMix resampled by oracle/_ref (tools/make_dustmix_table.py)",
           lambda_m=lam.tolist(), kappa_abs=kabs.tolist(), kappa_sca=ksca.tolist(), asymmpar=g.tolist())
path = os.path.join(ROOT, "skirt_b200", "data", "interstellar_dustmix.json")
json.dump(out, open(path, "w"))
print("wrote", path, "kext(V)~", np.interp(0.55e-6, lam, kabs + ksca))
