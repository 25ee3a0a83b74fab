"""Synthetic configurations of BASELINE.json assembled from the host-side mirror classes.

Each configuration is described once as plain data (`*_params`) so that the engine set-up below and the
reference-side description used by the tests and the benchmark's CPU legs are built from the same numbers."""
import math

import numpy as np

from . import simulation as sim

PC = sim.PC
LSUN = sim.LSUN

C1_BOX = (-25000 * PC, 25000 * PC, -25000 * PC, 25000 * PC, -5000 * PC, 5000 * PC)


def c1_params(n=100, packages=1e6):
    """C1: oligochromatic edge-on ExpDisk stars+dust, Cartesian n^3, 1 wavelength, FrameInstrument
    (doc/Part 1 - User Guide/SKIRT/Tutorial 1.txt:143-150,241-245,352-372)."""
    return dict(name="C1", sim="oligo", wavelengths=[0.55e-6], box=C1_BOX, n=n, packages=packages,
                stellar=[dict(geometry="expdisk", hR=4000 * PC, hz=350 * PC, L=[5e9 * LSUN * 1e-3])],
                dust=[dict(geometry="expdisk", hR=4000 * PC, hz=140 * PC, tau=1.0, lam=0.55e-6)],
                instruments=[dict(kind="frame", name="i88", distance=10e6 * PC, inclination=math.radians(88),
                                  Nxp=800, fovxp=50000 * PC, Nyp=200, fovyp=12500 * PC)])


def c2_params(n=100, nlambda=50, packages=2e6):
    """C2 (the configuration BASELINE.json's metric is quoted on): panchromatic Sersic bulge + exponential disk,
    50-wavelength logarithmic grid 0.1-1000 micron, InterstellarDustMix, absorption stored (no dust emission phase),
    SED + frame instruments; 1e8 packets = 50 wavelengths x 2e6 packets per wavelength."""
    return dict(name="C2", sim="pan", loggrid=(0.1e-6, 1000e-6, nlambda), box=C1_BOX, n=n, packages=packages,
                stellar=[dict(geometry="sersic", index=2.0, Re=1600 * PC, q=0.7, T=3500.0, Lbol=3e9 * LSUN),
                         dict(geometry="expdisk", hR=4000 * PC, hz=350 * PC, T=10000.0, Lbol=5e9 * LSUN)],
                dust=[dict(geometry="expdisk", hR=4000 * PC, hz=140 * PC, tau=1.0, lam=0.55e-6)],
                instruments=[dict(kind="frame", name="frame88", distance=10e6 * PC, inclination=math.radians(88),
                                  Nxp=800, fovxp=50000 * PC, Nyp=200, fovyp=12500 * PC),
                             dict(kind="sed", name="sed88", distance=10e6 * PC, inclination=math.radians(88))])


def _geometry(g):
    if g["geometry"] == "expdisk":
        return sim.ExpDiskGeometry(g["hR"], g["hz"], g.get("Rmax", 0.0), g.get("zmax", 0.0))
    if g["geometry"] == "sersic":
        return sim.SersicGeometry(g["index"], g["Re"], g.get("q", 1.0))
    raise sim.FatalError(f"unknown geometry {g['geometry']}")


def wavelength_grid(p):
    if p["sim"] == "oligo":
        return sim.OligoWavelengthGrid(p["wavelengths"])
    return sim.LogWavelengthGrid(*p["loggrid"])


def luminosities(p, lg):
    """per stellar component the luminosity per wavelength bin (W)"""
    out = []
    for s in p["stellar"]:
        if "L" in s:
            out.append(np.asarray(s["L"], dtype=np.float64))
        else:
            B = sim.planck_lambda(lg.lambdav, s["T"]) * lg.dlambdav
            out.append(s["Lbol"] * B / B.sum())
    return out


def build(p, device=0, rank=0, nranks=1, seed=4357, storeAbsorption=None, rho=None, engine=None):
    """MonteCarloSimulation (engine side) for a parameter dict"""
    lg = wavelength_grid(p)
    mix = sim.InterstellarDustMix(lg)
    n = p["n"]; b = p["box"]
    grid = sim.CartesianDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], sim.LinMesh(n), sim.LinMesh(n), sim.LinMesh(n))
    ds = sim.DustSystem(grid, [sim.DustComp(_geometry(d), mix, d["tau"], d["lam"]) for d in p["dust"]], lg, rho=rho)
    ss = sim.StellarSystem([sim.StellarComp(_geometry(s), L) for s, L in zip(p["stellar"], luminosities(p, lg))])
    ins = []
    for i in p["instruments"]:
        if i["kind"] == "sed":
            ins.append(sim.SEDInstrument(i["name"], i["distance"], i["inclination"], i.get("azimuth", 0.0), i.get("positionAngle", 0.0)))
        else:
            cls = sim.FrameInstrument if i["kind"] == "frame" else sim.SimpleInstrument
            ins.append(cls(i["name"], i["distance"], i["inclination"], i.get("azimuth", 0.0), i.get("positionAngle", 0.0),
                           i["Nxp"], i["fovxp"], i["Nyp"], i["fovyp"]))
    if storeAbsorption is None:
        storeAbsorption = p["sim"] == "pan"
    return sim.MonteCarloSimulation(lg, ss, ds, sim.InstrumentSystem(ins), packages=p["packages"], seed=seed,
                                    storeAbsorption=storeAbsorption, device=device, rank=rank, nranks=nranks, engine=engine)


def c1_oligo(n=100, packages=1e6, **kw):
    return build(c1_params(n, packages), **kw)


def c2_pan(n=100, nlambda=50, packages=2e6, **kw):
    return build(c2_params(n, nlambda, packages), **kw)
