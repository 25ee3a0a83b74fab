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


SPIRAL = dict(arms=2, pitch=math.radians(20), radius=4000 * PC, phase=0.0, weight=1.0, index=1)


def c3_params(maxlevel=8, packages=1e9, massfrac=1e-6, pixels=400):
    """C3: adaptive OctTreeDustGrid (levels 2..maxlevel, maxMassFraction 1e-6, Neighbor search), stars in a two-armed
    spiral disk (SpiralStructureGeometryDecorator on the ExpDisk; the dust disk stays axisymmetric, as the reference's
    FaceOnDustCompNormalization demands, FaceOnDustCompNormalization.cpp:72), forced scattering, 6 peel-off
    FrameInstruments at i = 0, 30, 60, 80, 88, 90 degrees; oligochromatic, 1e9 packets"""
    disk = lambda hz: dict(geometry="expdisk", hR=4000 * PC, hz=hz)
    return dict(name="C3", sim="oligo", wavelengths=[0.55e-6], box=C1_BOX, packages=packages,
                grid=dict(kind="octtree", minLevel=2, maxLevel=maxlevel, maxMassFraction=massfrac, searchMethod="Neighbor"),
                stellar=[dict(disk(350 * PC), spiral=SPIRAL, L=[5e9 * LSUN * 1e-3])],
                dust=[dict(disk(140 * PC), tau=1.0, lam=0.55e-6)],
                instruments=[dict(kind="frame", name=f"i{inc}", distance=10e6 * PC, inclination=math.radians(inc),
                                  Nxp=pixels, fovxp=50000 * PC, Nyp=pixels, fovyp=50000 * PC) for inc in (0, 30, 60, 80, 88, 90)])


def c4_params(particles=1000000, nlambda=100, packages=1e7):
    """C4: VoronoiDustGrid over synthetic SPH particles drawn from the ExpDisk density (Voro++ tessellation), 100-point
    logarithmic wavelength grid, SED + frame instruments; 1e9 packets = 100 wavelengths x 1e7"""
    return dict(name="C4", sim="pan", loggrid=(0.1e-6, 1000e-6, nlambda), box=C1_BOX, packages=packages,
                grid=dict(kind="voronoi", particles=int(particles), seed=0x5eed0004),
                stellar=[dict(geometry="expdisk", hR=4000 * PC, hz=350 * PC, T=8000.0, Lbol=8e9 * LSUN)],
                dust=[dict(geometry="expdisk", hR=4000 * PC, hz=140 * PC, tau=1.0, lam=0.55e-6)],
                instruments=[dict(kind="frame", name="frame60", distance=10e6 * PC, inclination=math.radians(60),
                                  Nxp=400, fovxp=50000 * PC, Nyp=400, fovyp=50000 * PC),
                             dict(kind="sed", name="sed60", distance=10e6 * PC, inclination=math.radians(60))])


def c5_params(depth=6, root=16, nlambda=50, packages=2e8, frac=None):
    """C5: full panchromatic run with dust emission and self-absorption cycles on an adaptive mesh: synthetic AMR, root
    16^3, 2x2x2 refinement where (peak density) x volume exceeds a threshold, down to `depth` levels; dust density =
    the mesh's own field (AdaptiveMeshDustDistribution); 1e10 packets = 50 wavelengths x 2e8, sharded over 8 GPUs"""
    if frac is None:
        frac = 2e-6 * 8.0 ** (6 - depth)
    return dict(name="C5", sim="pan", loggrid=(0.1e-6, 1000e-6, nlambda), box=C1_BOX, packages=packages, dustemission=True, selfabsorption=True,
                grid=dict(kind="amesh", root=int(root), depth=int(depth), frac=float(frac), tau=1.0, lam=0.55e-6),
                stellar=[dict(geometry="sersic", index=2.0, Re=1600 * PC, q=0.7, T=3500.0, Lbol=3e9 * LSUN),
                         dict(geometry="expdisk", hR=4000 * PC, hz=350 * PC, T=10000.0, Lbol=5e9 * LSUN)],
                dust=[dict(geometry="mesh", tau=1.0, lam=0.55e-6)],
                instruments=[dict(kind="frame", name="frame88", distance=10e6 * PC, inclination=math.radians(88),
                                  Nxp=800, fovxp=50000 * PC, Nyp=200, fovyp=12500 * PC),
                             dict(kind="sed", name="sed88", distance=10e6 * PC, inclination=math.radians(88))])


def sph_particles(n, box=C1_BOX, seed=0x5eed0004, hR=4000 * PC, hz=140 * PC):
    """synthetic SPH particle positions: exponential disk (R ~ Gamma(2, hR), z ~ Laplace(hz)) clipped to the box, plus 20 %
    uniform background so that the whole domain is tessellated into reasonable cells"""
    rng = np.random.Generator(np.random.Philox(key=seed))
    box = np.asarray(box, dtype=np.float64); lo = box[0::2]; hi = box[1::2]
    out = np.zeros((0, 3))
    while len(out) < n:
        m = n
        R = rng.gamma(2.0, hR, m); phi = rng.random(m) * 2 * np.pi; z = rng.laplace(0.0, hz, m)
        pts = np.concatenate([np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1), lo + rng.random((m // 4, 3)) * (hi - lo)])
        pts = pts[np.all((pts > lo) & (pts < hi), axis=1)]
        out = np.concatenate([out, pts])
    out = out[:n]
    rng.shuffle(out)
    return np.ascontiguousarray(out)


def synthetic_amesh(box=C1_BOX, root=16, depth=6, frac=2e-6, hR=4000 * PC, hz=140 * PC):
    """Synthetic adaptive mesh in the order of an adaptive mesh data file (AdaptiveMeshAsciiFile.cpp:43-100: depth first,
    children k -> j -> i): root x root x root children below the root, each refined 2x2x2 while the largest of 27 probe
    densities times the cell volume exceeds frac x (total mass), down to `depth` levels below the root cells.
    Returns (nxyz[N, 3], value[N]): value = un-normalised exponential-disk density at the centre of a leaf."""
    box = np.asarray(box, dtype=np.float64); lo0 = box[0::2]; hi0 = box[1::2]
    dens = lambda p: np.exp(-np.hypot(p[..., 0], p[..., 1]) / hR) * np.exp(-np.abs(p[..., 2]) / hz)
    total = 2 * np.pi * hR ** 2 * 2 * hz
    B = 9 ** depth                                  # path digits: 0 = the node itself, 1..8 its children; root children in front
    n = root
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")          # local Morton order: i fastest
    idx = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    lo = lo0 + idx * (hi0 - lo0) / n; hi = lo0 + (idx + 1) * (hi0 - lo0) / n                 # Box::fracpos
    key = (np.arange(n ** 3, dtype=np.int64) + 1) * B
    keys, nx, val = [np.zeros(1, np.int64)], [np.full((1, 3), n, np.int32)], [np.zeros(1)]
    off = np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)])              # k -> j -> i, i fastest
    probes = np.array([[a, b, c] for a in (0.0, 0.5, 1.0) for b in (0.0, 0.5, 1.0) for c in (0.0, 0.5, 1.0)])
    for level in range(depth + 1):
        w = hi - lo
        peak = np.max(dens(lo[:, None, :] + probes[None, :, :] * w[:, None, :]), axis=1)
        refine = (peak * np.prod(w, axis=1) > frac * total) & (level < depth)
        keys.append(key); val.append(np.where(refine, 0.0, dens(0.5 * (lo + hi))))
        nx.append(np.where(refine[:, None], 2, 0).astype(np.int32) * np.ones((1, 3), np.int32))
        if not refine.any():
            break
        plo, phi_, pkey = lo[refine], hi[refine], key[refine]
        step = 9 ** (depth - level - 1)
        lo = (plo[:, None, :] + off[None, :, :] * (phi_ - plo)[:, None, :] / 2).reshape(-1, 3)
        hi = (plo[:, None, :] + (off[None, :, :] + 1) * (phi_ - plo)[:, None, :] / 2).reshape(-1, 3)
        key = (pkey[:, None] + (np.arange(8, dtype=np.int64)[None, :] + 1) * step).ravel()
    keys = np.concatenate(keys); order = np.argsort(keys, kind="stable")
    return np.concatenate(nx)[order], np.concatenate(val)[order]


def _geometry(g):
    if g["geometry"] == "expdisk":
        geom = sim.ExpDiskGeometry(g["hR"], g["hz"], g.get("Rmax", 0.0), g.get("zmax", 0.0))
    elif g["geometry"] == "sersic":
        geom = sim.SersicGeometry(g["index"], g["Re"], g.get("q", 1.0))
    elif g["geometry"] == "mesh":
        return None             # the dust is the adaptive mesh's own density field
    else:
        raise sim.FatalError(f"unknown geometry {g['geometry']}")
    sp = g.get("spiral")
    if sp:
        geom = sim.SpiralStructureGeometryDecorator(geom, sp["arms"], sp["pitch"], sp["radius"], sp["phase"], sp["weight"], sp["index"])
    return geom


def dust_grid(p, lg=None, mix=None):
    """the dust grid of a parameter dict (and, for the adaptive mesh, the density units that give the requested face-on
    optical depth through the centre)"""
    b = p["box"]; g = p.get("grid")
    if g is None:
        n = p["n"]
        return sim.CartesianDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], sim.LinMesh(n), sim.LinMesh(n), sim.LinMesh(n))
    if g["kind"] in ("octtree", "bintree"):
        cls = sim.OctTreeDustGrid if g["kind"] == "octtree" else sim.BinTreeDustGrid
        return cls(b[0], b[1], b[2], b[3], b[4], b[5], g["minLevel"], g["maxLevel"], g.get("searchMethod", "Neighbor"),
                   g.get("sampleCount", 100), g.get("maxOpticalDepth", 0.0), g.get("maxMassFraction", 1e-6))
    if g["kind"] == "voronoi":
        return sim.VoronoiDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], sph_particles(g["particles"], b, g.get("seed", 0x5eed0004)))
    if g["kind"] == "amesh":
        hz = 140 * PC
        nxyz, val = synthetic_amesh(b, g["root"], g["depth"], g["frac"], hz=hz)
        # face-on optical depth tau at lam through the centre: integral of exp(-|z|/hz) dz = 2 hz (the box cuts it at +-5 kpc)
        kv = float(10.0 ** np.interp(math.log10(g["lam"]), np.log10(lg.lambdav), np.log10(mix.kappaext))) if lg.Nlambda > 1 else float(mix.kappaext[0])
        units = g["tau"] / (kv * 2 * hz * (1.0 - math.exp(-0.5 * (b[5] - b[4]) / hz)))
        return sim.AdaptiveMeshDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], nxyz, val, densityUnits=units)
    raise sim.FatalError(f"unknown grid kind {g['kind']}")


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
    grid = dust_grid(p, lg, mix)
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
