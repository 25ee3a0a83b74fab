"""Shared helpers for the test-suite: synthetic configurations, seeded ray batches,
and the reference harness (oracle/_ref) specs that describe them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PC = 3.08567758e16          # Units.cpp:17-30
GOLDEN = os.path.join(ROOT, "tests", "golden")

# InterstellarDustMix at 0.55 micron on an OligoWavelengthGrid, evaluated by the reference itself
# (oracle/_ref, skr_interstellar_mix); regenerate with tests/golden/make_golden.py
MIX_V = dict(kabs=848.25415325, ksca=1751.62943942, g=0.53723952)

C1_BOX = np.array([-25000., 25000., -25000., 25000., -5000., 5000.]) * PC


def box_line(box):
    return "box " + " ".join(repr(float(v)) for v in box)


def rays(n, box, seed, scale=1.2):
    """SURVEY.md 8d synthetic rays: r uniform in `scale` x bounding box, k isotropic."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float64)
    c = 0.5 * (box[0::2] + box[1::2]); w = (box[1::2] - box[0::2])
    r = c + (rng.random((n, 3)) - 0.5) * w * scale
    k = rng.normal(size=(n, 3)); k /= np.linalg.norm(k, axis=1)[:, None]
    return np.ascontiguousarray(r), np.ascontiguousarray(k)


def adversarial_rays(box, axes=None, seed=7):
    """axis-aligned, |k_i|<1e-15, starts on faces/edges/corners, grazing, outside & pointing away."""
    box = np.asarray(box, dtype=np.float64)
    lo = box[0::2]; hi = box[1::2]; c = 0.5 * (lo + hi); w = hi - lo
    rs, ks = [], []
    e = np.eye(3)
    for a in range(3):
        for sgn in (1.0, -1.0):
            for start in (c, lo - 0.1 * w, hi + 0.1 * w, lo, hi, c + 0.25 * w):
                rs.append(np.array(start, dtype=np.float64)); ks.append(sgn * e[a])
    # tiny components
    for a in range(3):
        k = np.array([0.6, 0.8, 0.0]); k = np.roll(k, a); k[(a + 2) % 3] = 5e-16
        rs.append(c + 0.01 * w); ks.append(k / np.linalg.norm(k))
        rs.append(lo - 0.2 * w); ks.append(np.abs(k) / np.linalg.norm(k))
    # corners and edges, diagonal directions
    for sx in (0, 1):
        for sy in (0, 1):
            for sz in (0, 1):
                corner = np.where([sx, sy, sz], hi, lo)
                d = c - corner
                rs.append(corner); ks.append(d / np.linalg.norm(d))
                rs.append(corner - 0.05 * d); ks.append(d / np.linalg.norm(d))
                rs.append(corner); ks.append(-d / np.linalg.norm(d))
    # grazing along a face
    for a in range(3):
        k = np.zeros(3); k[(a + 1) % 3] = 1.0
        p = c.copy(); p[a] = hi[a]; p[(a + 1) % 3] = lo[(a + 1) % 3] - 0.1 * w[(a + 1) % 3]
        rs.append(p); ks.append(k)
        p2 = p.copy(); p2[a] = lo[a]
        rs.append(p2); ks.append(k)
    if axes is not None:
        # starts exactly on interior cell borders
        rng = np.random.default_rng(seed)
        xv, yv, zv = axes
        for _ in range(24):
            p = np.array([rng.choice(xv), rng.choice(yv), rng.choice(zv)])
            k = rng.normal(size=3); k /= np.linalg.norm(k)
            rs.append(p); ks.append(k)
    return np.ascontiguousarray(np.array(rs)), np.ascontiguousarray(np.array(ks))


def spec_c1(packages=1e5, n=40, threads=1, seed=4357, mesh="lin", grid=None, dustsamples=20, storeabs=0, instruments=None,
            tau=1.0, box=C1_BOX):
    """C1: oligochromatic edge-on ExpDisk stars+dust (Tutorial 1.txt:241-245,352-372), 1 wavelength."""
    if grid is None:
        grid = f"grid cartesian {n} {n} {n} {mesh} {mesh} {mesh}"
    if instruments is None:
        instruments = [f"instrument frame i88 {1e7*PC!r} {float(np.radians(88))!r} 0 0 200 {50000*PC!r} 50 {12500*PC!r}",
                       f"instrument sed s88 {1e7*PC!r} {float(np.radians(88))!r} 0 0"]
    lines = ["sim oligo", f"threads {threads}", f"seed {seed}", f"packages {packages!r}", "wavelengths 0.55e-6",
             box_line(box), grid, f"dustsamples {dustsamples}", f"storeabs {storeabs}",
             f"stellar expdisk {4000*PC!r} {350*PC!r} 0 0",
             f"dust {tau!r} 0.55e-6 expdisk {4000*PC!r} {140*PC!r} 0 0"] + list(instruments)
    return "\n".join(lines) + "\n"


def mix_v():
    return [([MIX_V["kabs"]], [MIX_V["ksca"]], [MIX_V["g"]])]


def assert_paths_equal(a, b, rtol=1e-12, label=""):
    """bit-exact cell sequences + counts; ds/s/dtau/tau to rtol relative (north_star: 1e-12)."""
    assert np.array_equal(a["offsets"], b["offsets"]), f"{label}: segment counts differ"
    assert np.array_equal(a["m"], b["m"]), f"{label}: cell index sequences differ"
    for key in ("ds", "s", "dtau", "tau"):
        x, y = a[key], b[key]
        denom = np.maximum(np.abs(y), 1e-300)
        err = np.max(np.abs(x - y) / denom) if len(x) else 0.0
        assert err <= rtol, f"{label}: {key} relative error {err:g} > {rtol:g}"


def paths_bit_identical(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("offsets", "m", "ds", "s", "dtau", "tau"))


# ---- synthetic adaptive mesh: refine 2x2x2 where rho*V exceeds a threshold ----------
def expdisk_density(x, y, z, hR=4000 * PC, hz=140 * PC):
    R = np.hypot(x, y)
    return np.exp(-R / hR) * np.exp(-np.abs(z) / hz)


def make_amesh(box=C1_BOX, root=(4, 4, 4), max_depth=4, frac=2e-3):
    """Returns (nxyz[N,3], value[N]) in the reference's file order (AdaptiveMeshAsciiFile.cpp:43-100:
    depth-first, children looped k -> j -> i)."""
    box = np.asarray(box, dtype=np.float64)
    lo = box[0::2]; hi = box[1::2]
    total = 2 * np.pi * (4000 * PC) ** 2 * 2 * 140 * PC      # integral of the un-normalised density
    nxyz, val = [], []

    def visit(lo, hi, depth, n):
        nxyz.append(n); val.append(0.0)
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    idx = np.array([i, j, k]); nn = np.array(n)
                    clo = lo + idx * (hi - lo) / nn
                    chi = lo + (idx + 1) * (hi - lo) / nn
                    c = 0.5 * (clo + chi)
                    rho = float(expdisk_density(*c))
                    # refinement criterion: largest density over a 3x3x3 probe lattice times the cell volume
                    g3 = [clo + f * (chi - clo) for f in (0.0, 0.5, 1.0)]
                    peak = max(float(expdisk_density(gx[0], gy[1], gz[2])) for gx in g3 for gy in g3 for gz in g3)
                    mass = peak * np.prod(chi - clo)
                    if depth < max_depth and mass > frac * total:
                        visit(clo, chi, depth + 1, (2, 2, 2))
                    else:
                        nxyz.append((0, 0, 0)); val.append(rho)

    visit(lo, hi, 0, tuple(root))
    return np.array(nxyz, dtype=np.int32), np.array(val, dtype=np.float64)


# the grids with symmetries, as spec lines of the reference harness (radius 18 kpc; the C1 disk has hR = 4 kpc, hz = 140 pc)
SYM_GRIDS = {
    "sphere1d": f"sphere1d {18000*PC!r} 40 pow 30.0",
    "sphere2d": f"sphere2d {18000*PC!r} 30 pow 20.0 16 lin",                  # 16 polar bins: a border at pi/2 exists
    "sphere2d_odd": f"sphere2d {18000*PC!r} 24 lin 9 lin",                    # 9 polar bins: the reference inserts the border at pi/2
    "cylinder2d": f"cylinder2d {18000*PC!r} {-3000*PC!r} {3000*PC!r} 36 pow 25.0 30 sympow 15.0",
}


SPHERE1D_DUST_MASS = 1e37          # kg: a V-band optical depth of order one through the centre of the Sersic sphere


def sym_grid_mirror(kind):
    """the same grid through the product's host mirror (skirt_b200.simulation)"""
    from skirt_b200 import simulation as sim
    if kind == "sphere1d":
        return sim.Sphere1DDustGrid(18000 * PC, sim.PowMesh(40, 30.0))
    if kind == "sphere2d":
        return sim.Sphere2DDustGrid(18000 * PC, sim.PowMesh(30, 20.0), sim.LinMesh(16))
    if kind == "sphere2d_odd":
        return sim.Sphere2DDustGrid(18000 * PC, sim.LinMesh(24), sim.LinMesh(9))
    return sim.Cylinder2DDustGrid(18000 * PC, -3000 * PC, 3000 * PC, sim.PowMesh(36, 25.0), sim.SymPowMesh(30, 15.0))


def spec_grid(kind, search=1, box=C1_BOX, minlevel=2, maxlevel=5, massfrac=2e-4, packages=1e4, threads=1, extra=()):
    """Specs for the non-Cartesian grids; dust = C1's ExpDisk (or the mesh's own densities for amesh)."""
    head = ["sim oligo", f"threads {threads}", "seed 4357", f"packages {packages!r}", "wavelengths 0.55e-6", box_line(box)]
    stellar = f"stellar expdisk {4000*PC!r} {350*PC!r} 0 0"
    dust = f"dust 1.0 0.55e-6 expdisk {4000*PC!r} {140*PC!r} 0 0"
    instr = f"instrument sed s88 {1e7*PC!r} {float(np.radians(88))!r} 0 0"
    if kind in ("octtree", "bintree"):
        grid = f"grid {kind} {minlevel} {maxlevel} {search} {massfrac!r} 0 50"
        lines = head + [grid, "dustsamples 10", stellar, dust, instr]
    elif kind == "voronoi":
        lines = head + ["grid voronoi file", "dustsamples 10", stellar, dust, instr]
    elif kind == "amesh":
        lines = head + ["grid amesh", "ameshdust 1e-24", stellar, instr]
    elif kind in SYM_GRIDS:
        if kind == "sphere1d":      # DustGrid.cpp:37: the grid's dimension must not be lower than the geometries' -> spherical stars and dust
            stellar = f"stellar sersic 2.0 {1500*PC!r} 1.0"; dust = f"dustmass {SPHERE1D_DUST_MASS!r} sersic 1.5 {3000*PC!r} 1.0"
        lines = head + ["grid " + SYM_GRIDS[kind], "dustsamples 10", stellar, dust, instr]
    elif kind in ("particletree_oct", "particletree_bin"):      # maxlevel doubles as the number of extra levels (0 or 1 in the tests)
        lines = head + [f"grid particletree {kind[-3:]} {0 if maxlevel > 1 else maxlevel}", "dustsamples 10", stellar, dust, instr]
    else:
        raise ValueError(kind)
    return "\n".join(lines + list(extra)) + "\n"


def voronoi_particles(n, box=C1_BOX, seed=11):
    """particles drawn from the ExpDisk density (clipped to the box) plus a uniform background."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float64); lo = box[0::2]; hi = box[1::2]
    out = []
    while len(out) < n:
        m = n
        R = rng.gamma(2.0, 4000 * PC, m); phi = rng.random(m) * 2 * np.pi
        z = rng.laplace(0.0, 350 * PC, m)
        p = np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1)
        u = lo + rng.random((m // 4, 3)) * (hi - lo)
        p = np.concatenate([p, u])
        ok = np.all((p > lo) & (p < hi), axis=1)
        out.extend(p[ok].tolist())
    out = np.array(out[:n])
    rng.shuffle(out)
    return np.ascontiguousarray(out)


# ---- configuration dictionaries shared by the reference harness spec and the engine set-up -----------------
def cfg_c1(n=40, packages=1e5, instruments=None, storeabs=0, tau=1.0, grid=None, seed=4357, threads=1, dustsamples=20):
    """C1 (BASELINE.json configs[0]): oligochromatic edge-on ExpDisk stars + dust, 1 wavelength, FrameInstrument."""
    if instruments is None:
        instruments = [dict(kind=1, name="i88", distance=1e7 * PC, inclination=float(np.radians(88)), azimuth=0.0, positionAngle=0.0,
                            Nxp=200, fovxp=50000 * PC, Nyp=50, fovyp=12500 * PC),
                       dict(kind=2, name="s88", distance=1e7 * PC, inclination=float(np.radians(88)), azimuth=0.0, positionAngle=0.0)]
    return dict(sim="oligo", wavelengths=[0.55e-6], box=C1_BOX, packages=packages, seed=seed, threads=threads,
                grid=grid or f"grid cartesian {n} {n} {n} lin lin lin", dustsamples=dustsamples, storeabs=storeabs,
                sources=[dict(geometry=1, p=[4000 * PC, 350 * PC, 0.0, 0.0, 0.0], L=[1.0])],
                dust=[dict(tau=tau, lam=0.55e-6, geometry=1, p=[4000 * PC, 140 * PC, 0.0, 0.0, 0.0], mix=mix_v()[0])],
                instruments=instruments, mwr=1e4, minscatt=0.0, xi=0.5, ebias=0.5)


FULL_NSCATT = 2


def cfg_full(threads=1):
    """C1 geometry at tau_V = 3 seen by a FullInstrument with 2 scattering levels plus an SEDInstrument in the same
    direction (tests/golden/mc_full.npz, made by tests/golden/make_full_golden.py)"""
    ins = [dict(kind=4, name="full", distance=1e7 * PC, inclination=float(np.radians(80)), azimuth=0.0, positionAngle=0.0,
                Nxp=40, fovxp=50000 * PC, Nyp=16, fovyp=20000 * PC, scatteringLevels=FULL_NSCATT),
           dict(kind=2, name="s80", distance=1e7 * PC, inclination=float(np.radians(80)), azimuth=0.0, positionAngle=0.0)]
    return cfg_c1(n=24, packages=2e5, instruments=ins, tau=3.0, threads=threads)


def _geom_words(g):
    if g["geometry"] == 1:
        p = g["p"]; w = f"expdisk {p[0]!r} {p[1]!r} {p[2]!r} {p[3]!r}"
    else:
        p = g["p"]; w = f"sersic {g['n']!r} {p[0]!r} {p[1]!r}"
    sp = g.get("spiral")
    if sp:
        w += f" spiral {sp['arms']} {sp['pitch']!r} {sp['radius']!r} {sp['phase']!r} {sp['weight']!r} {sp['index']}"
    return w


def ref_spec(cfg):
    lines = [f"sim {cfg['sim']}", f"threads {cfg.get('threads', 1)}", f"seed {cfg.get('seed', 4357)}", f"packages {float(cfg['packages'])!r}",
             f"minweightreduction {cfg.get('mwr', 1e4)!r}", f"minscatt {cfg.get('minscatt', 0.0)!r}", f"scattbias {cfg.get('xi', 0.5)!r}",
             f"emissionbias {cfg.get('ebias', 0.5)!r}"]
    if cfg["sim"] == "oligo":
        lines.append("wavelengths " + " ".join(repr(float(v)) for v in cfg["wavelengths"]))
    else:
        lg = cfg["loggrid"]; lines.append(f"loggrid {lg[0]!r} {lg[1]!r} {lg[2]}")
    lines += [box_line(cfg["box"]), cfg["grid"], f"dustsamples {cfg.get('dustsamples', 20)}", f"storeabs {cfg.get('storeabs', 0)}"]
    if "units" in cfg:
        lines.append(f"units {cfg['units'][0]} {cfg['units'][1]}")
    for s in cfg["sources"]:
        lines.append("stellar " + _geom_words(s))
    for d in cfg.get("dust", []):
        lines.append(f"dust {d['tau']!r} {d['lam']!r} " + _geom_words(d))
    if cfg.get("ameshdust"):
        lines.append(f"ameshdust {cfg['ameshdust']!r}")
    for ins in cfg["instruments"]:
        if ins["kind"] == 6:
            lines.append(f"perspective {ins['name']} {ins['Nxp']} {ins['Nyp']} {ins['fovxp']!r} " +
                         " ".join(repr(float(ins[k])) for k in ("viewX", "viewY", "viewZ", "crossX", "crossY", "crossZ", "upX", "upY", "upZ", "focal")))
            continue
        kind = {1: "frame", 2: "sed", 3: "simple", 4: "full", 5: "multiframe"}[ins["kind"]]
        w = f"instrument {kind} {ins['name']} {ins['distance']!r} {ins['inclination']!r} {ins.get('azimuth', 0.0)!r} {ins.get('positionAngle', 0.0)!r}"
        if ins["kind"] == 5:
            w += f" {int(ins.get('writeTotal', True))} {int(ins.get('writeStellarComps', False))} {len(ins['frames'])}"
            for f in ins["frames"]:
                w += f" {f['Nxp']} {f['fovxp']!r} {f['Nyp']} {f['fovyp']!r} {f.get('xpc', 0.0)!r} {f.get('ypc', 0.0)!r}"
            lines.append(w)
            continue
        if ins["kind"] != 2:
            w += f" {ins['Nxp']} {ins['fovxp']!r} {ins['Nyp']} {ins['fovyp']!r}"
        if ins["kind"] == 4:
            w += f" {ins.get('scatteringLevels', 0)}"
        lines.append(w)
    return "\n".join(lines) + "\n"


def make_ref(cfg, **kw):
    from oracle import skirtref as sr
    return sr.RefSim(ref_spec(cfg), luminosities=[s["L"] for s in cfg["sources"]],
                     mixes=[d["mix"] for d in cfg.get("dust", [])] + list(cfg.get("extra_mixes", [])), **kw)


def setup_engine(e, cfg, tables, medium, L=None):
    """feeds one engine with the flattened state (tables/medium as produced by the reference harness or
    by the host-side builders) plus sources and instruments of cfg"""
    e.set_grid(tables)
    e.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    if L is None:
        L = np.array([s["L"] for s in cfg["sources"]], dtype=np.float64)
    e.sources(cfg["sources"], L, cfg.get("ebias", 0.5))
    e.instruments(cfg["instruments"])
    return e


# ---- golden fixtures (tests/golden/*.npz, generated from the reference by tests/golden/make_golden.py) ---------
GEOM_CASES = ["cart_lin", "cart_sympow", "octtree_s0", "octtree_s1", "octtree_s2", "bintree_s0", "bintree_s1", "amesh", "voronoi"]
# fixtures of tests/golden/make_symmetric_golden.py: grids with symmetries, a particle tree, a barycentric octree
GEOM_CASES_MORE = ["sphere1d", "sphere2d", "sphere2d_odd", "cylinder2d", "particletree_oct", "octtree_bary_s1"]


def load_golden(name):
    """-> (tables, medium, data) of one geometry fixture; data holds rays and the reference's answers"""
    z = np.load(os.path.join(GOLDEN, f"geom_{name}.npz"))
    tables, medium, data = {}, {}, {}
    for key in z.files:
        v = z[key]
        if key.startswith("grid_"):
            k2 = key[5:]
            if k2 == "kind":
                v = str(v)
            elif v.ndim == 0:
                v = v.item()
            tables[k2] = v
        elif key.startswith("med_"):
            medium[key[4:]] = v
        else:
            data[key] = v
    data["paths"] = {k[5:]: data.pop(k) for k in list(data) if k.startswith("path_")}
    return tables, medium, data


def load_golden_mc(name="mc_c1"):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    tables = {k[5:]: (str(z[k]) if k == "grid_kind" else z[k]) for k in z.files if k.startswith("grid_")}
    medium = {k[4:]: z[k] for k in z.files if k.startswith("med_")}
    data = {k: z[k] for k in z.files if not k.startswith(("grid_", "med_"))}
    return tables, medium, data


def zscores(mean_a, sem_a, mean_b, sem_b):
    sig = np.sqrt(sem_a ** 2 + sem_b ** 2)
    ok = sig > 0
    return (mean_a - mean_b)[ok] / sig[ok]


# ---- the Monte Carlo gate of SURVEY.md 8d(ii) ---------------------------------------------------------------------------------
def mc_gate(a, r, label, signal=0.15, min_bins=0.0, total_sigma=4.0, total_rel=0.01):
    """a[Ba, ...], r[Br, ...]: the same output from Ba engine batches and Br reference batches (>= 16 each).
    Contract: z-scores per bin with sigma^2 = sem_gpu^2 + sem_ref^2; |z| < 3 for 99.7 % of the bins with signal and no
    systematic offset, totals consistent.  The 99.7 % is the Gaussian figure: with the sems estimated from B batches, z
    follows Student's t with the Welch-Satterthwaite degrees of freedom of the bin (2B-2 when both sides are equally noisy,
    B-1 when one side dominates), so a bin lies beyond 3 sigma with probability p_i = 2 sf_t(3, dof_i) (0.5 % ... 0.9 % at
    B = 16) and E = sum p_i such bins are expected.  Bins are not independent (the packets of a batch cross many cells), which
    widens the scatter of that count beyond the binomial: allowed are E + 6 sqrt(E) + 3.  No bin may lie beyond the value that the
    largest of N Student-t deviates exceeds once in ten thousand trials (5.5 sigma at least).
    A systematic offset is a mean z beyond max(0.15, 3.5/sqrt(N), 4 standard errors of that mean as estimated from the batches).
    The batch totals must agree within the Student-t equivalent of total_sigma Gaussian sigmas (Welch) and total_rel relative --
    the test with power against a bias, since the bins of a total add coherently.
    False alarms: the thresholds are set so that a full GPU run (~120 gates, the reference's threads drawing different streams
    every time) fails by chance in about 1 % of the runs; at the 3.5 sigma / 3.5/sqrt(N) settings of round 1's contract three of
    four consecutive full runs of this round failed in one gate or another, each time a different one."""
    from scipy import stats
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1); r = np.asarray(r, dtype=np.float64).reshape(len(r), -1)
    scale = float(np.max(np.abs(r))) or 1.0          # (the squares of very small luminosities would underflow)
    a = a / scale; r = r / scale
    Ba, Br = len(a), len(r)
    assert Ba >= 16 and Br >= 16, f"{label}: the gate needs at least 16 batches on both sides ({Ba}, {Br})"
    ta, tr = a.sum(1), r.sum(1)
    wa, wr = ta.var(ddof=1) / Ba, tr.var(ddof=1) / Br
    st = np.sqrt(wa + wr)
    zt = (ta.mean() - tr.mean()) / st
    # zt follows Student's t (Welch-Satterthwaite degrees of freedom, ~30 for 16 + 16 batches), whose tails are wider than a
    # Gaussian's: the limit is the t quantile with the two-sided tail probability of `total_sigma` Gaussian sigmas (4 sigma =
    # 6.3e-5: the ~120 gates of a full GPU run then raise a false alarm in about 1 % of the runs; 3.5 sigma would do so in 6 %)
    dof_t = (wa + wr) ** 2 / max(wa ** 2 / (Ba - 1) + wr ** 2 / (Br - 1), 1e-300)
    total_limit = float(stats.t.isf(stats.norm.sf(total_sigma), dof_t))
    # the relative bound applies as far as the statistics of the two batch sets resolve it (small runs; the reference's threads draw
    # different streams from run to run): it never binds tighter than the sigma gate
    rel_bound = None if total_rel is None else max(total_rel, total_limit * st / abs(tr.mean()))
    assert abs(zt) < total_limit and (rel_bound is None or abs(ta.mean() / tr.mean() - 1) < rel_bound), \
        f"{label}: totals differ by {zt:.2f} sigma (gpu {ta.mean():.6g}, reference {tr.mean():.6g})"
    if a.shape[1] < 2:
        return dict(zt=zt)
    ma, mr = a.mean(0), r.mean(0); sa, sr_ = a.std(0, ddof=1) / np.sqrt(Ba), r.std(0, ddof=1) / np.sqrt(Br)
    # bins with signal: batch mean known to better than 15 % on both sides (sparser bins have skewed, far-from-Gaussian batch statistics)
    ok = (sa > 0) & (sr_ > 0) & (sa < signal * np.abs(ma)) & (sr_ < signal * np.abs(mr))       # (Stokes Q, U, V are signed)
    N = int(ok.sum())
    assert N >= max(1, min_bins * a.shape[1]), f"{label}: only {N} of {a.shape[1]} bins carry signal"
    h = np.hypot(sa[ok], sr_[ok])                    # (no explicit squares: bins ~1e-170 x the brightest one would underflow)
    z = (ma[ok] - mr[ok]) / h
    fa, fr = (sa[ok] / h) ** 2, (sr_[ok] / h) ** 2
    dof = 1.0 / (fa ** 2 / (Ba - 1) + fr ** 2 / (Br - 1))
    E = float(np.sum(2 * stats.t.sf(3.0, dof)))
    allowed = int(np.ceil(E + 6 * np.sqrt(E) + 3))      # (generous: the bins of a batch are correlated, outliers come in groups)
    nout = int(np.sum(np.abs(z) >= 3))
    assert nout <= allowed, f"{label}: {nout} of {N} bins beyond 3 sigma ({E:.1f} expected, {allowed} allowed)"
    zmax = max(5.5, float(stats.t.isf(0.5e-4 / N, float(np.min(dof)))))
    assert np.max(np.abs(z)) < zmax, f"{label}: a bin differs by {np.max(np.abs(z)):.1f} sigma (limit {zmax:.1f} for {N} bins)"
    # no systematic offset: the mean z over the bins.  Its standard error is NOT 1/sqrt(N) -- the bins of a batch are correlated (its
    # packets cross many cells) -- so it is estimated from the batches themselves: mean z = mean_b(u_b) - mean_b(v_b) with
    # u_b = mean over bins of a_b / h (engine batches) and v_b likewise for the reference batches
    u, v = (a[:, ok] / h).mean(1), (r[:, ok] / h).mean(1)
    se = float(np.sqrt(u.var(ddof=1) / Ba + v.var(ddof=1) / Br))
    lim = max(0.15, 3.5 / np.sqrt(N), 4.0 * se)
    assert abs(z.mean()) < lim, f"{label}: systematic offset, mean z = {z.mean():.3f} over {N} bins (limit {lim:.3f}, standard error {se:.3f})"
    return dict(zt=zt, z=z, bins=N, outliers=nout)


def sym_rays(R=18000 * PC, n=40000, seed=911):
    """rays for the grids with symmetries: isotropic rays from inside and outside, plus rays from the origin, through the centre,
    inside and parallel to the equatorial plane, along and parallel to the z axis, and from far outside"""
    r, k = rays(n, [-R, R, -R, R, -R, R], seed, scale=1.3)
    rng = np.random.default_rng(12)
    r2 = (rng.random((3000, 3)) - 0.5) * 2 * R; k2 = rng.normal(size=(3000, 3))
    r2[:400] = 0.0
    k2[400:800] = -r2[400:800]
    r2[800:1200, 2] = 0.0; k2[800:1200, 2] = 0.0
    r2[1200:1600, :2] = 0.0; k2[1200:1600, :2] = 0.0
    k2[1600:2000, 2] = 0.0
    k2[2000:2400, :2] = 0.0
    r2[2400:2700] *= 3.0
    k2 /= np.linalg.norm(k2, axis=1, keepdims=True)
    return np.concatenate([r, r2]), np.concatenate([k, k2])
