"""Generates the golden vectors under tests/golden/ from the reference's OWN code (oracle/_ref/libskirtref.so,
built in place from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Every fixture holds the flattened grid tables and medium exactly as the reference built them (so that the
engine and the restated oracle consume the same state), seeded rays, and the reference's DustGrid::path() +
fillOpticalDepth() records, whichcell() and opticaldepth() answers for them.  The Monte Carlo fixture holds the
batch mean and standard error of the reference's detector arrays and absorption table for a small C1 run."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                   # noqa: E402
from oracle import skirtref as sr               # noqa: E402

NRAYS = 384


def geometry_case(name, S, seed):
    S.setup()
    t = S.grid_tables(); med = S.medium()
    axes = (t["xv"], t["yv"], t["zv"]) if t["kind"] == "cartesian" else None
    r, k = common.rays(NRAYS, common.C1_BOX, seed)
    ra, ka = common.adversarial_rays(common.C1_BOX, axes)
    r = np.concatenate([r, ra]); k = np.concatenate([k, ka])
    p = S.path_batch(r, k, ell=0)
    rng = np.random.default_rng(seed + 1)
    dist = rng.random(len(r)) * 30000 * common.PC
    out = dict(r=r, k=k, distance=dist, whichcell=S.whichcell(r), tau_inf=S.opticaldepth_batch(r, k, 0),
               tau_dist=S.opticaldepth_batch(r, k, 0, dist), warnings=np.array([sr.lib().skr_warnings()]))
    for key, v in p.items():
        out["path_" + key] = v
    for key, v in t.items():
        out["grid_" + key] = np.asarray(v)
    for key, v in med.items():
        out["med_" + key] = v
    path = os.path.join(HERE, f"geom_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name:14s} cells {S.Ncells:6d} rays {len(r)} segments {p['offsets'][-1]:7d} -> {os.path.getsize(path)/1024:.0f} KiB")


def mc_case(batches=16):
    cfg = common.cfg_c1(n=24, packages=2e5, threads=os.cpu_count() or 1, storeabs=1)
    S = common.make_ref(cfg).setup()
    t = S.grid_tables(); med = S.medium()
    fr, se, la = [], [], []
    for b in range(batches):
        S.reset(9000 + 13 * b); S.run_stellar(); ins = S.instruments()
        fr.append(ins[0]["frame"].copy()); se.append(ins[1]["sed"].copy()); la.append(S.labs().ravel().copy())
    out = dict(Npp=np.array([S.packages_per_lambda()]), batches=np.array([batches]), L=S.luminosities())
    for nm, a in (("frame", fr), ("sed", se), ("labs", la)):
        a = np.array(a)
        out[nm + "_mean"] = a.mean(0); out[nm + "_sem"] = a.std(0, ddof=1) / np.sqrt(len(a))
        tot = a.reshape(len(a), -1).sum(1)
        out[nm + "_total_mean"] = np.array([tot.mean()]); out[nm + "_total_sem"] = np.array([tot.std(ddof=1) / np.sqrt(len(a))])
    for key, v in t.items():
        out["grid_" + key] = np.asarray(v)
    for key, v in med.items():
        out["med_" + key] = v
    path = os.path.join(HERE, "mc_c1.npz")
    np.savez_compressed(path, **out)
    print(f"mc_c1: {batches} batches of {S.packages_per_lambda():g} packets -> {os.path.getsize(path)/1024:.0f} KiB")


def sampler_case():
    """launch samplers: moments of the reference's StellarSystem::launch for the geometries the engine supports"""
    from skirt_b200 import configs
    from oracle import refspec
    p = configs.c2_params(n=4, nlambda=8, packages=10)
    spec, L, mixes = refspec.reference_spec(p, threads=1, dustsamples=1)
    S = sr.RefSim(spec, luminosities=L, mixes=mixes).setup()
    out = {}
    S.reset(4357); out["uniforms_4357"] = S.uniforms(2000)      # raw MT19937 stream of thread 0 (Random.cpp:89-126)
    for ell in (0, 4):
        r, k, Lw = S.sample_launch(ell, 200000)
        out[f"r_mean_{ell}"] = r.mean(0); out[f"r_absmean_{ell}"] = np.abs(r).mean(0); out[f"r_sq_{ell}"] = (r * r).mean(0)
        out[f"k_mean_{ell}"] = k.mean(0); out[f"L_mean_{ell}"] = np.array([Lw.mean()]); out[f"L_sq_{ell}"] = np.array([(Lw * Lw).mean()])
        out[f"R_quant_{ell}"] = np.quantile(np.hypot(r[:, 0], r[:, 1]), [0.1, 0.25, 0.5, 0.75, 0.9])
        out[f"z_quant_{ell}"] = np.quantile(np.abs(r[:, 2]), [0.1, 0.25, 0.5, 0.75, 0.9])
    out["L"] = np.array(L)
    # an exponential disk with two spiral arms (C3's stellar geometry)
    spiral = dict(arms=2, pitch=float(np.radians(20)), radius=4000 * common.PC, phase=0.0, weight=1.0, index=1)
    cfg = common.cfg_c1(n=4, packages=10)
    cfg["sources"][0]["spiral"] = spiral
    Ssp = common.make_ref(cfg).setup()
    r, k, Lw = Ssp.sample_launch(0, 400000)
    phi = np.arctan2(r[:, 1], r[:, 0]); R = np.hypot(r[:, 0], r[:, 1])
    out["spiral_R_quant"] = np.quantile(R, [0.1, 0.25, 0.5, 0.75, 0.9])
    ring = (R > 3000 * common.PC) & (R < 5000 * common.PC)
    out["spiral_phi_hist"] = np.histogram(phi[ring], bins=24, range=(-np.pi, np.pi))[0] / ring.sum()
    out["spiral_n"] = np.array([ring.sum()])
    np.savez_compressed(os.path.join(HERE, "launch_c2.npz"), **out)
    print("launch_c2 written")


if __name__ == "__main__":
    mk = lambda spec, **kw: sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw)
    geometry_case("cart_lin", mk(common.spec_c1(n=20)), 101)
    geometry_case("cart_sympow", mk(common.spec_c1(n=16, mesh="sympow 30")), 102)
    for s in (0, 1, 2):
        geometry_case(f"octtree_s{s}", mk(common.spec_grid("octtree", search=s, maxlevel=4)), 103)
    for s in (0, 1):
        geometry_case(f"bintree_s{s}", mk(common.spec_grid("bintree", search=s, maxlevel=10)), 104)
    geometry_case("amesh", mk(common.spec_grid("amesh"), amesh=common.make_amesh(max_depth=3)), 105)
    geometry_case("voronoi", mk(common.spec_grid("voronoi"), particles=common.voronoi_particles(1500)), 106)
    mc_case()
    sampler_case()
