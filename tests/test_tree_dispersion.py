"""TreeDustGrid's density-dispersion subdivision criterion (TreeDustGrid.cpp:63,192,215-221 with
TreeNodeSampleDensityCalculator::densityDispersion, TreeNodeSampleDensityCalculator.cpp:62-67).

CPU part: the decision logic of the host mirrors with a stand-in for the device sampling, and the property validation of
both hosts.  GPU part (skg_sample_boxes_dispersion through the C ABI): the masses are bit-identical to skg_sample_boxes,
the sampled dispersion of a box lies between 0 and the exact dispersion of the density over the box and approaches it with
the sample count, and a tree grown on the criterion subdivides exactly the nodes the criterion allows."""
import os
import subprocess

import numpy as np
import pytest

import common
from skirt_b200 import hostlib, simulation as sim

PC = common.PC
HR, HZ = 4000 * PC, 140 * PC
# with the levels used below the deepest judged nodes have an exact dispersion of 0.927-0.938: their samples decide
FRAC = 0.92
RUN = os.path.join(common.ROOT, "skirt_b200", "skirt_b200_run")
needs_host = pytest.mark.skipif(not hostlib.lib_available(), reason="skirt_b200/libskirthost.so not built")


def exact_dispersion(boxes, hR=HR, hz=HZ):
    """(max - min) / max of an untruncated exponential disk over each box: the density falls monotonically with R and |z|"""
    b = np.asarray(boxes, float).reshape(-1, 6)

    def nearest_farthest(lo, hi):
        near = np.where((lo <= 0) & (hi >= 0), 0.0, np.minimum(np.abs(lo), np.abs(hi)))
        return near, np.maximum(np.abs(lo), np.abs(hi))
    xn, xf = nearest_farthest(b[:, 0], b[:, 3]); yn, yf = nearest_farthest(b[:, 1], b[:, 4]); zn, zf = nearest_farthest(b[:, 2], b[:, 5])
    return -np.expm1(-(np.hypot(xf, yf) - np.hypot(xn, yn)) / hR - (zf - zn) / hz)


class FakeEngine:
    """sample_boxes with numpy's generator instead of the device: the same quantities, for the host logic on CPU"""
    def __init__(self, geometry):
        self.geometry = geometry; self.calls = []

    def sample_boxes(self, boxes, geometries, norm, sample_count=100, seed=4357, dispersion=False):
        b = np.asarray(boxes, float).reshape(-1, 6); rng = np.random.default_rng(seed)
        u = rng.random((len(b), sample_count, 3))
        p = b[:, None, :3] + u * (b[:, None, 3:] - b[:, None, :3])
        rho = norm[0] * self.geometry.density(p[..., 0], p[..., 1], p[..., 2])
        mass = rho.mean(axis=1) * np.prod(b[:, 3:] - b[:, :3], axis=1)
        self.calls.append(dispersion)
        if not dispersion:
            return mass
        mx, mn = rho.max(axis=1), rho.min(axis=1)
        return mass, np.where(mx > 0, (mx - mn) / np.where(mx > 0, mx, 1), 0.0)


def check_tree_against_criterion(t, frac, minlevel, maxlevel):
    box = t["box"].reshape(-1, 6); inner = t["child0"] >= 0; level = t["level"]
    ex = exact_dispersion(box)
    # a sampled dispersion never exceeds the exact one: a node above the forced levels was split only if the criterion allows it
    judged = inner & (level >= minlevel)
    assert judged.any() and (ex[judged] >= frac).all()
    assert level[~inner].max() <= maxlevel and level[~inner].min() >= minlevel
    # and the criterion did stop somewhere: leaves above the deepest level exist, where the disk is smooth
    assert (level[~inner] < maxlevel).any() and (level[~inner] == maxlevel).any()


@needs_host
@pytest.mark.parametrize("cls", [sim.OctTreeDustGrid, sim.BinTreeDustGrid])
def test_dispersion_criterion_in_the_python_mirror(cls):
    b = common.C1_BOX; geo = sim.ExpDiskGeometry(HR, HZ)
    lo, hi = (2, 6) if cls is sim.OctTreeDustGrid else (6, 16)
    fake = FakeEngine(geo)
    g = cls(b[0], b[1], b[2], b[3], b[4], b[5], lo, hi, "Neighbor", 200, 0.0, 0.0, FRAC).build(fake, [geo.sampler()], [1.0])
    assert all(fake.calls) and len(fake.calls) == hi - lo - 1
    check_tree_against_criterion(g.tables(), FRAC, lo, hi)
    # without any criterion the tree is complete (TreeDustGrid.cpp:192), with the mass criterion alone the sampler is not asked for dispersions
    fake2 = FakeEngine(geo)
    full = cls(b[0], b[1], b[2], b[3], b[4], b[5], 1, 3, "Neighbor", 10, 0.0, 0.0, 0.0).build(fake2, [geo.sampler()], [1.0])
    assert full.numCells() == (8 if cls is sim.OctTreeDustGrid else 2) ** 3 and not fake2.calls
    fake3 = FakeEngine(geo)
    cls(b[0], b[1], b[2], b[3], b[4], b[5], 2, 4, "Neighbor", 50, 0.0, 1e-3, 0.0).build(fake3, [geo.sampler()], [1.0])
    assert fake3.calls and not any(fake3.calls)
    # either criterion asks for a split: adding the dispersion criterion to the mass criterion never gives a coarser tree
    both = cls(b[0], b[1], b[2], b[3], b[4], b[5], lo, hi, "Neighbor", 200, 0.0, 1e-3, FRAC).build(FakeEngine(geo), [geo.sampler()], [1.0])
    assert both.numCells() >= g.numCells()


def test_dispersion_fraction_validation(tmp_path):
    b = common.C1_BOX
    with pytest.raises(sim.FatalError, match="maximum density dispersion fraction should be positive"):
        sim.OctTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], maxDensDispFraction=-0.1)
    with pytest.raises(sim.FatalError, match="maximum mass fraction should be positive"):
        sim.OctTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], maxMassFraction=-1e-6)
    with pytest.raises(sim.FatalError, match="maximum mean optical depth should be positive"):
        sim.BinTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], maxOpticalDepth=-1.0)


def _run_text(grid_line, packages=1e5):
    return "\n".join(["sim oligo", f"packages {packages!r}", "seed 4357", "wavelengths 0.55e-6", common.box_line(common.C1_BOX), grid_line,
                      "storeabs 1", f"dustmix table {common.MIX_V['kabs']!r} {common.MIX_V['ksca']!r} {common.MIX_V['g']!r}",
                      f"dust 1.0 0.55e-6 expdisk {HR!r} {HZ!r} 0 0", f"stellar 1.0 expdisk {HR!r} {350*PC!r} 0 0",
                      f"instrument sed s88 {1e7*PC!r} {float(np.radians(88))!r} 0 0"]) + "\n"


# ---- on the device ----------------------------------------------------------------------------------------------------
def _boxes(rng, n):
    b = common.C1_BOX
    lo = np.stack([rng.uniform(b[0], b[1], n), rng.uniform(b[2], b[3], n), rng.uniform(b[4], b[5], n)], axis=1)
    w = np.stack([rng.uniform(20, 3000, n), rng.uniform(20, 3000, n), rng.uniform(5, 400, n)], axis=1) * PC
    return np.concatenate([lo, lo + w], axis=1)


@pytest.mark.gpu
def test_box_dispersion_on_the_device(engine):
    geo = sim.ExpDiskGeometry(HR, HZ); g = geo.sampler()
    boxes = _boxes(np.random.default_rng(11), 5000)
    ex = exact_dispersion(boxes)
    mass = engine.sample_boxes(boxes, [g], [1.0], 100, seed=77)
    m100, d100 = engine.sample_boxes(boxes, [g], [1.0], 100, seed=77, dispersion=True)
    assert np.array_equal(mass, m100)                                   # the same samples, the same sum
    assert (d100 >= 0).all() and (d100 <= ex * (1 + 1e-12) + 1e-15).all()
    m2, d2 = engine.sample_boxes(boxes, [g], [1.0], 100, seed=77, dispersion=True)
    assert np.array_equal(d2, d100) and np.array_equal(m2, m100)        # reproducible
    _, d4000 = engine.sample_boxes(boxes, [g], [1.0], 4000, seed=78, dispersion=True)
    assert (d4000 <= ex * (1 + 1e-12) + 1e-15).all()
    # more samples reach further into the corners: the estimate grows towards the exact value
    assert np.median(d4000 / ex) > np.median(d100 / ex) > 0.45 and np.median(d4000 / ex) > 0.8
    # masses against the host mirror's density on a fine midpoint lattice, on a few boxes
    for q in range(0, 5000, 1250):
        bx = boxes[q]; n = 40
        ax = [bx[i] + (np.arange(n) + 0.5) / n * (bx[i + 3] - bx[i]) for i in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing="ij")
        want = geo.density(X, Y, Z).mean() * np.prod(bx[3:] - bx[:3])
        m4000 = engine.sample_boxes(bx[None], [g], [1.0], 4000, seed=5)[0]
        assert abs(m4000 / want - 1) < 0.1
    # a truncated disk: boxes wholly outside the truncation radius hold no dust and have no dispersion (maxrho == 0)
    cut = sim.ExpDiskGeometry(HR, HZ, 10000 * PC, 0.0).sampler()
    far = boxes[np.hypot(np.minimum(np.abs(boxes[:, 0]), np.abs(boxes[:, 3])), np.minimum(np.abs(boxes[:, 1]), np.abs(boxes[:, 4]))) > 10500 * PC]
    far = far[(far[:, 0] * far[:, 3] > 0) & (far[:, 1] * far[:, 4] > 0)]
    assert len(far) > 100
    mf, df = engine.sample_boxes(far, [cut], [1.0], 50, seed=3, dispersion=True)
    assert not mf.any() and not df.any()
    # two components: the dispersion is that of the summed density -- a constant-ratio second copy changes nothing
    _, dd = engine.sample_boxes(boxes, [g, g], [1.0, 3.0], 100, seed=77, dispersion=True)
    assert np.allclose(dd, d100, rtol=1e-12, atol=1e-15)
    with pytest.raises(Exception, match="Number of random samples must be at least 1"):
        engine.sample_boxes(boxes[:4], [g], [1.0], 0, seed=1, dispersion=True)


@pytest.mark.gpu
@needs_host
@pytest.mark.parametrize("kind", ["octtree", "bintree"])
def test_tree_grown_on_the_dispersion_criterion(tmp_path, engine, kind):
    b = common.C1_BOX; geo = sim.ExpDiskGeometry(HR, HZ)
    cls = sim.OctTreeDustGrid if kind == "octtree" else sim.BinTreeDustGrid
    lo, hi = (2, 6) if kind == "octtree" else (6, 16)
    grid = cls(b[0], b[1], b[2], b[3], b[4], b[5], lo, hi, "Neighbor", 100, 0.0, 0.0, FRAC).build(engine, [geo.sampler()], [1.0])
    check_tree_against_criterion(grid.tables(), FRAC, lo, hi)
    if not os.path.exists(RUN):
        pytest.skip("skirt_b200_run not built")
    # the C++ host grows its own tree on the same criterion (its own sampling streams: the trees are statistically alike)
    f = tmp_path / "sim.txt"; f.write_text(_run_text(f"grid {kind} {lo} {hi} 1 0 100 0 {FRAC}"))
    r = subprocess.run([RUN, str(f), str(tmp_path / "out")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    import json
    st = json.loads(r.stdout.strip().splitlines()[-1])
    assert abs(st["cells"] / grid.numCells() - 1) < 0.15


@pytest.mark.gpu
@pytest.mark.parametrize("kind,search,maxlevel", [("octtree", 0, 5), ("octtree", 1, 5), ("octtree", 2, 5), ("bintree", 0, 12), ("bintree", 1, 12)])
def test_device_walkers_follow_barycentric_trees(engine, kind, search, maxlevel):
    """trees grown by the reference with OctTreeDustGrid::barycentric / BinTreeDustGrid's Barycenter direction method (nodes split
    off-centre / across non-alternating axes): the table-driven device walkers follow them bit for bit with every search method
    (tools/check_barycentric.py is the same check as a script; its B200 output is profiles/r02_z_barycentric_gpu.txt)"""
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    spec = common.spec_grid(kind, search=search, maxlevel=maxlevel).replace(" 0 50\n", " 1 50\n")
    S = sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v()).setup()
    t, med = S.grid_tables(), S.medium()
    r, k = common.rays(20000, common.C1_BOX, 31)
    ref = S.path_batch(r, k, ell=0, nthreads=os.cpu_count() or 1)
    engine.set_grid(t); engine.medium(med["rho"], med["kext"], med["ksca"], med["g"])
    assert common.paths_bit_identical(engine.path_batch(r, k, ell=0), ref)
    assert np.array_equal(engine.whichcell(r[:5000]), S.whichcell(r[:5000]))


def test_barycentric_subdivision_is_refused_by_the_hosts():
    """the host mirrors grow centre-split trees only: the option is an error, not a silent fallback"""
    b = common.C1_BOX
    with pytest.raises(sim.FatalError, match="barycentric subdivision is not supported"):
        sim.OctTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], barycentric=True)
    with pytest.raises(sim.FatalError, match="barycentric subdivision is not supported"):
        sim.BinTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], directionMethod="Barycenter")
    sim.BinTreeDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], directionMethod="Alternating")
