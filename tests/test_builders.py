"""Product-side grid builders (skirt_b200/libskirthost.so: include/skirthost.h, skirt_b200/host/GridBuilders.cpp) against
the reference's own grid classes (oracle/_ref): for the same subdivision decisions / mesh file / particles the flattened
tables -- node boxes and ids, cell numbers, neighbour lists IN THE REFERENCE'S ORDER, wall neighbours, Voro++ neighbour
lists, block lists and search trees -- are identical.  CPU only."""
import os

import numpy as np
import pytest

import common
from skirt_b200 import configs, hostlib

pytestmark = pytest.mark.skipif(not hostlib.lib_available(), reason="skirt_b200/libskirthost.so not built")


def _ref(spec, **kw):
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    return sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw).setup()


@pytest.mark.parametrize("kind,kw", [("octtree", dict(minlevel=2, maxlevel=6, massfrac=2e-5)),
                                     ("bintree", dict(minlevel=4, maxlevel=14, massfrac=2e-5))])
@pytest.mark.parametrize("search", [0, 1])
def test_tree_builder_reproduces_the_reference_tree(kind, kw, search):
    ref = _ref(common.spec_grid(kind, search=search, **kw)).grid_tables()
    tb = hostlib.TreeBuilder(0 if kind == "octtree" else 1, common.C1_BOX, kw["minlevel"], kw["maxlevel"])
    child0 = ref["child0"]; pos = {tuple(b): i for i, b in enumerate(ref["box"].reshape(-1, 6))}
    levels = []

    def decide(level, boxes):       # the reference's own decisions: a node was subdivided iff it has children
        levels.append(level)
        return np.array([child0[pos[tuple(b)]] >= 0 for b in boxes])
    tb.grow(decide)
    mine = tb.finish(search)
    assert min(levels) == kw["minlevel"] + 1 and max(levels) == kw["maxlevel"] - 1
    keys = ["box", "child0", "parent", "cell"] + (["nbrStart", "nbrIds"] if search == 1 else [])
    for k in keys:
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{kind}: {k} differs from the reference"
    inner = child0 >= 0         # BinTreeNode::_dir is only ever set for nodes that were split
    assert np.array_equal(mine["dir"][inner], np.asarray(ref["dir"])[inner])
    assert mine["Ncells"] == int((child0 < 0).sum())


@pytest.mark.parametrize("tt,extra", [("oct", 0), ("oct", 1), ("bin", 0), ("bin", 1)])
def test_particle_tree_builder_reproduces_the_reference_tree(tt, extra):
    """ParticleTreeDustGrid: node ids follow the order in which the particles are added (ParticleTreeDustGrid.cpp:36-72)"""
    pts = common.voronoi_particles(3000, seed=5)
    pts[7] = [9e30, 0, 0]                        # a particle outside the box is skipped (whichnode returns 0)
    ref = _ref(common.spec_grid("particletree_" + tt, maxlevel=extra), particles=pts).grid_tables()
    mine = hostlib.build_particle_tree(0 if tt == "oct" else 1, common.C1_BOX, pts, extra)
    assert ref["search"] == 3 and mine["search"] == 3
    for k in ("box", "child0", "parent", "cell"):
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{tt}: {k} differs from the reference"
    inner = ref["child0"] >= 0
    assert np.array_equal(mine["dir"][inner], np.asarray(ref["dir"])[inner])
    assert mine["Ncells"] == int((ref["child0"] < 0).sum()) >= 2999
    with pytest.raises(hostlib.HostError, match="share a position"):
        hostlib.build_particle_tree(0, common.C1_BOX, np.zeros((2, 3)), 0)


@pytest.mark.parametrize("kind", list(common.SYM_GRIDS))
def test_symmetric_grid_mirrors_reproduce_the_reference_borders(kind):
    """Sphere1D / Sphere2D / Cylinder2D: the border arrays (and the polar border the reference inserts at pi/2) are identical"""
    ref = _ref(common.spec_grid(kind)).grid_tables()
    mine = common.sym_grid_mirror(kind).tables()
    assert mine["kind"] == ref["kind"]
    for k in ref:
        if k != "kind":
            assert np.array_equal(np.asarray(mine[k]), np.asarray(ref[k])), f"{kind}: {k} differs from the reference"
    if kind == "sphere2d_odd":
        assert len(mine["thetav"]) == 11 and mine["cv"][5] == 0.0


def test_tree_builder_validation():
    with pytest.raises(hostlib.HostError, match="Maximum tree level should be larger"):
        hostlib.TreeBuilder(0, common.C1_BOX, 3, 3)
    with pytest.raises(hostlib.HostError, match="maximum tree level should be at least 2"):
        hostlib.TreeBuilder(0, common.C1_BOX, 0, 1)
    tb = hostlib.TreeBuilder(1, common.C1_BOX, 1, 3)
    tb.grow(lambda level, boxes: np.ones(len(boxes), bool))
    with pytest.raises(hostlib.HostError, match="Bookkeeping method is not compatible with binary tree"):
        tb.finish(2)
    t = tb.finish(0)
    assert t["Ncells"] == 8 and len(t["child0"]) == 15


def test_adaptive_mesh_builder_reproduces_the_reference_mesh():
    am = common.make_amesh(root=(4, 4, 4), max_depth=4, frac=2e-3)
    S = _ref(common.spec_grid("amesh"), amesh=am)
    ref = S.grid_tables()
    mine = hostlib.build_adaptive_mesh(common.C1_BOX, am[0])
    for k in ("box", "nxyz", "child0", "cell", "wallNbr"):
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{k} differs from the reference"
    assert np.array_equal(mine["volume"], S.volumes())
    # densities: the reference's table is value x units in cell order
    np.testing.assert_allclose(am[1][mine["fileIndex"]] * 1e-24, S.medium()["rho"].ravel(), rtol=1e-15)
    # the vectorised generator of the product side writes the same file order as the recursive one of the tests
    nx, val = configs.synthetic_amesh(root=4, depth=4, frac=2e-3)
    assert np.array_equal(nx, am[0]) and np.allclose(val, am[1], rtol=1e-14)
    with pytest.raises(hostlib.HostError, match="Reached end of file"):
        hostlib.build_adaptive_mesh(common.C1_BOX, am[0][:-3])
    with pytest.raises(hostlib.HostError, match="Superfluous data"):
        hostlib.build_adaptive_mesh(common.C1_BOX, np.concatenate([am[0], [[0, 0, 0]]]))


def test_voronoi_builder_reproduces_the_reference_mesh():
    if not hostlib.voronoi_available():
        pytest.skip("libskirthost.so was built without Voro++")
    pts = common.voronoi_particles(3000)
    S = _ref(common.spec_grid("voronoi"), particles=pts)
    ref = S.grid_tables()
    mine = hostlib.build_voronoi_mesh(common.C1_BOX, pts)
    assert mine["nb"] == ref["nb"]
    for k in ("particles", "nbrStart", "nbrIds", "blkStart", "blkIds", "blkTree", "kdM", "kdAxis", "kdUp", "kdLeft", "kdRight", "cellBox"):
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{k} differs from the reference"
    assert np.array_equal(mine["volume"], S.volumes())
    assert abs(mine["volume"].sum() / np.prod(common.C1_BOX[1::2] - common.C1_BOX[0::2]) - 1) < 1e-9


def test_host_library_exports_its_header():
    import ctypes, os, re
    hdr = open(os.path.join(common.ROOT, "include", "skirthost.h")).read()
    names = sorted(set(re.findall(r"\b(skh_[a-z_]+)\s*\(", hdr)))
    L = ctypes.CDLL(hostlib.LIB_PATH)
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in skirthost.h but not exported"


def test_two_phase_grid_weights():
    """TwoPhaseDustGrid (TwoPhaseDustGrid.cpp:18-39): two weights, the high one in a fraction `fillingFactor` of the cells, volume mean 1;
    the density table of the dust system carries them (DustSystem.cpp:165-176)"""
    from skirt_b200 import simulation as sim
    PC = common.PC
    ext = [-1e4 * PC, 1e4 * PC] * 3
    g = sim.TwoPhaseDustGrid(*ext, sim.LinMesh(20), sim.LinMesh(20), sim.LinMesh(20), fillingFactor=0.2, contrast=50.0)
    w = g.weights(); hi, lo = 50.0 / (50 * 0.2 + 0.8), 1.0 / (50 * 0.2 + 0.8)
    assert set(np.unique(w)) == {hi, lo} and abs((w == hi).mean() - 0.2) < 0.02 and abs(w.mean() - 1) < 0.1
    plain = sim.CartesianDustGrid(*ext, sim.LinMesh(20), sim.LinMesh(20), sim.LinMesh(20))
    lg = sim.OligoWavelengthGrid([0.55e-6]); mix = sim.TableDustMix(common.MIX_V["kabs"], common.MIX_V["ksca"], common.MIX_V["g"])
    comp = sim.DustComp(sim.ExpDiskGeometry(4000 * PC, 140 * PC), mix, 1.0, 0.55e-6)
    a, b = sim.DustSystem(g, [comp], lg).rho, sim.DustSystem(plain, [comp], lg).rho
    np.testing.assert_allclose(a[:, 0], b[:, 0] * w, rtol=1e-15)
    with pytest.raises(sim.FatalError, match="filling factor"):
        sim.TwoPhaseDustGrid(*ext, sim.LinMesh(4), sim.LinMesh(4), sim.LinMesh(4), fillingFactor=1.0, contrast=2.0)


def _logmesh_golden():
    """tests/golden/logmesh.txt: borders printed by the reference's own NR::zerologgrid (make_logmesh_golden.sh), hex floats"""
    path = os.path.join(os.path.dirname(__file__), "golden", "logmesh.txt")
    for line in open(path):
        t = line.split()
        yield int(t[0]), float.fromhex(t[1]), np.array([float.fromhex(x) for x in t[2:]])


def test_logmesh_mirror_reproduces_the_reference_borders():
    """LogMesh (LogMesh.cpp:47-53 -> NR::zerologgrid): the Python mirror, bit for bit, and its parameter check"""
    from skirt_b200 import simulation as sim
    cases = list(_logmesh_golden())
    assert len(cases) == 20
    for n, tc, want in cases:
        got = sim.LogMesh(n, tc).mesh()
        assert got.shape == (n + 1,) and np.array_equal(got, want), (n, tc)
        assert got[0] == 0.0 and got[1] == pytest.approx(tc) and got[-1] == pytest.approx(1.0)
    assert np.array_equal(sim.LogMesh(1, 0.5).mesh(), [0.0, 1.0])
    with pytest.raises(sim.FatalError):
        sim.LogMesh(10, 1.0)
    # a radial mesh of a spherical grid
    g = sim.Sphere1DDustGrid(18000 * common.PC, sim.LogMesh(30, 1e-3))
    rv = g.tables()["rv"]
    assert len(rv) == 31 and rv[0] == 0.0 and np.all(np.diff(rv) > 0) and rv[-1] == pytest.approx(18000 * common.PC)


def test_logmesh_cpp_mirror_reproduces_the_reference_borders(tmp_path):
    """the C++ host layer's LogMesh (skirt_b200/host/SimulationItems.hpp) against the same golden borders"""
    import shutil, subprocess
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no C++ compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "m.cpp"
    src.write_text('#include <cstdio>\n#include <cstdlib>\n#include "SimulationItems.hpp"\n'
                   'int main(int argc, char** argv) { skirt::LogMesh m; m.setNumBins(std::atoi(argv[1])); m.setCentralBinFraction(std::strtod(argv[2], 0));\n'
                   '  for (double v : m.mesh()) std::printf("%a\\n", v); return 0; }\n')
    exe = tmp_path / "m"
    subprocess.run([cxx, "-std=c++17", "-O2", f"-I{root}/skirt_b200/host", f"-I{root}/include", str(src), "-o", str(exe)], check=True)
    for n, tc, want in list(_logmesh_golden())[::3]:
        out = subprocess.run([str(exe), str(n), tc.hex()], check=True, capture_output=True, text=True).stdout.split()
        assert np.array_equal([float.fromhex(x) for x in out], want), (n, tc)


@pytest.mark.parametrize("kind,kw", [("octtree", dict(minlevel=2, maxlevel=5, massfrac=2e-4)),
                                     ("bintree", dict(minlevel=4, maxlevel=12, massfrac=2e-4))])
def test_tree_builder_reproduces_barycentric_reference_trees(kind, kw):
    """OctTreeDustGrid::barycentric (nodes split at the barycentre of their dust, BaryOctTreeNode.cpp:27-30) and BinTreeDustGrid's
    Barycenter direction method (BaryBinTreeNode.cpp:34-58) through skh_tree_subdivide_at: with the reference's own decisions and
    split points (octree: read off its children; binary tree: a barycentre nearest to a wall along the axis the reference chose)
    the tables and the neighbour lists are the reference's"""
    spec = common.spec_grid(kind, search=1, **kw).replace(" 0 50\n", " 1 50\n")
    ref = _ref(spec).grid_tables()
    child0 = ref["child0"]; rbox = ref["box"].reshape(-1, 6); rdir = np.asarray(ref["dir"])
    pos = {tuple(b): i for i, b in enumerate(rbox)}
    tb = hostlib.TreeBuilder(0 if kind == "octtree" else 1, common.C1_BOX, kw["minlevel"], kw["maxlevel"])

    def decide(level, boxes):
        ids = np.array([pos[tuple(b)] for b in boxes]); flags = child0[ids] >= 0
        bary = 0.5 * (boxes[:, :3] + boxes[:, 3:])
        for q, (l, f) in enumerate(zip(ids, flags)):
            if not f:
                continue
            if kind == "octtree":
                bary[q] = rbox[child0[l], 3:]
            else:
                d = rdir[l]; bary[q, d] = boxes[q, d] + 0.01 * (boxes[q, d + 3] - boxes[q, d])
        return flags, bary
    tb.grow(decide)
    mine = tb.finish(1)
    for k in ("box", "child0", "parent", "cell", "nbrStart", "nbrIds"):
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{kind}: {k} differs from the reference"
    inner = child0 >= 0
    assert np.array_equal(mine["dir"][inner], rdir[inner])
    if kind == "octtree":       # the splits really are off-centre
        mid = 0.5 * (rbox[inner, :3] + rbox[inner, 3:])
        assert (np.abs(rbox[child0[inner], 3:] - mid).max(axis=1) > 1e-6 * np.abs(rbox[0]).max()).mean() > 0.5
    else:                       # and the directions do not simply alternate
        assert (rdir[inner] != mine["level"][inner] % 3).any()


def test_barycentric_direction_rule_ties():
    """BaryBinTreeNode.cpp:44-54: strict comparisons, ties go to the later axis"""
    def first_dir(bary):
        tb = hostlib.TreeBuilder(1, [0., 1., 0., 1., 0., 1.], 0, 2)
        tb.subdivide(None)                                      # level 0 <= minLevel: regular, across x
        level, n, need = tb.frontier()
        assert need and n == 2
        boxes = tb.frontier_boxes()
        b = boxes[:, :3] + np.asarray(bary) * (boxes[:, 3:] - boxes[:, :3])
        tb.subdivide(np.ones(2, bool), b)
        while tb.frontier()[1]:
            tb.subdivide(None)                                  # the last level: nothing is subdivided
        t = tb.finish(0)
        return int(t["dir"][1])
    assert first_dir([0.5, 0.5, 0.5]) == 2          # all equal -> z
    assert first_dir([0.25, 0.5, 0.5]) == 0 and first_dir([0.5, 0.75, 0.5]) == 1 and first_dir([0.5, 0.5, 0.125]) == 2
    assert first_dir([0.25, 0.25, 0.5]) == 1        # dx == dy < dz -> y
    assert first_dir([0.25, 0.5, 0.25]) == 2        # dx == dz < dy -> z


NESTED_CASES = [(1e-7, 1e-3, 12, 5e-6, 4e-5, 9), (9e-8, 2e-3, 50, 1e-6, 3e-6, 40), (1e-7, 1e-3, 2, 2e-7, 3e-7, 2), (1e-7, 1e-3, 30, 1.0000001e-7, 9.99e-4, 3)]


def test_nested_log_wavelength_grid_mirrors(tmp_path):
    """NestedLogWavelengthGrid (NestedLogWavelengthGrid.cpp:21-60): wavelengths and bin widths of the Python and the C++ mirror
    bit for bit against the reference's own class (a pan simulation set up on it), and its property checks"""
    import shutil, subprocess
    from skirt_b200 import simulation as sim
    from oracle import refspec, skirtref as sr
    with pytest.raises(sim.FatalError, match="properly nested"):
        sim.NestedLogWavelengthGrid(1e-7, 1e-3, 10, 1e-8, 1e-5, 5)
    with pytest.raises(sim.FatalError, match="low-resolution grid should be at least 2"):
        sim.NestedLogWavelengthGrid(1e-7, 1e-3, 1, 1e-6, 1e-5, 5)
    with pytest.raises(sim.FatalError, match="high-resolution subgrid should be at least 2"):
        sim.NestedLogWavelengthGrid(1e-7, 1e-3, 10, 1e-6, 1e-5, 1)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = None
    if cxx:
        src = tmp_path / "w.cpp"
        src.write_text('#include <cstdio>\n#include <cstdlib>\n#include "SimulationItems.hpp"\n'
                       'int main(int, char** a) { skirt::NestedLogWavelengthGrid g; g.setMinWavelength(std::strtod(a[1], 0)); g.setMaxWavelength(std::strtod(a[2], 0));\n'
                       '  g.setPoints(std::atoi(a[3])); g.setMinWavelengthSubGrid(std::strtod(a[4], 0)); g.setMaxWavelengthSubGrid(std::strtod(a[5], 0));\n'
                       '  g.setPointsSubGrid(std::atoi(a[6])); g.setup();\n'
                       '  for (int i = 0; i < g.Nlambda(); i++) std::printf("%a %a\\n", g.lambda(i), g.dlambda(i)); return 0; }\n')
        exe = tmp_path / "w"
        subprocess.run([cxx, "-std=c++17", "-O2", f"-I{common.ROOT}/skirt_b200/host", f"-I{common.ROOT}/include", str(src), "-o", str(exe)], check=True)
    spec = L = mixes = None
    if sr.available():
        spec, L, mixes = refspec.reference_spec(configs.c2_params(n=4, nlambda=8, packages=10), threads=1, dustsamples=1)
    for args in NESTED_CASES:
        g = sim.NestedLogWavelengthGrid(*args); n = g.Nlambda
        assert np.all(np.diff(g.lambdav) > 0) and g.lambdav[0] == pytest.approx(args[0]) and g.lambdav[-1] == pytest.approx(args[1])
        if exe:
            out = subprocess.run([str(exe)] + [float(v).hex() if isinstance(v, float) else str(v) for v in args], check=True, capture_output=True, text=True).stdout.split()
            vals = np.array([float.fromhex(x) for x in out]).reshape(-1, 2)
            assert np.array_equal(vals[:, 0], g.lambdav) and np.array_equal(vals[:, 1], g.dlambdav)
        if spec:
            line = [l for l in spec.splitlines() if l.startswith("loggrid")][0]
            S = sr.RefSim(spec.replace(line, "nestedloggrid %r %r %d %r %r %d" % args), luminosities=[np.ones(n).tolist() for _ in L],
                          mixes=[(np.ones(n), np.ones(n), np.zeros(n)) for _ in mixes]).setup()
            lam, dlam = S.wavelengths()
            assert np.array_equal(lam, g.lambdav) and np.array_equal(dlam, g.dlambdav), args


def test_file_wavelength_grid_mirrors(tmp_path):
    """FileWavelengthGrid (FileWavelengthGrid.cpp:22-47): micron -> m by division, sorted, PanWavelengthGrid bin widths; the
    Python and the C++ mirror agree bit for bit"""
    import shutil, subprocess
    from skirt_b200 import simulation as sim
    lam_um = np.array([0.55, 0.1, 2.2, 1000.0, 24.0, 0.3333333333333333])
    f = tmp_path / "grid.dat"; f.write_text(f"{len(lam_um)}\n" + "\n".join(repr(float(v)) for v in lam_um) + "\n")
    g = sim.FileWavelengthGrid(str(f))
    assert np.array_equal(g.lambdav, np.sort(lam_um / 1e6)) and g.Nlambda == 6
    mid = np.sqrt(g.lambdav[:-1] * g.lambdav[1:])
    assert np.array_equal(g.dlambdav, np.concatenate([mid, g.lambdav[-1:]]) - np.concatenate([g.lambdav[:1], mid]))
    assert g.dlambdav.sum() == pytest.approx(g.lambdav[-1] - g.lambdav[0], rel=1e-12)
    with pytest.raises(sim.FatalError, match="Could not open the data file"):
        sim.FileWavelengthGrid(str(tmp_path / "missing.dat"))
    (tmp_path / "short.dat").write_text("2\n1.0\n2.0\n")
    with pytest.raises(sim.FatalError, match="at least three bins"):
        sim.FileWavelengthGrid(str(tmp_path / "short.dat"))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        return
    src = tmp_path / "w.cpp"
    src.write_text('#include <cstdio>\n#include "SimulationItems.hpp"\n'
                   'int main(int, char** a) { skirt::FileWavelengthGrid g; g.setFilename(a[1]); g.setup();\n'
                   '  for (int i = 0; i < g.Nlambda(); i++) std::printf("%a %a\\n", g.lambda(i), g.dlambda(i)); return 0; }\n')
    exe = tmp_path / "w"
    subprocess.run([cxx, "-std=c++17", "-O2", f"-I{common.ROOT}/skirt_b200/host", f"-I{common.ROOT}/include", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(f)], check=True, capture_output=True, text=True).stdout.split()
    vals = np.array([float.fromhex(x) for x in out]).reshape(-1, 2)
    assert np.array_equal(vals[:, 0], g.lambdav) and np.array_equal(vals[:, 1], g.dlambdav)


@pytest.mark.parametrize("kind", list(common.SYM_GRIDS))
def test_symmetric_grid_mirrors_reproduce_the_golden_borders(kind):
    """the same comparison against the committed fixtures (tests/golden/make_symmetric_golden.py), which needs no reference library"""
    tables, _, _ = common.load_golden(kind)
    mine = common.sym_grid_mirror(kind).tables()
    assert mine["kind"] == tables["kind"]
    for key in ("rv", "thetav", "cv", "Rv", "zv"):
        if key in tables:
            assert np.array_equal(np.asarray(mine[key]), np.asarray(tables[key])), f"{kind}: {key}"


@pytest.mark.parametrize("name,kind,lo,hi,bary", [("octtree_s1", 0, 2, 4, False), ("bintree_s1", 1, 2, 10, False), ("octtree_bary_s1", 0, 2, 4, True)])
def test_tree_builder_reproduces_the_golden_trees(name, kind, lo, hi, bary):
    """the tree builder against the trees in the committed fixtures (no reference library needed): with the fixture's decisions
    (and, for the barycentric octree, its split points) the node tables and neighbour lists are the reference's"""
    ref, _, _ = common.load_golden(name)
    child0 = np.asarray(ref["child0"]); rbox = np.asarray(ref["box"]).reshape(-1, 6)
    pos = {tuple(b): i for i, b in enumerate(rbox)}
    tb = hostlib.TreeBuilder(kind, common.C1_BOX, lo, hi)

    def decide(level, boxes):
        ids = np.array([pos[tuple(b)] for b in boxes]); flags = child0[ids] >= 0
        if not bary:
            return flags
        pts = 0.5 * (boxes[:, :3] + boxes[:, 3:])
        pts[flags] = rbox[child0[ids[flags]], 3:]
        return flags, pts
    tb.grow(decide)
    mine = tb.finish(1)
    for k in ("box", "child0", "parent", "cell", "nbrStart", "nbrIds"):
        assert np.array_equal(np.asarray(mine[k]).ravel(), np.asarray(ref[k]).ravel()), f"{name}: {k} differs from the reference"
