"""Pins the restated CPU oracle (oracle/liboracle.so) against the reference: every golden vector under
tests/golden/ (generated from the reference's own translation units) must be reproduced bit for bit, and --
in the build container, where oracle/_ref/libskirtref.so exists -- so must fresh seeded rays at larger sizes."""
import numpy as np
import pytest

import common
from oracle import oracle_py, skirtref

pytestmark = pytest.mark.skipif(not oracle_py.available(), reason="oracle/liboracle.so not built")


@pytest.mark.parametrize("name", common.GEOM_CASES + common.GEOM_CASES_MORE)
def test_oracle_reproduces_golden_paths(name):
    tables, medium, d = common.load_golden(name)
    o = oracle_py.Oracle(tables, medium)
    got = o.path_batch(d["r"], d["k"], ell=0)
    assert common.paths_bit_identical(got, d["paths"]), name
    assert np.array_equal(o.whichcell(d["r"]), d["whichcell"])
    assert np.array_equal(o.opticaldepth(d["r"], d["k"], 0), d["tau_inf"])
    assert np.array_equal(o.opticaldepth(d["r"], d["k"], 0, d["distance"]), d["tau_dist"])
    geo = o.path_batch(d["r"], d["k"], ell=None)
    assert np.array_equal(geo["m"], d["paths"]["m"]) and np.array_equal(geo["s"], d["paths"]["s"]) and not geo["tau"].any()


def test_mt19937_stream_matches_reference_golden():
    z = np.load(common.GOLDEN + "/launch_c2.npz")
    if "uniforms_4357" in z.files:
        assert np.array_equal(oracle_py.uniforms(4357, len(z["uniforms_4357"])), z["uniforms_4357"])
    u = oracle_py.uniforms(4357, 100000)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.005


@pytest.mark.skipif(not skirtref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", ["cart", "octtree1", "octtree2", "bintree0", "amesh", "voronoi"])
def test_oracle_matches_reference_on_fresh_rays(kind):
    mk = lambda spec, **kw: skirtref.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw)
    if kind == "cart":
        S = mk(common.spec_c1(n=50, mesh="sympow 20"))
    elif kind.startswith("octtree"):
        S = mk(common.spec_grid("octtree", search=int(kind[-1]), maxlevel=5))
    elif kind.startswith("bintree"):
        S = mk(common.spec_grid("bintree", search=int(kind[-1]), maxlevel=12))
    elif kind == "amesh":
        S = mk(common.spec_grid("amesh"), amesh=common.make_amesh())
    else:
        S = mk(common.spec_grid("voronoi"), particles=common.voronoi_particles(4000))
    S.setup()
    tables, medium = S.grid_tables(), S.medium()
    r, k = common.rays(20000, common.C1_BOX, 31)
    ref = S.path_batch(r, k, ell=0, nthreads=8)
    got = oracle_py.Oracle(tables, medium).path_batch(r, k, ell=0)
    assert common.paths_bit_identical(got, ref), kind
    assert np.array_equal(oracle_py.Oracle(tables, medium).whichcell(r[:4000]), S.whichcell(r[:4000]))


@pytest.mark.skipif(not skirtref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind", ["sphere1d", "sphere2d", "sphere2d_odd", "cylinder2d"])
def test_oracle_symmetric_grids_match_the_reference(kind):
    """the restated Sphere1DDustGrid / Sphere2DDustGrid / Cylinder2DDustGrid walkers: paths, optical depths and cell lookups bit for bit on the rays
    the GPU parity test uses (isotropic + through the centre, along the axes, in the equatorial plane, from the origin), on the
    reference's own border tables and on those of the product's host mirror"""
    S = skirtref.RefSim(common.spec_grid(kind), luminosities=[[1.0]], mixes=common.mix_v()).setup()
    tables, medium = S.grid_tables(), S.medium()
    r, k = common.sym_rays()
    ref = S.path_batch(r, k, ell=0, nthreads=8)
    assert len(ref["m"]) > len(r)
    for t in (tables, common.sym_grid_mirror(kind).tables()):
        o = oracle_py.Oracle(t, medium)
        assert o.Ncells == S.Ncells
        assert common.paths_bit_identical(o.path_batch(r, k, ell=0), ref), kind
        assert np.array_equal(o.whichcell(r), S.whichcell(r))
        assert np.array_equal(o.opticaldepth(r[:5000], k[:5000], 0), S.opticaldepth_batch(r[:5000], k[:5000], 0))
    # randomPositionInCell: the same MT19937 stream, the same draws in the same order
    S.reset(4357)
    assert np.array_equal(S.random_positions(7, 2000), oracle_py.Oracle(tables, medium).random_positions(7, 4357, 2000))


@pytest.mark.skipif(not skirtref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("tt,extra", [("oct", 0), ("bin", 1)])
def test_oracle_particle_tree_walker_matches_the_reference(tt, extra):
    """ParticleTreeDustGrid::path (its own loop: nearest wall with a positive distance, search from the root;
    ParticleTreeDustGrid.cpp:258-325), including rays parallel to coordinate planes and axes"""
    pts = common.voronoi_particles(3000, seed=5)
    S = skirtref.RefSim(common.spec_grid("particletree_" + tt, maxlevel=extra), luminosities=[[1.0]], mixes=common.mix_v(), particles=pts).setup()
    tables, medium = S.grid_tables(), S.medium()
    assert tables["search"] == 3
    r, k = common.rays(20000, common.C1_BOX, 31)
    k[:300, 0] = 0; k[300:600, 1] = 0; k[600:900, 2] = 0; k[900:1000, :2] = 0
    k /= np.linalg.norm(k, axis=1, keepdims=True)
    o = oracle_py.Oracle(tables, medium)
    assert common.paths_bit_identical(o.path_batch(r, k, ell=0), S.path_batch(r, k, ell=0, nthreads=8))
    assert np.array_equal(o.whichcell(r), S.whichcell(r))


BARY_CASES = [("octtree", 0, 5), ("octtree", 1, 5), ("octtree", 2, 5), ("bintree", 0, 12), ("bintree", 1, 12)]


def barycentric_reference_tree(kind, search, maxlevel):
    """a tree grown by the reference with OctTreeDustGrid::barycentric / BinTreeDustGrid's Barycenter direction method"""
    spec = common.spec_grid(kind, search=search, maxlevel=maxlevel).replace(" 0 50\n", " 1 50\n")
    return skirtref.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v()).setup()


@pytest.mark.skipif(not skirtref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("kind,search,maxlevel", BARY_CASES)
def test_oracle_follows_barycentric_trees(kind, search, maxlevel):
    S = barycentric_reference_tree(kind, search, maxlevel)
    r, k = common.rays(20000, common.C1_BOX, 31)
    o = oracle_py.Oracle(S.grid_tables(), S.medium())
    assert common.paths_bit_identical(o.path_batch(r, k, ell=0), S.path_batch(r, k, ell=0, nthreads=8))
    assert np.array_equal(o.whichcell(r[:5000]), S.whichcell(r[:5000]))
