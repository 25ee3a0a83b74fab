"""Golden vectors for the grids added after the first fixtures: the grids with symmetries (Sphere1D, Sphere2D with and without a
mesh point at pi/2, Cylinder2D), the particle tree and a barycentric octree -- from the reference's OWN code
(oracle/_ref/libskirtref.so), in the layout of make_golden.py (tables and medium as the reference built them, seeded + adversarial
rays, the reference's path records, whichcell and optical depths).  Run in the build container only:

    python tests/golden/make_symmetric_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                        # noqa: E402
from make_golden import common, sr              # noqa: E402

if __name__ == "__main__":
    mk = lambda spec, **kw: sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw)
    for i, kind in enumerate(common.SYM_GRIDS):
        mg.geometry_case(kind, mk(common.spec_grid(kind)), 201 + i)
    mg.geometry_case("particletree_oct", mk(common.spec_grid("particletree_oct", maxlevel=0), particles=common.voronoi_particles(800, seed=5)), 211)
    mg.geometry_case("octtree_bary_s1", mk(common.spec_grid("octtree", search=1, maxlevel=4).replace(" 0 50\n", " 1 50\n")), 212)
