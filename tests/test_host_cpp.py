"""The C++ host layer (skirt_b200/host: simulation items with the reference's names over the C ABI) through its
command-line driver.  CPU part: argument/property validation with the reference's error messages and the
no-fallback rule; GPU part: a C1 run end to end, compared with the Python mirror and the reference's golden run."""
import json
import os
import subprocess

import numpy as np
import pytest

import common

RUN = os.path.join(common.ROOT, "skirt_b200", "skirt_b200_run")
PC = common.PC


def params(packages=2e5, n=24, extra=()):
    lines = ["sim oligo", f"packages {packages!r}", "seed 4357", "wavelengths 0.55e-6", common.box_line(common.C1_BOX),
             f"grid cartesian {n} {n} {n} lin lin lin", "storeabs 1",
             f"dustmix table {common.MIX_V['kabs']!r} {common.MIX_V['ksca']!r} {common.MIX_V['g']!r}",
             f"dust 1.0 0.55e-6 expdisk {4000*PC!r} {140*PC!r} 0 0",
             f"stellar 1.0 expdisk {4000*PC!r} {350*PC!r} 0 0",
             f"instrument frame i88 {1e7*PC!r} {float(np.radians(88))!r} 0 0 200 {50000*PC!r} 50 {12500*PC!r}",
             f"instrument sed s88 {1e7*PC!r} {float(np.radians(88))!r} 0 0"] + list(extra)
    return "\n".join(lines) + "\n"


def run(tmp_path, text):
    f = tmp_path / "sim.txt"; f.write_text(text)
    return subprocess.run([RUN, str(f), str(tmp_path / "out")], capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
def test_property_validation_messages(tmp_path):
    r = run(tmp_path, params(extra=["packages -5"]))
    assert r.returncode == 1 and "Number of photon packages is negative" in r.stderr
    r = run(tmp_path, params().replace("expdisk %r %r 0 0" % (4000 * PC, 140 * PC), "expdisk -1 %r 0 0" % (140 * PC)))
    assert r.returncode == 1 and "radial scale length hR should be positive" in r.stderr
    r = run(tmp_path, params().replace(common.box_line(common.C1_BOX), "box 1 0 0 1 0 1"))
    assert r.returncode == 1 and "extent of the box should be positive in the X direction" in r.stderr


@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run(tmp_path, params(packages=100.0, n=4))
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
def test_c1_through_the_cpp_host(tmp_path, engine):
    r = run(tmp_path, params())
    assert r.returncode == 0, r.stderr
    st = json.loads(r.stdout.strip().splitlines()[-1])
    assert st["packets"] == 200000 and st["cells"] == 24 ** 3
    load = lambda name: np.fromfile(tmp_path / f"out_{name}.f64")
    sed, frame, labs, rho = load("s88_sed"), load("i88_frame"), load("Labs"), load("rho")
    assert frame.size == 200 * 50 and labs.size == 24 ** 3
    # the same configuration through the Python mirror: identical density table, statistically identical results
    from skirt_b200 import simulation as sim
    lg = sim.OligoWavelengthGrid([0.55e-6])
    b = common.C1_BOX
    grid = sim.CartesianDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], sim.LinMesh(24), sim.LinMesh(24), sim.LinMesh(24))
    mix = sim.TableDustMix(common.MIX_V["kabs"], common.MIX_V["ksca"], common.MIX_V["g"])
    ds = sim.DustSystem(grid, [sim.DustComp(sim.ExpDiskGeometry(4000 * PC, 140 * PC), mix, 1.0, 0.55e-6)], lg)
    np.testing.assert_allclose(rho, ds.rho.ravel(), rtol=1e-12)
    cfg = common.cfg_c1(n=24, packages=2e5, storeabs=1)
    common.setup_engine(engine, cfg, grid.tables(), ds.medium(), np.array([[1.0]]))
    engine.reset_results(); engine.run_stellar(2e5, store_absorption=True, seed=99)
    for name, a, b_ in (("sed", sed.sum(), engine.fetch_sed(1).sum()), ("frame", frame.sum(), engine.fetch_frame(0).sum()),
                        ("labs", labs.sum(), engine.fetch_labs().sum())):
        assert abs(a / b_ - 1) < 0.01, f"{name}: C++ host {a} vs Python mirror {b_}"
    # against the reference's own 16 runs of this configuration (golden), whose density table comes from 20 random
    # samples per cell instead of the lattice (a coarse 24^3 grid, so the tables differ at the few-percent level)
    _, _, g = common.load_golden_mc()
    assert abs(sed[0] / g["sed_total_mean"][0] - 1) < 0.06
    assert abs(frame.sum() / g["frame_total_mean"][0] - 1) < 0.06
    assert abs(labs.sum() / g["labs_total_mean"][0] - 1) < 0.12


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
def test_pan_flow_through_the_cpp_host(tmp_path):
    """stellar emission -> self-absorption cycles (until convergence) -> dust emission, all driven by the C++ host layer:
    energy bookkeeping of the result"""
    dat = os.path.join(common.ROOT, "skirt_b200", "data", "interstellar_dustmix.dat")
    text = "\n".join([
        "sim pan", "packages 20000.0", "loggrid 1e-07 0.001 25", common.box_line(common.C1_BOX), "grid cartesian 16 16 16 lin lin lin",
        f"dustmix interstellar {dat}", f"dust 1.0 5.5e-07 expdisk {4000*PC!r} {140*PC!r} 0 0",
        f"stellar bb:3500:1.15e36 sersic 2.0 {1600*PC!r} 0.7", f"stellar bb:10000:1.92e36 expdisk {4000*PC!r} {350*PC!r} 0 0",
        f"instrument sed s {1e7*PC!r} 0.5 0 0", "dustemission 1", "selfabs 1"]) + "\n"
    r = run(tmp_path, text)
    assert r.returncode == 0, r.stderr
    st = json.loads(r.stdout.strip().splitlines()[-1])
    assert 4 <= st["selfabs_cycles"] <= 40
    sed = np.fromfile(tmp_path / "out_s_sed.f64"); labs = np.fromfile(tmp_path / "out_Labs.f64")
    Lstar = 1.15e36 + 1.92e36
    # every emitted watt is either seen directly / after scattering, or absorbed and re-emitted by the dust: the SED
    # integrated over wavelength recovers the stellar luminosity to within the anisotropy of one viewing direction
    assert 0.7 < sed.sum() / Lstar < 1.3
    assert 0.02 < labs.sum() / Lstar < 0.6
    assert sed[-8:].sum() > 0 and sed[:5].sum() > 0


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
@pytest.mark.parametrize("kind", ["octtree", "bintree", "amesh", "voronoi", "sphere2d_odd", "cylinder2d"])
def test_other_grids_through_the_cpp_host(tmp_path, engine, kind):
    """the C++ host builds tree / adaptive-mesh / Voronoi grids itself (skirt_b200/host/GridBuilders.cpp; tree subdivision and
    cell densities sampled on the device) and shoots through them; the Python mirror does the same through libskirthost.so:
    the two product-side hosts agree on the detected and absorbed luminosity"""
    from skirt_b200 import configs, hostlib, simulation as sim
    mixline = f"dustmix table {common.MIX_V['kabs']!r} {common.MIX_V['ksca']!r} {common.MIX_V['g']!r}"
    dust = f"dust 1.0 0.55e-6 expdisk {4000*PC!r} {140*PC!r} 0 0"
    b = common.C1_BOX
    lg = sim.OligoWavelengthGrid([0.55e-6])
    mix = sim.TableDustMix(common.MIX_V["kabs"], common.MIX_V["ksca"], common.MIX_V["g"])
    comp = sim.DustComp(sim.ExpDiskGeometry(4000 * PC, 140 * PC), mix, 1.0, 0.55e-6)
    if kind in ("octtree", "bintree"):
        lo, hi = (2, 5) if kind == "octtree" else (6, 14)
        grid_line = f"grid {kind} {lo} {hi} 1 1e-4 50"
        cls = sim.OctTreeDustGrid if kind == "octtree" else sim.BinTreeDustGrid
        grid = cls(b[0], b[1], b[2], b[3], b[4], b[5], lo, hi, "Neighbor", 50, 0.0, 1e-4)
    elif kind in common.SYM_GRIDS:
        grid_line = "grid " + common.SYM_GRIDS[kind]; grid = common.sym_grid_mirror(kind)
    elif kind == "amesh":
        nxyz, val = configs.synthetic_amesh(root=4, depth=3, frac=2e-3)
        f = tmp_path / "mesh.txt"
        f.write_text("# synthetic adaptive mesh\n" + "\n".join(f"! {n[0]} {n[1]} {n[2]}" if n[0] else repr(float(v)) for n, v in zip(nxyz, val)) + "\n")
        grid_line = f"grid amesh {f} 1e-24"; dust = "meshdust"
        grid = sim.AdaptiveMeshDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], nxyz, val, densityUnits=1e-24)
        comp = sim.DustComp(None, mix, 1.0, 0.55e-6)
    else:
        if not hostlib.voronoi_available():
            pytest.skip("libskirthost.so was built without Voro++")
        pts = common.voronoi_particles(3000)
        f = tmp_path / "particles.txt"; np.savetxt(f, pts, fmt="%.17g")
        grid_line = f"grid voronoi {f}"
        grid = sim.VoronoiDustGrid(b[0], b[1], b[2], b[3], b[4], b[5], pts)
    text = "\n".join(["sim oligo", "packages 200000.0", "seed 4357", "wavelengths 0.55e-6", common.box_line(b), grid_line, "storeabs 1", mixline, dust,
                      f"stellar 1.0 expdisk {4000*PC!r} {350*PC!r} 0 0", f"instrument sed s88 {1e7*PC!r} {float(np.radians(88))!r} 0 0"]) + "\n"
    r = run(tmp_path, text)
    assert r.returncode == 0, r.stderr
    st = json.loads(r.stdout.strip().splitlines()[-1])
    sed = np.fromfile(tmp_path / "out_s88_sed.f64"); labs = np.fromfile(tmp_path / "out_Labs.f64")
    # the same configuration through the Python mirror (its own tree: the device samples differ, the grids are statistically alike)
    ds = sim.DustSystem(grid, [comp], lg)
    ss = sim.StellarSystem([sim.StellarComp(sim.ExpDiskGeometry(4000 * PC, 350 * PC), [1.0])])
    ins = sim.InstrumentSystem([sim.SEDInstrument("s88", 1e7 * PC, float(np.radians(88)))])
    m = sim.MonteCarloSimulation(lg, ss, ds, ins, packages=2e5, storeAbsorption=True, engine=engine).setup()
    engine.reset_results(); m.runstellaremission()
    assert abs(st["cells"] / engine.Ncells - 1) < (0.15 if kind in ("octtree", "bintree") else 1e-12)
    assert st["packets"] == 200000 and labs.size == st["cells"]
    assert abs(sed[0] / engine.fetch_sed(0)[0] - 1) < 0.03, f"{kind}: SED {sed[0]} vs {engine.fetch_sed(0)[0]}"
    assert abs(labs.sum() / engine.fetch_labs().sum() - 1) < 0.03
