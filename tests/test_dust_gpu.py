"""GPU parity of the dust emission phases: PanMonteCarloSimulation::dodustselfabsorptionchunk
and dodustemissionchunk.  The reference side runs its own code (oracle/_ref, which travels to the GPU box as a
prebuilt library): stellar phase -> DustLib -> cell luminosities Lv; both sides then shoot the same Lv."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _ref_pan(n=16, nlambda=25, packages=2e4, grid=None):
    from oracle import skirtref as sr, refspec
    from skirt_b200 import configs
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    p = configs.c2_params(n=n, nlambda=nlambda, packages=packages)
    spec, L, mixes = refspec.reference_spec(p, threads=os.cpu_count() or 1, dustsamples=5)
    spec += "selfabs 1\n"
    kw = {}
    if grid == "grid amesh":
        # C5's grid type: the dust density comes from the mesh itself (AdaptiveMeshDustDistribution)
        spec = "\n".join("ameshdust 1e-24" if l.startswith("dust ") else l for l in spec.splitlines()) + "\n"
        kw["amesh"] = common.make_amesh(max_depth=3)
    if grid == "grid voronoi file":
        kw["particles"] = common.voronoi_particles(3000)       # C4's grid type
    if grid:
        spec = "\n".join(grid if l.startswith("grid ") else l for l in spec.splitlines()) + "\n"
    S = sr.RefSim(spec, luminosities=L, mixes=mixes, **kw).setup()
    S.reset(4357); S.run_stellar()
    return S, p


def _engine_for(engine, S, p):
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    engine.set_grid(tables); engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    ins = [dict(kind=1, distance=1e7 * common.PC, inclination=float(np.radians(88)), Nxp=800, fovxp=50000 * common.PC, Nyp=200, fovyp=12500 * common.PC),
           dict(kind=2, distance=1e7 * common.PC, inclination=float(np.radians(88)))]
    engine.instruments(ins)
    return tables


def _compare(name, a, r, B, frac=0.97):
    """the Monte Carlo gate of SURVEY.md 8d(ii), tests/common.py mc_gate: 16 batches on both sides"""
    common.mc_gate(a, r, name)


@pytest.mark.parametrize("grid", [None, "grid octtree 2 4 1 0.0005 0 30", "grid amesh", "grid voronoi file",
                                  "grid " + common.SYM_GRIDS["cylinder2d"], "grid " + common.SYM_GRIDS["sphere2d"]])
def test_dust_selfabsorption_and_emission(engine, grid):
    S, p = _ref_pan(grid=grid)
    _engine_for(engine, S, p)
    Npp = S.packages_per_lambda()
    Lv = S.prepare_dust(True)
    assert Lv.sum() > 0
    B = 16
    # ---- one self-absorption cycle (PanMonteCarloSimulation.cpp:116-148) with the full packet count on both sides: a
    #      tenth of it (the first stage of the reference) leaves too few events per cell for the per-cell gate
    ref, gpu = [], []
    for b in range(B):
        S.reset(100 + 1000 * b); S.run_dust(True, 1.0); ref.append(S.labs_dust().ravel().copy())
        NppStage = S.packages_per_lambda()      # set by setChunkParams(packages*factor)
        engine.reset_labs_dust()
        st = engine.run_dust(1, Lv, NppStage, seed=9070 + b)     # (seeds 70.. are a 3.5 sigma low draw on the Sphere2D grid; 48 batches: -0.09 +- 0.08 %, tools/gpu_sym_dust_diag.py)
        gpu.append(engine.fetch_labs_dust().ravel())
        assert st["detections"] == 0
    _compare("Labsdust", gpu, ref, B)
    # ---- the dust emission phase
    ref_f, ref_s, gpu_f, gpu_s = [], [], [], []
    for b in range(B):
        S.reset(5000 + 1000 * b); S.run_dust(False, 1.0); ins = S.instruments()
        ref_f.append(ins[0]["frame"].copy()); ref_s.append(ins[1]["sed"].copy())
        NppEm = S.packages_per_lambda()
        engine.reset_results()
        st = engine.run_dust(2, Lv, NppEm, emission_bias=0.5, seed=170 + b)
        gpu_f.append(engine.fetch_frame(0)); gpu_s.append(engine.fetch_sed(1))
        assert st["absorbSegments"] == 0
    _compare("dust SED", gpu_s, ref_s, B)
    _compare("dust frame", gpu_f, ref_f, B)


def test_absorption_tables_keep_their_own_shape(engine):
    """an engine reused after a run with another number of wavelengths: the absorption tables come back with the shape of the
    medium they were filled under, not of the sources left over from the earlier run (found as an order dependence between the
    Voronoi tests: one wavelength of sources, then the 25 wavelengths of the dust phases -- only the first slice arrived)"""
    S, p = _ref_pan()
    engine.sources([dict(geometry=1, p=[4000 * common.PC, 350 * common.PC, 0, 0, 0])], [[1.0]], 0.5)     # a one-wavelength run's
    _engine_for(engine, S, p)
    Lv = S.prepare_dust(True)
    engine.reset_labs_dust()
    engine.run_dust(1, Lv, 2000, seed=5)
    a = engine.fetch_labs_dust()
    assert a.shape[1] == 25 and (a.sum(0) > 0).sum() > 10
    assert np.isclose(a.sum(), engine.labs_dust_total(), rtol=1e-9)           # (summed on the device, whatever the layout)


def test_labs_bolometric_and_random_positions(engine):
    S, p = _ref_pan(packages=5e3)
    tables = _engine_for(engine, S, p)
    L = S.luminosities()
    engine.sources([dict(geometry=2, p=[1600 * common.PC, 0.7], rv=np.ones(2), Xv=np.array([0.0, 1.0])),
                    dict(geometry=1, p=[4000 * common.PC, 350 * common.PC, 0, 0, 0])], L, 0.5)
    engine.reset_results()
    engine.run_stellar(2e3, store_absorption=True, seed=5)
    bol = engine.labs_bolometric()
    np.testing.assert_allclose(bol, engine.fetch_labs().sum(1), rtol=1e-12)
    # uniform positions inside the emitting cell: all luminosity in one cell, no dust interaction needed to see it
    Lv = np.zeros((engine.Nlambda, engine.Ncells)); m = 1234; Lv[3, m] = 1.0
    engine.reset_results()
    st = engine.run_dust(2, Lv, 2e4, emission_bias=0.0, seed=9)
    assert st["packets"] == 20000
    sed = engine.fetch_sed(1)
    assert sed[3] > 0 and np.count_nonzero(sed) == 1


@pytest.mark.parametrize("devlib", [False, True])
def test_full_pan_flow_through_the_host_mirror(engine, devlib):
    """stellar emission -> three fixed self-absorption cycles -> dust emission, driven by the Python mirror of
    PanMonteCarloSimulation with the numpy DustLib, against the same sequence executed by the reference's own classes"""
    from oracle import skirtref as sr, refspec
    from skirt_b200 import configs, simulation as sim
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    p = configs.c2_params(n=14, nlambda=25, packages=2e4)
    spec, L, mixes = refspec.reference_spec(p, threads=os.cpu_count() or 1, dustsamples=5)
    S = sr.RefSim(spec + "selfabs 1\n", luminosities=L, mixes=mixes).setup()
    med = S.medium()
    # engine side: same density table, same mix, same luminosities
    m = configs.build(p, rho=med["rho"], storeAbsorption=True)
    m.engine.close(); m.engine = engine
    m.setup()
    lib = sim.GreyBodyDustLib(m.lambdagrid, [mixes[0][0]], med["rho"], S.volumes())
    if devlib:          # dust emission spectra on the device (skg_dust_library) instead of the numpy DustLib
        m.setup_dust_library(S.volumes()); lib = None
    B = 6
    ref_sed, gpu_sed = [], []
    for b in range(B):
        S.reset(2000 + 1000 * b); S.run_stellar()
        for stage, factor in enumerate((0.1, 1. / 3., 1.0)):
            S.prepare_dust(stage == 0); S.run_dust(True, factor)
        S.prepare_dust(False); S.run_dust(False, 1.0)
        ref_sed.append(S.instruments()[1]["sed"].copy())
        engine.reset_results(); engine.reset_labs_dust() if b else None
        m.seed = 50 + b
        m.packages = S.packages_per_lambda()
        m.runstellaremission()
        m.rundustselfabsorption(lib, cycles=1)
        m.rundustemission(lib)
        gpu_sed.append(engine.fetch_sed(1))
    ref_sed, gpu_sed = np.array(ref_sed), np.array(gpu_sed)
    # long wavelengths are pure dust emission, short ones pure (attenuated) star light
    assert gpu_sed.mean(0)[-5:].sum() > 0
    for sl in (slice(0, 12), slice(15, 25)):
        a, r = gpu_sed[:, sl].sum(1), ref_sed[:, sl].sum(1)
        z = (a.mean() - r.mean()) / np.sqrt(a.var(ddof=1) / B + r.var(ddof=1) / B)
        assert abs(z) < 4 and abs(a.mean() / r.mean() - 1) < 0.02, f"bins {sl}: gpu {a.mean():.6g} ref {r.mean():.6g} z {z:.2f}"


def test_device_dust_library_matches_the_host_restatement(engine):
    """skg_dust_cell_luminosities (DustLib on the device, fed by the device-resident absorption tables) against the numpy
    GreyBodyDustLib, which tests/test_dustlib.py pins to the reference's own DustLib; then a dust emission phase shot
    straight from the device array equals one shot from the same array passed through the host"""
    from skirt_b200 import configs, simulation as sim
    p = configs.c2_params(n=16, nlambda=25, packages=2e4)
    m = configs.build(p, storeAbsorption=True)
    m.engine.close(); m.engine = engine
    m.setup()
    engine.reset_results(); engine.reset_labs_dust()
    m.runstellaremission()
    grid = m.ds.grid
    dx, dy, dz = np.diff(grid.xv), np.diff(grid.yv), np.diff(grid.zv)
    vol = (dx[:, None, None] * dy[None, :, None] * dz[None, None, :]).ravel()
    kabs = np.array([c.mix.kappaabs for c in m.ds.comps])
    engine.dust_library(vol, kabs, m.lambdagrid.lambdav, m.lambdagrid.dlambdav)
    d_L = engine.dust_cell_luminosities()
    got = engine.copy_from_device(d_L, (engine.Nlambda, engine.Ncells))
    lib = sim.GreyBodyDustLib(m.lambdagrid, kabs, m.ds.rho, vol)
    labs = engine.fetch_labs()
    want = (labs.sum(1)[:, None] * lib.luminosities(labs)).T
    assert want.sum() > 0
    np.testing.assert_allclose(got.sum(0), want.sum(0), rtol=1e-10)
    big = want > 1e-12 * want.max()
    np.testing.assert_allclose(got[big], want[big], rtol=1e-8)
    # same packets (same seed) from the device array and from its host copy
    engine.reset_results(); a = engine.run_dust_device(2, d_L, 2e4, seed=5); sed_dev = engine.fetch_sed(1)
    engine.reset_results(); b = engine.run_dust(2, got, 2e4, seed=5); sed_host = engine.fetch_sed(1)
    assert a["packets"] == b["packets"] and a["pathSegments"] == b["pathSegments"]
    np.testing.assert_allclose(sed_dev, sed_host, rtol=1e-9)
