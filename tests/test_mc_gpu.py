"""GPU parity of the Monte Carlo outputs (north_star gate 2): frames, SEDs and per-cell absorbed luminosity of
the engine's stellar emission phase must agree with the reference's within 3 sigma of the combined Monte Carlo
noise (the RNG streams necessarily differ: Philox per packet vs MT19937 per thread).  The reference side is the
golden batch statistics in tests/golden/mc_c1.npz (16 independent runs of the reference's own code)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _run_batches(engine, tables, medium, cfg, L, Npp, batches, store=True):
    common.setup_engine(engine, cfg, tables, medium, L)
    fr, se, la = [], [], []
    for b in range(batches):
        engine.reset_results()
        st = engine.run_stellar(Npp, store_absorption=store, seed=500 + b)
        assert st["packets"] == int(Npp)
        fr.append(engine.fetch_frame(0)); se.append(engine.fetch_sed(1))
        if store:
            la.append(engine.fetch_labs().ravel())
    return [np.array(a) for a in (fr, se, la)]


def test_c1_outputs_within_three_sigma(engine):
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=2e5, storeabs=1)
    B = 16
    fr, se, la = _run_batches(engine, tables, medium, cfg, g["L"], float(g["Npp"][0]), B)
    for name, a in (("frame", fr), ("sed", se), ("labs", la)):
        mean = a.mean(0); sem = a.std(0, ddof=1) / np.sqrt(B)
        z = common.zscores(mean, sem, g[name + "_mean"], g[name + "_sem"])
        tot = a.reshape(B, -1).sum(1)
        zt = (tot.mean() - g[name + "_total_mean"][0]) / np.hypot(tot.std(ddof=1) / np.sqrt(B), g[name + "_total_sem"][0])
        assert abs(zt) < 3.5, f"{name}: total differs by {zt:.2f} sigma"
        if len(z) > 10:
            # per-bin z scores: Student-t tails with 15+15 degrees of freedom are wider than a Gaussian's
            assert np.mean(np.abs(z) < 3) > 0.98, f"{name}: only {np.mean(np.abs(z) < 3):.4f} of bins within 3 sigma"
            assert abs(z.mean()) < 0.1, f"{name}: systematic offset, mean z = {z.mean():.3f}"
            assert 0.8 < z.std() < 1.3, f"{name}: z scatter {z.std():.3f}"


def test_energy_budget_without_scattering_bias(engine):
    """conservation: detected SED at a face-on observer with no dust equals L/(4 pi) x 4 pi = emitted luminosity"""
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=1e5)
    med0 = dict(medium); med0["rho"] = np.zeros_like(medium["rho"])
    common.setup_engine(engine, cfg, tables, med0, g["L"])
    engine.reset_results()
    engine.run_stellar(1e5, seed=3)
    sed = engine.fetch_sed(1)
    np.testing.assert_allclose(sed[0], g["L"].sum(), rtol=1e-9)     # every packet peels off its full weight once


def test_streams_are_reproducible_and_disjoint(engine):
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=2e4)
    common.setup_engine(engine, cfg, tables, medium, g["L"])
    out = []
    for off in (0, 0, 20000):
        engine.reset_results(); engine.run_stellar(2e4, total_packages=4e4, seed=11, stream_offset=off)
        out.append(engine.fetch_sed(1)[0])
    assert np.isclose(out[0], out[1], rtol=1e-9)        # same stream -> same result up to atomic summation order
    assert out[0] != out[2]                              # disjoint Philox counters -> a different realisation


@pytest.mark.parametrize("kind", ["octtree", "bintree", "amesh", "voronoi", "particletree_oct", "sphere1d", "sphere2d", "cylinder2d"])
def test_other_grids_against_reference_runs(engine, kind):
    """the same 3-sigma gate on the hierarchical / unstructured grids, against runs of the reference's own code
    (oracle/_ref travels to the GPU box as a prebuilt library; skipped where it is absent)"""
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    import os
    extra = ("storeabs 1",)
    kw = {}
    if kind == "amesh":
        kw["amesh"] = common.make_amesh(max_depth=3)
    if kind in ("voronoi", "particletree_oct"):
        kw["particles"] = common.voronoi_particles(3000)
    spec = common.spec_grid(kind, search=1, maxlevel=4 if kind == "octtree" else 10, packages=1e5, threads=os.cpu_count() or 1, extra=extra)
    S = sr.RefSim(spec, luminosities=[[1.0]], mixes=common.mix_v(), **kw).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    if kind in common.SYM_GRIDS:
        tables = common.sym_grid_mirror(kind).tables()          # the product's own host mirror of these grids
    Npp = S.packages_per_lambda()
    engine.set_grid(tables); engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    if kind == "sphere1d":
        from skirt_b200 import simulation as sim
        engine.sources([sim.SersicGeometry(2.0, 1500 * common.PC).sampler()], L, 0.5)
    else:
        engine.sources([dict(geometry=1, p=[4000 * common.PC, 350 * common.PC, 0, 0, 0])], L, 0.5)
    engine.instruments([dict(kind=2, distance=1e7 * common.PC, inclination=float(np.radians(88)))])
    B = 16
    ref_s, ref_l, gpu_s, gpu_l = [], [], [], []
    for b in range(B):
        S.reset(300 + 1000 * b); S.run_stellar()      # Random seeds thread t with seed+t: keep the batches disjoint
        ref_s.append(S.instruments()[0]["sed"].copy()); ref_l.append(S.labs().ravel().copy())
        engine.reset_results(); engine.run_stellar(Npp, store_absorption=True, seed=40 + b)
        gpu_s.append(engine.fetch_sed(0)); gpu_l.append(engine.fetch_labs().ravel())
    # the gate of SURVEY.md 8d(ii) (tests/common.py mc_gate): 16 batches on both sides, per-cell z-scores, totals at 3.5 sigma
    common.mc_gate(gpu_s, ref_s, f"{kind}/sed")
    common.mc_gate(gpu_l, ref_l, f"{kind}/labs", min_bins=0.4)


def test_c2_benchmark_configuration_against_reference_runs(engine):
    """the configuration bench.py measures (C2: Sersic bulge + exponential disk, 50 wavelengths, InterstellarDustMix,
    absorption stored, frame + SED), at a reduced grid / packet count, against the reference's PanMonteCarloSimulation:
    SED per wavelength, frame and absorbed luminosity per wavelength within the combined Monte Carlo noise"""
    import os
    from oracle import skirtref as sr, refspec
    from skirt_b200 import configs
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    p = configs.c2_params(n=32, nlambda=50, packages=2e4)
    spec, L, mixes = refspec.reference_spec(p, threads=os.cpu_count() or 1, dustsamples=5)
    S = sr.RefSim(spec, luminosities=L, mixes=mixes).setup()
    med = S.medium()
    m = configs.build(p, rho=med["rho"], storeAbsorption=True)
    m.engine.close(); m.engine = engine
    m.setup()
    np.testing.assert_allclose(m.ds.kext, med["kext"], rtol=1e-9)          # same opacities on both sides
    m.packages = S.packages_per_lambda()
    B = 16
    ref = dict(sed=[], frame=[], labs=[]); gpu = dict(sed=[], frame=[], labs=[])
    for b in range(B):
        S.reset(77 + 1000 * b); S.run_stellar(); ins = S.instruments()
        ref["sed"].append(ins[1]["sed"].copy()); ref["frame"].append(ins[0]["frame"].reshape(50, -1).sum(1)); ref["labs"].append(S.labs().sum(0))
        engine.reset_results(); m.seed = 400 + b; m.runstellaremission()
        gpu["sed"].append(engine.fetch_sed(1)); gpu["frame"].append(engine.fetch_frame(0).reshape(50, -1).sum(1)); gpu["labs"].append(engine.fetch_labs().sum(0))
    for name in ref:
        common.mc_gate(gpu[name], ref[name], "C2/" + name)


def test_configuration_errors_are_reported(engine):
    """misuse of the C ABI comes back as an error message (the reference-side adapter turns it into FATALERROR)"""
    import skirt_b200 as sk
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=1e3)
    common.setup_engine(engine, cfg, tables, medium, g["L"])
    with pytest.raises(sk.EngineError, match="scattBias"):
        engine.run_stellar(1e3, scatt_bias=1.5)
    with pytest.raises(sk.EngineError, match="wavelength range"):
        engine.run_stellar(1e3, ell_end=5)
    with pytest.raises(sk.EngineError, match="negative|1e15"):
        engine.run_stellar(-5)
    # a medium with another number of wavelengths invalidates the detector arrays
    med2 = {k: (np.concatenate([v, v], axis=1) if k != "rho" else v) for k, v in medium.items()}
    engine.medium(med2["rho"], med2["kext"], med2["ksca"], med2["g"])
    engine.sources(cfg["sources"], np.array([[1.0, 1.0]]), 0.5)
    with pytest.raises(sk.EngineError, match="skg_instruments again"):
        engine.run_stellar(1e3)
    with pytest.raises(sk.EngineError, match="unsupported source geometry"):
        engine.sources([dict(geometry=7, p=[1, 1])], np.array([[1.0, 1.0]]), 0.5)
    with pytest.raises(sk.EngineError, match="skg_dust_library"):
        engine.dust_cell_luminosities()


def test_full_instrument_channels_against_reference_runs(engine):
    """FullInstrument::detect (FullInstrument.cpp:107-172, unpolarised): transparent / direct / scattered data cubes and SEDs
    plus one per scattering level, against 16 runs of the reference's own FullInstrument (tests/golden/mc_full.npz)"""
    tables, medium, g = common.load_golden_mc("mc_full")
    cfg = common.cfg_full()
    common.setup_engine(engine, cfg, tables, medium, g["L"])
    B = 16; Npp = float(g["Npp"][0])
    chans = (0, 1, 2, 5, 6)
    fr = {c: [] for c in range(7)}; se = {c: [] for c in range(7)}; tot = []
    for b in range(B):
        engine.reset_results()
        engine.run_stellar(Npp, seed=900 + b)
        for c in range(7):
            fr[c].append(engine.fetch_frame_channel(0, c)); se[c].append(engine.fetch_sed_channel(0, c))
        tot.append(engine.fetch_sed(1))
    fr = {c: np.array(v) for c, v in fr.items()}; se = {c: np.array(v) for c, v in se.items()}; tot = np.array(tot)
    # exact structure of one run: no dust emission channels in the stellar phase; transparent SED = emitted luminosity;
    # direct + scattered = what the SEDInstrument in the same direction saw; levels 1 and 2 are part of 'scattered'
    assert not fr[3].any() and not fr[4].any() and not se[3].any() and not se[4].any()
    np.testing.assert_allclose(se[0][:, 0], g["L"].sum(), rtol=1e-9)
    np.testing.assert_allclose(se[1] + se[2], tot, rtol=1e-9)
    assert np.all(se[5] + se[6] < se[2]) and np.all(fr[5] + fr[6] <= fr[2] * (1 + 1e-12))
    for c in chans:
        for name, a in (("frame", fr[c]), ("sed", se[c])):
            mean = a.mean(0); sem = a.std(0, ddof=1) / np.sqrt(B)
            gm, gs = g[f"{name}{c}_mean"], g[f"{name}{c}_sem"]
            zt = (a.reshape(B, -1).sum(1).mean() - gm.sum()) / max(np.hypot(a.reshape(B, -1).sum(1).std(ddof=1) / np.sqrt(B), np.sqrt((gs ** 2).sum())), 1e-12 * gm.sum())
            assert abs(zt) < 3.5, f"channel {c} {name}: total differs by {zt:.2f} sigma"
            z = common.zscores(mean, sem, gm, gs)
            if len(z) > 10:
                assert np.mean(np.abs(z) < 3) > 0.97, f"channel {c} {name}: only {np.mean(np.abs(z) < 3):.4f} of bins within 3 sigma"
                assert abs(z.mean()) < 0.15, f"channel {c} {name}: systematic offset, mean z = {z.mean():.3f}"
    z = (tot.mean(0) - g["sedtotal_mean"]) / np.hypot(tot.std(0, ddof=1) / np.sqrt(B), g["sedtotal_sem"])
    assert np.all(np.abs(z) < 3.5)


def test_full_instrument_errors(engine):
    tables, medium, g = common.load_golden_mc("mc_full")
    cfg = common.cfg_full()
    common.setup_engine(engine, cfg, tables, medium, g["L"])
    with pytest.raises(Exception):
        engine.fetch_frame_channel(0, 7)                 # 5 + 2 channels
    with pytest.raises(Exception):
        engine.fetch_sed_channel(1, 0)                   # an SEDInstrument has no channels
    bad = [dict(cfg["instruments"][0], scatteringLevels=-1)]
    with pytest.raises(Exception):
        engine.instruments(bad)


def test_c3_configuration_against_reference_runs(engine):
    """C3 at a reduced size: adaptive octree (Neighbor search), stars in a spiral-arm exponential disk
    (SpiralStructureGeometryDecorator; the dust disk stays axisymmetric, as the reference's face-on normalisation demands), forced scattering, six frame instruments at inclinations 0 ... 90 degrees (five
    observer directions on the engine, one traversal each per peel-off), against runs of the reference's own classes"""
    import os
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    PC = common.PC
    sp = f"spiral 2 {float(np.radians(20))!r} {4000 * PC!r} 0.0 1.0 1"
    incl = (0, 30, 60, 80, 88, 90)
    lines = ["sim oligo", f"threads {os.cpu_count() or 1}", "seed 4357", "packages 100000.0", "wavelengths 0.55e-6", common.box_line(common.C1_BOX),
             "grid octtree 2 6 1 2e-05 0 50", "dustsamples 10", "storeabs 1",
             f"stellar expdisk {4000 * PC!r} {350 * PC!r} 0 0 {sp}", f"dust 2.0 0.55e-6 expdisk {4000 * PC!r} {140 * PC!r} 0 0"]
    for i in incl:
        lines.append(f"instrument frame f{i} {1e7 * PC!r} {float(np.radians(i))!r} 0 0 60 {50000 * PC!r} 60 {50000 * PC!r}")
    S = sr.RefSim("\n".join(lines) + "\n", luminosities=[[1.0]], mixes=common.mix_v()).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    Npp = S.packages_per_lambda()
    engine.set_grid(tables); engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    spiral = dict(arms=2, pitch=float(np.radians(20)), radius=4000 * PC, phase=0.0, weight=1.0, index=1)
    engine.sources([dict(geometry=1, p=[4000 * PC, 350 * PC, 0, 0, 0], spiral=spiral)], L, 0.5)
    engine.instruments([dict(kind=1, distance=1e7 * PC, inclination=float(np.radians(i)), Nxp=60, fovxp=50000 * PC, Nyp=60, fovyp=50000 * PC) for i in incl])
    B = 16
    ref = {i: [] for i in incl}; gpu = {i: [] for i in incl}; ref_l, gpu_l = [], []
    for b in range(B):
        S.reset(900 + 1000 * b); S.run_stellar(); ins = S.instruments()
        engine.reset_results(); st = engine.run_stellar(Npp, store_absorption=True, seed=60 + b)
        for q, i in enumerate(incl):
            ref[i].append(ins[q]["frame"].copy()); gpu[i].append(engine.fetch_frame(q))
        ref_l.append(S.labs().sum()); gpu_l.append(engine.fetch_labs().sum())
    assert st["scatterings"] > 0.5 * st["packets"]                     # forced scattering: tau_V = 2 face-on
    for i in incl:
        out = common.mc_gate(gpu[i], ref[i], f"C3/frame i={i}")
        assert out["bins"] > 200
    common.mc_gate(np.array(gpu_l)[:, None], np.array(ref_l)[:, None], "C3/labs total")


def test_results_snapshot_travels_while_the_engine_goes_on(engine):
    """skg_results_snapshot / skg_fetch_snapshot_async: the shadow copies hold the arrays of the moment of the snapshot, bit for
    bit what the blocking fetches return, also when the engine has reset and refilled its accumulators in the meantime"""
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=5e4, storeabs=1)
    common.setup_engine(engine, cfg, tables, medium, g["L"])
    engine.reset_results()
    engine.run_stellar(5e4, store_absorption=True, seed=77)
    want = {(1, 0): engine.fetch_frame(0).ravel(), (2, 1): engine.fetch_sed(1).ravel(), (0, 0): engine.fetch_labs().ravel()}
    engine.results_snapshot()
    host = {}
    for key in want:
        n = engine.fetch_snapshot_async(key[0], key[1], None)
        assert n == want[key].size
        host[key] = engine.pinned_empty((n,))
        engine.fetch_snapshot_async(key[0], key[1], host[key])
    engine.reset_results()                                   # the engine goes on: the accumulators change under the transfer
    engine.run_stellar(5e4, store_absorption=True, seed=78)
    engine.fetch_snapshot_wait()
    for key in want:
        assert np.array_equal(host[key], want[key]), key
    assert not np.array_equal(engine.fetch_labs().ravel(), want[(0, 0)])
    with pytest.raises(Exception):
        engine.fetch_snapshot_async(9, 0, None)              # no such accumulator
