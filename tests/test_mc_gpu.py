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
