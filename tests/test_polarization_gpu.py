"""Polarisation: Stokes vectors carried by the packets, scattering angles sampled from the Mueller matrix,
the polarised peel-off weight and the Stokes Q, U, V arrays of FullInstruments -- against runs of the reference's own classes
(DustMix.cpp:540-731, StokesVector.cpp, FullInstrument.cpp:107-172, MonteCarloSimulation.cpp:319-363) with the same polarised
dust mix (the Mueller matrix of Thomson scattering, ElectronDustMix.cpp:41-59, on the opacities of the interstellar mix)."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
PC = common.PC


def thomson_mueller(Nlambda, Ntheta=181):
    ct = np.cos(np.arange(Ntheta) * (np.pi / (Ntheta - 1)))
    row = lambda v: np.tile(v, (Nlambda, 1))
    return row(0.5 * (ct * ct + 1.)), row(0.5 * (ct * ct - 1.)), row(ct), row(np.zeros(Ntheta))


def _instruments():
    full = lambda name, inc, pa: dict(kind=4, name=name, distance=1e7 * PC, inclination=float(np.radians(inc)), azimuth=0.0,
                                      positionAngle=float(np.radians(pa)), Nxp=24, fovxp=50000 * PC, Nyp=12, fovyp=25000 * PC, scatteringLevels=1)
    return [full("f80", 80, 0), full("f80pa", 80, 30), full("f20", 20, 0),
            dict(kind=2, name="s80", distance=1e7 * PC, inclination=float(np.radians(80)), azimuth=0.0, positionAngle=0.0)]


def test_polarised_full_instrument_against_reference_runs(engine):
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    ins = _instruments()
    # the reference runs on ONE thread here: DustMix::samplePhi builds its cumulative phi table in a member shared by all
    # threads (const_cast, DustMix.cpp:729), a data race that biases the reference's own multi-threaded polarised runs
    cfg = common.cfg_c1(n=20, packages=1e5, instruments=ins, tau=2.0, threads=1, dustsamples=10)
    mu = thomson_mueller(1)
    S = sr.RefSim(common.ref_spec(cfg), luminosities=[[1.0]], mixes=common.mix_v(), mueller=[mu]).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    Npp = S.packages_per_lambda()
    engine.set_grid(tables); engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    engine.medium_polarization(*[v[None] for v in mu])
    engine.sources(cfg["sources"], L, 0.5); engine.instruments(ins)
    B = 16
    names = ("scattered", "level1", "Q", "U", "V")
    chans = dict(scattered=2, level1=5, Q=6, U=7, V=8)          # 5 + scatteringLevels(1) + {0, 1, 2}: Stokes Q, U, V
    ref = {(i, c): [] for i in range(3) for c in names}; gpu = {(i, c): [] for i in range(3) for c in names}
    ref_sed, gpu_sed = [], []
    for b in range(B):
        S.reset(7000 + 1000 * b); S.run_stellar()
        engine.reset_results(); st = engine.run_stellar(Npp, seed=900 + b)
        for i in range(3):
            nf = ins[i]["Nxp"] * ins[i]["Nyp"]
            for c in names:
                ref[(i, c)].append(S.full_channel(i, chans[c], nf)[0]); gpu[(i, c)].append(engine.fetch_frame_channel(i, chans[c]))
        ref_sed.append(S.instruments()[3]["sed"].copy()); gpu_sed.append(engine.fetch_sed(3))
    assert st["scatterings"] > 0.5 * st["packets"]
    common.mc_gate(np.array(gpu_sed), np.array(ref_sed), "polarised/SED")
    for i in range(3):
        name = ins[i]["name"]
        common.mc_gate(gpu[(i, "scattered")], ref[(i, "scattered")], f"{name}/scattered flux")
        common.mc_gate(gpu[(i, "level1")], ref[(i, "level1")], f"{name}/first scattering level")
        # Stokes Q and U: signed, and their frame totals can cancel by symmetry -> no relative gate on the totals
        q = common.mc_gate(gpu[(i, "Q")], ref[(i, "Q")], f"{name}/Stokes Q", signal=0.3, total_rel=None)
        u = common.mc_gate(gpu[(i, "U")], ref[(i, "U")], f"{name}/Stokes U", signal=0.3, total_rel=None)
        assert q.get("bins", 0) + u.get("bins", 0) > 20, f"{name}: too few pixels with a significant polarised flux"
        # Thomson scattering: S34 = 0 and V starts at 0, so V stays exactly zero on both sides
        assert not np.any(np.array(gpu[(i, "V")])) and not np.any(np.array(ref[(i, "V")]))
    # the degree of linear polarisation of the scattered light at i = 80 deg is substantial and the same on both sides
    Pg = np.hypot(np.mean(gpu[(0, "Q")], 0).sum(), np.mean(gpu[(0, "U")], 0).sum()) / np.mean(gpu[(0, "scattered")], 0).sum()
    Pr = np.hypot(np.mean(ref[(0, "Q")], 0).sum(), np.mean(ref[(0, "U")], 0).sum()) / np.mean(ref[(0, "scattered")], 0).sum()
    assert Pr > 0.02 and abs(Pg / Pr - 1) < 0.05, f"polarisation degree of the scattered flux: gpu {Pg:.4f} reference {Pr:.4f}"
    # a rotated instrument sees the same polarised flux in a rotated frame: Q^2 + U^2 of the total is invariant
    P0 = np.hypot(np.mean(gpu[(0, "Q")], 0).sum(), np.mean(gpu[(0, "U")], 0).sum())
    P1 = np.hypot(np.mean(gpu[(1, "Q")], 0).sum(), np.mean(gpu[(1, "U")], 0).sum())
    assert abs(P1 / P0 - 1) < 0.05


def test_polarisation_setup_errors(engine):
    import skirt_b200 as sk
    tables, medium, g = common.load_golden_mc()
    cfg = common.cfg_c1(n=24, packages=1e3, instruments=_instruments())
    common.setup_engine(engine, cfg, tables, medium, g["L"])           # instruments allocated for an unpolarised medium
    engine.medium_polarization(*[v[None] for v in thomson_mueller(1)])
    with pytest.raises(sk.EngineError, match="call skg_instruments again"):
        engine.run_stellar(1e3, seed=1)
    engine.instruments(cfg["instruments"])
    st = engine.run_stellar(1e3, seed=1)
    assert st["packets"] == 1000
    with pytest.raises(sk.EngineError, match="S11 must be positive"):
        engine.medium_polarization(np.zeros((1, 1, 5)), np.zeros((1, 1, 5)), np.zeros((1, 1, 5)), np.zeros((1, 1, 5)))
    engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])      # a new medium is unpolarised again
    engine.instruments(cfg["instruments"])
    with pytest.raises(sk.EngineError, match="channel out of range"):
        engine.fetch_frame_channel(0, 6)


@pytest.mark.parametrize("polarised", [False, True])
def test_continuous_peel_off_scattering_against_reference_runs(engine, polarised):
    """MonteCarloSimulation::continuouspeeloffscattering (MonteCarloSimulation.cpp:367-434, continuousScattering = true): a
    peel-off packet from a random position inside EVERY segment of every path instead of one per interaction point --
    the same expectation values with less noise.  Against runs of the reference in that mode; and the engine's two modes
    agree with each other on the scattered flux."""
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    ins = _instruments()[:1] + _instruments()[3:]
    cfg = common.cfg_c1(n=16, packages=1e4, instruments=ins, tau=2.0, threads=1 if polarised else (os.cpu_count() or 1), dustsamples=10)
    mu = thomson_mueller(1)
    S = sr.RefSim(common.ref_spec(cfg) + "continuousscattering 1\n", luminosities=[[1.0]], mixes=common.mix_v(), mueller=[mu] if polarised else None).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    Npp = S.packages_per_lambda()
    engine.set_grid(tables); engine.medium(medium["rho"], medium["kext"], medium["ksca"], medium["g"])
    if polarised:
        engine.medium_polarization(*[v[None] for v in mu])
    engine.sources(cfg["sources"], L, 0.5); engine.instruments(ins)
    nf = ins[0]["Nxp"] * ins[0]["Nyp"]
    B = 16
    chans = [2, 5] + ([6] if polarised else [])
    ref = {c: [] for c in chans}; gpu = {c: [] for c in chans}; ref_sed, gpu_sed = [], []
    for b in range(B):
        S.reset(9000 + 1000 * b); S.run_stellar()
        engine.reset_results(); st = engine.run_stellar(Npp, seed=1900 + b, continuous_scattering=True)
        for c in chans:
            ref[c].append(S.full_channel(0, c, nf)[0]); gpu[c].append(engine.fetch_frame_channel(0, c))
        ref_sed.append(S.instruments()[1]["sed"].copy()); gpu_sed.append(engine.fetch_sed(1))
    assert st["paths"] > 20 * st["packets"]          # one peel-off traversal per path segment
    common.mc_gate(np.array(gpu_sed), np.array(ref_sed), "continuous/SED")
    common.mc_gate(gpu[2], ref[2], "continuous/scattered flux")
    common.mc_gate(gpu[5], ref[5], "continuous/first scattering level")
    if polarised:
        common.mc_gate(gpu[6], ref[6], "continuous/Stokes Q", signal=0.3, total_rel=None)
    # the standard mode (peel-off at the interaction points) estimates the same scattered flux -- up to the approximation the
    # continuous mode makes: it places the peel-off point uniformly inside a segment, which differs from the exponential
    # distribution of the interaction points where a cell is optically thick (dtau ~ 1 on this coarse 16^3 grid): a few per cent
    std = []
    for b in range(B):
        engine.reset_results(); engine.run_stellar(Npp * 4, seed=2900 + b)
        std.append(engine.fetch_frame_channel(0, 2).sum())
    cont = np.array(gpu[2]).sum(1)
    assert abs(cont.mean() / np.mean(std) - 1) < 0.05, f"continuous {cont.mean():.6g} vs standard {np.mean(std):.6g}"
