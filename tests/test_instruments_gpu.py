"""GPU parity of the remaining instrument classes.  PerspectiveInstrument (PerspectiveInstrument.cpp:290-350): peel-off towards the eye
of a pinhole camera, optical depth up to the viewport plane, perspective projection.  MultiFrameInstrument (MultiFrameInstrument.cpp:85-99, InstrumentFrame.cpp:153-187): one frame per wavelength,
each with its own pixel grid, recording the total flux and the flux of every stellar component separately -- against runs of
the reference's own MultiFrameInstrument (oracle/_ref), 16 batches on both sides through tests/common.mc_gate."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
PC = common.PC


def _cfg(threads, write_total=True, write_comps=True):
    cfg = common.cfg_c1(n=24, packages=1e5, tau=2.0, threads=threads)
    cfg["wavelengths"] = [0.55e-6, 2.2e-6]
    mv = common.MIX_V
    cfg["dust"][0]["mix"] = ([mv["kabs"], 0.2 * mv["kabs"]], [mv["ksca"], 0.3 * mv["ksca"]], [mv["g"], 0.3])
    cfg["sources"] = [dict(geometry=1, p=[4000 * PC, 350 * PC, 0.0, 0.0, 0.0], L=[1.0, 0.4]),
                      dict(geometry=1, p=[1500 * PC, 600 * PC, 0.0, 0.0, 0.0], L=[0.3, 0.9])]
    frames = [dict(Nxp=40, fovxp=50000 * PC, Nyp=12, fovyp=15000 * PC),
              dict(Nxp=24, fovxp=30000 * PC, Nyp=20, fovyp=24000 * PC, xpc=2000 * PC, ypc=-1000 * PC)]
    cfg["instruments"] = [dict(kind=5, name="mf", distance=1e7 * PC, inclination=float(np.radians(75)), azimuth=0.3, positionAngle=0.2,
                               frames=frames, writeTotal=write_total, writeStellarComps=write_comps),
                          dict(kind=1, name="fr", distance=1e7 * PC, inclination=float(np.radians(75)), azimuth=0.3, positionAngle=0.2,
                               Nxp=40, fovxp=50000 * PC, Nyp=12, fovyp=15000 * PC)]
    return cfg


def test_multiframe_instrument_against_reference_runs(engine):
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    cfg = _cfg(os.cpu_count() or 1)
    S = common.make_ref(cfg).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    common.setup_engine(engine, cfg, tables, medium, L)
    Npp = S.packages_per_lambda()
    B = 16
    keys = [(w, ell) for w in (-1, 0, 1) for ell in (0, 1)]
    ref = {k: [] for k in keys}; gpu = {k: [] for k in keys}
    for b in range(B):
        S.reset(1300 + 1000 * b); S.run_stellar()
        engine.reset_results(); engine.run_stellar(Npp, seed=60 + b)
        for w, ell in keys:
            ref[(w, ell)].append(S.multiframe(0, w, ell).copy()); gpu[(w, ell)].append(engine.fetch_multiframe(0, w, ell).ravel())
        # exact structure of one run: the components add up to the total; frame 0 has the pixel grid of the FrameInstrument in the
        # same direction, which therefore saw the same flux at that wavelength
        for ell in (0, 1):
            np.testing.assert_allclose(gpu[(0, ell)][-1] + gpu[(1, ell)][-1], gpu[(-1, ell)][-1], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(engine.fetch_frame(1).reshape(2, -1)[0], gpu[(-1, 0)][-1], rtol=1e-9, atol=1e-300)
    assert gpu[(-1, 1)][0].size == 24 * 20 and ref[(-1, 1)][0].size == 24 * 20
    for w, ell in keys:
        common.mc_gate(gpu[(w, ell)], ref[(w, ell)], f"multiframe/{'total' if w < 0 else 'stellar_%d' % w}/{ell}", min_bins=0.2)


def test_multiframe_instrument_errors_and_options(engine):
    from skirt_b200.binding import EngineError
    tables, medium, g = common.load_golden_mc()
    cfg = _cfg(1, write_total=False, write_comps=True)
    cfg["wavelengths"] = [0.55e-6]; cfg["sources"] = [dict(s, L=[s["L"][0]]) for s in cfg["sources"]]
    bad = [dict(cfg["instruments"][0])]
    with pytest.raises(EngineError, match="Number of instrument frames must equal number of wavelengths"):
        common.setup_engine(engine, dict(cfg, instruments=bad), tables, medium)          # two frames, one wavelength
    bad = [dict(cfg["instruments"][0], frames=[dict(Nxp=0, fovxp=1.0, Nyp=4, fovyp=1.0)])]
    with pytest.raises(EngineError, match="Number of pixels was not set"):
        common.setup_engine(engine, dict(cfg, instruments=bad), tables, medium)
    ok = [dict(cfg["instruments"][0], frames=cfg["instruments"][0]["frames"][:1])]
    common.setup_engine(engine, dict(cfg, instruments=ok), tables, medium)
    engine.reset_results(); engine.run_stellar(2e4, seed=5)
    with pytest.raises(EngineError, match="does not record the total flux"):
        engine.fetch_multiframe(0, -1, 0)
    a, b = engine.fetch_multiframe(0, 0, 0), engine.fetch_multiframe(0, 1, 0)
    assert a.shape == (12, 40) and a.sum() > 0 and b.sum() > 0
    with pytest.raises(EngineError, match="does not record this stellar component"):
        engine.fetch_multiframe(0, 2, 0)


def _persp_cfg(threads):
    cfg = common.cfg_c1(n=24, packages=1e5, tau=2.0, threads=threads)
    cams = [dict(kind=6, name="outside", Nxp=32, Nyp=24, fovxp=16000 * PC, viewX=15000 * PC, viewY=4000 * PC, viewZ=6000 * PC,
                 crossX=0.0, crossY=0.0, crossZ=0.0, upX=0.0, upY=0.0, upZ=1.0, focal=12000 * PC),
            # a camera INSIDE the dusty disk looking along it: packets behind the viewport are ignored, the optical depth ends at its plane
            dict(kind=6, name="inside", Nxp=24, Nyp=16, fovxp=3000 * PC, viewX=3000 * PC, viewY=500 * PC, viewZ=100 * PC,
                 crossX=-2000 * PC, crossY=0.0, crossZ=0.0, upX=0.0, upY=0.0, upZ=1.0, focal=1500 * PC),
            # looking straight down the z axis: the other branch of the rotation set-up (PerspectiveInstrument.cpp:77-87)
            dict(kind=6, name="pole", Nxp=20, Nyp=20, fovxp=20000 * PC, viewX=0.0, viewY=0.0, viewZ=12000 * PC,
                 crossX=0.0, crossY=0.0, crossZ=0.0, upX=0.0, upY=1.0, upZ=0.0, focal=9000 * PC)]
    cfg["instruments"] = cams + [dict(kind=2, name="sed", distance=1e7 * PC, inclination=float(np.radians(60)))]
    return cfg


def test_perspective_instrument_against_reference_runs(engine):
    from oracle import skirtref as sr
    if not sr.available():
        pytest.skip("oracle/_ref/libskirtref.so not present")
    cfg = _persp_cfg(os.cpu_count() or 1)
    S = common.make_ref(cfg).setup()
    tables, medium, L = S.grid_tables(), S.medium(), S.luminosities()
    common.setup_engine(engine, cfg, tables, medium, L)
    Npp = S.packages_per_lambda()
    B = 16
    ref = {i: [] for i in range(4)}; gpu = {i: [] for i in range(4)}
    for b in range(B):
        S.reset(2300 + 1000 * b); S.run_stellar(); ins = S.instruments()
        engine.reset_results(); engine.run_stellar(Npp, seed=160 + b)
        for i in range(3):
            ref[i].append(ins[i]["frame"].copy()); gpu[i].append(engine.fetch_frame(i).ravel())
        ref[3].append(ins[3]["sed"].copy()); gpu[3].append(engine.fetch_sed(3))
    for i, name in enumerate(("outside", "inside", "pole")):
        assert gpu[i][0].size == ref[i][0].size and np.sum(gpu[i]) > 0
        common.mc_gate(gpu[i], ref[i], f"perspective/{name}", min_bins=0.3)
    common.mc_gate(gpu[3], ref[3], "perspective/sed")


def test_perspective_instrument_errors(engine):
    from skirt_b200.binding import EngineError
    tables, medium, g = common.load_golden_mc()
    cfg = _persp_cfg(1)
    for change, msg in ((dict(fovxp=0.0), "Viewport width was not set"), (dict(upZ=0.0), "Upwards direction was not set"),
                        (dict(focal=0.0), "Focal length was not set"), (dict(crossX=15000 * PC, crossY=4000 * PC, crossZ=6000 * PC), "Crosshair is too close")):
        with pytest.raises(EngineError, match=msg):
            common.setup_engine(engine, dict(cfg, instruments=[dict(cfg["instruments"][0], **change)]), tables, medium)
    common.setup_engine(engine, dict(cfg, instruments=cfg["instruments"][:1]), tables, medium)
    with pytest.raises(EngineError, match="continuous scattering with a PerspectiveInstrument"):
        engine.run_stellar(1e3, seed=1, continuous_scattering=True)
