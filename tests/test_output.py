"""Output side: calibration and wire formats against the reference's own Instrument::write()
(golden vectors from tests/golden/make_output_golden.py: 2 simulation kinds x 3 unit systems x 3 flux styles)."""
import os

import numpy as np
import pytest

import common
from skirt_b200 import output, simulation as sim

G = np.load(os.path.join(common.GOLDEN, "output_units.npz"))
UNITS = {0: output.SIUnits, 1: output.StellarUnits, 2: output.ExtragalacticUnits}
CASES = [(s, u, f) for s in ("oligo", "pan") for u in (0, 1, 2) for f in (0, 1, 2)]


def _grid(simkind, key):
    lam, dlam = G[key + "wavelengths"]             # the reference's lambda and dlambda
    lg = sim.OligoWavelengthGrid(lam) if simkind == "oligo" else sim.LogWavelengthGrid(0.2e-6, 50e-6, 5)
    np.testing.assert_allclose(lg.lambdav, lam, rtol=1e-15)
    np.testing.assert_allclose(lg.dlambdav, dlam, rtol=1e-14)
    return lg


def _instr(key, name):
    d = {k: G[f"{key}{name}_{k}"][0] for k in ("kind", "distance", "Nxp", "fovxp", "Nyp", "fovyp") if f"{key}{name}_{k}" in G.files}
    for k in ("kind", "Nxp", "Nyp"):
        if k in d:
            d[k] = int(d[k])
    return d


@pytest.mark.parametrize("simkind,us,style", CASES)
def test_calibration_matches_reference_write(simkind, us, style):
    key = f"{simkind}_{us}{style}_"
    lg = _grid(simkind, key); units = UNITS[us](style)
    for name in ("fr", "sm"):
        d = _instr(key, name)
        cal = output.calibrate_frames(G[key + name + "_raw_frame"], lg, d, units).ravel()
        ref = G[key + name + "_cal_frame"]
        assert ref.max() > 0
        np.testing.assert_allclose(cal, ref, rtol=4e-15, atol=0)
    for name in ("sd", "sm"):
        d = _instr(key, name)
        cal = output.calibrate_sed(G[key + name + "_raw_sed"], lg, d, units)
        ref = G[key + name + "_cal_sed"]            # rows: lambda, flux
        np.testing.assert_allclose(units.owavelength(lg.lambdav), ref[:, 0], rtol=2e-15)
        np.testing.assert_allclose(cal, ref[:, 1], rtol=4e-15, atol=0)


def test_fits_and_sed_files(tmp_path):
    key = "pan_22_"
    lg = _grid("pan", key); units = output.ExtragalacticUnits(2); d = _instr(key, "fr")
    cube = output.calibrate_frames(G[key + "fr_raw_frame"], lg, d, units)
    path = str(tmp_path / "x_fr_total.fits")
    output.write_fits(path, cube, d["Nxp"], d["Nyp"], lg.Nlambda, units.out("length", d["fovxp"] / d["Nxp"]),
                      units.out("length", d["fovyp"] / d["Nyp"]), 0.0, 0.0, units.unit("surfacebrightness"), units.unit("length"),
                      stamp="2026-01-01T00:00:00")
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T"
    hdr, data = output.read_fits(path)
    assert hdr["BITPIX"] == -32 and hdr["NAXIS"] == 3 and (hdr["NAXIS1"], hdr["NAXIS2"], hdr["NAXIS3"]) == (12, 7, 5)
    assert hdr["BUNIT"] == "MJy/sr" and hdr["CTYPE1"] == "pc" and hdr["CRPIX1"] == 6.5 and hdr["CRPIX2"] == 4.0
    assert hdr["CDELT1"] == pytest.approx(40000 / 12, rel=1e-12)
    np.testing.assert_array_equal(data, cube.astype(np.float32))          # FLOAT_IMG: the reference stores fp32 pixels
    # 2-D image when there is a single wavelength (ffcrim naxis = 2)
    okey = "oligo_00_"; olg = _grid("oligo", okey); ounits = output.SIUnits(0); od = _instr(okey, "fr")
    ocube = output.calibrate_frames(G[okey + "fr_raw_frame"], olg, od, ounits)
    output.write_fits(path, ocube, od["Nxp"], od["Nyp"], 1, 1.0, 1.0, 0.0, 0.0, "W/m2/sr", "m")
    hdr, data = output.read_fits(path)
    assert hdr["NAXIS"] == 2 and "NAXIS3" not in hdr and data.shape == (1, 7, 12)
    # SED text file: header lines + 'e' format with 8 decimals (TextOutFile.cpp:45-85)
    F = output.calibrate_sed(G[key + "sd_raw_sed"], lg, _instr(key, "sd"), units)
    spath = str(tmp_path / "x_sd_sed.dat")
    output.write_sed(spath, lg, [F], ["total flux"], units)
    lines = open(spath).read().splitlines()
    assert lines[0] == "# column 1: lambda (micron)" and lines[1] == "# column 2: total flux; F_nu (Jy)"
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[2:]])
    np.testing.assert_allclose(rows, G[key + "sd_cal_sed"], rtol=1e-8)
    assert all(len(v) >= 14 and "e" in v for v in lines[2].split())
    with pytest.raises(sim.FatalError):
        output.write_fits(path, cube, 3, 3, 3, 1, 1, 0, 0, "x", "y")


# ---- the C++ host layer's Instrument::write() (skirt_b200/host/Output.cpp) through the driver's --write-only mode ----
RUN = os.path.join(common.ROOT, "skirt_b200", "skirt_b200_run")


@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
@pytest.mark.parametrize("simkind,us,style", [("pan", 2, 2), ("pan", 1, 1), ("oligo", 0, 0), ("pan", 2, 0)])
def test_cpp_write_matches_reference(tmp_path, simkind, us, style):
    import subprocess
    key = f"{simkind}_{us}{style}_"
    lam, _ = G[key + "wavelengths"]
    lines = [f"sim {simkind}", "wavelengths " + " ".join(repr(float(v)) for v in lam) if simkind == "oligo" else "loggrid 0.2e-6 50e-6 5",
             f"units {('si', 'stellar', 'extragalactic')[us]} {('neutral', 'wavelength', 'frequency')[style]}"]
    prefix = str(tmp_path / "out")
    for name in ("fr", "sd", "sm"):
        d = _instr(key, name); kind = {1: "frame", 2: "sed", 3: "simple"}[d["kind"]]
        w = f"instrument {kind} {name} {float(d['distance'])!r} 0.5 0 0"
        if kind != "sed":
            w += f" {d['Nxp']} {float(d['fovxp'])!r} {d['Nyp']} {float(d['fovyp'])!r}"
            G[key + name + "_raw_frame"].astype(np.float64).tofile(f"{prefix}_{name}_frame.f64")
        if kind != "frame":
            G[key + name + "_raw_sed"].astype(np.float64).tofile(f"{prefix}_{name}_sed.f64")
        lines.append(w)
    f = tmp_path / "sim.txt"; f.write_text("\n".join(lines) + "\n")
    r = subprocess.run([RUN, "--write-only", str(f), prefix], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    units = UNITS[us](style)
    for name in ("fr", "sm"):
        d = _instr(key, name)
        hdr, cube = output.read_fits(f"{prefix}_{name}_total.fits")
        ref = G[key + name + "_cal_frame"]
        np.testing.assert_array_equal(cube.ravel(), ref.astype(np.float32))       # bit-identical fp32 pixels
        assert hdr["BUNIT"] == units.unit("surfacebrightness") and hdr["CTYPE1"] == units.unit("length")
        assert hdr["NAXIS"] == (2 if len(lam) == 1 else 3) and hdr["CRPIX1"] == (d["Nxp"] + 1) / 2
        assert hdr["CDELT2"] == pytest.approx(units.out("length", d["fovyp"] / d["Nyp"]), rel=1e-14)
        # the Python writer produces the same bytes (apart from the DATE card)
        ppath = str(tmp_path / "py.fits")
        output.write_fits(ppath, output.calibrate_frames(G[key + name + "_raw_frame"], _grid(simkind, key), d, units), d["Nxp"], d["Nyp"], len(lam),
                          units.out("length", d["fovxp"] / d["Nxp"]), units.out("length", d["fovyp"] / d["Nyp"]), 0.0, 0.0,
                          units.unit("surfacebrightness"), units.unit("length"), stamp=hdr["DATE"])
        assert open(ppath, "rb").read() == open(f"{prefix}_{name}_total.fits", "rb").read()
    for name in ("sd", "sm"):
        lines = open(f"{prefix}_{name}_sed.dat").read().splitlines()
        assert lines[0] == f"# column 1: lambda ({units.uwavelength()})"
        assert lines[1] == f"# column 2: total flux; {units.sfluxdensity()} ({units.ufluxdensity()})"
        rows = np.array([[float(v) for v in ln.split()] for ln in lines[2:]])
        np.testing.assert_allclose(rows, G[key + name + "_cal_sed"], rtol=1e-8)


def test_full_instrument_write_matches_reference(tmp_path):
    """FullInstrument::write (FullInstrument.cpp:176-236): total = direct + scattered, one FITS cube per non-empty channel
    (total, direct, scattered, transparent, scatteringlevelN -- no dust cubes without dust emission) and one SED column per
    channel (zeros for the empty dust channels), against the files the reference wrote for the same raw detector arrays"""
    import types
    F = np.load(os.path.join(common.GOLDEN, "mc_full.npz"))
    cfg = common.cfg_full()
    d = cfg["instruments"][0]
    ins = sim.FullInstrument("full", d["distance"], d["inclination"], pixelsX=d["Nxp"], fieldOfViewX=d["fovxp"], pixelsY=d["Nyp"],
                             fieldOfViewY=d["fovyp"], scatteringLevels=d["scatteringLevels"])
    assert ins.channel_names() == ["transparent", "direct", "scattered", "dustdirect", "dustscattered", "scatteringlevel1", "scatteringlevel2"]
    lg = sim.OligoWavelengthGrid(cfg["wavelengths"])
    nf = d["Nxp"] * d["Nyp"]
    res = {}
    for c, cname in enumerate(ins.channel_names()):
        res[f"full_{cname}_frame"] = F[f"last_frame{c}"] if f"last_frame{c}" in F.files else np.zeros(nf)
        res[f"full_{cname}_sed"] = F[f"last_sed{c}"] if f"last_sed{c}" in F.files else np.zeros(1)
    fake = types.SimpleNamespace(lambdagrid=lg, isys=types.SimpleNamespace(instruments=[ins]), ds=object(), dustemission=False)
    out = output.write_instruments(fake, res, str(tmp_path), prefix="t_", units=output.SIUnits(0))
    cubes = sorted(k for k in out if k.endswith(".fits"))
    assert cubes == sorted(f"t_full_{n}.fits" for n in ("total", "direct", "scattered", "transparent", "scatteringlevel1", "scatteringlevel2"))
    for n in ("total", "direct", "scattered", "transparent", "scatteringlevel1", "scatteringlevel2"):
        ref = F["written_" + n]
        assert ref.max() > 0
        np.testing.assert_allclose(out[f"t_full_{n}.fits"].ravel(), ref, rtol=4e-15, atol=0)
        hdr, data = output.read_fits(str(tmp_path / f"t_full_{n}.fits"))
        np.testing.assert_array_equal(data.ravel(), ref.astype(np.float32))
    rows = F["written_sed_rows"]              # lambda + 8 columns: total, direct, scattered, dust, dustscattered, transparent, level 1, level 2
    assert rows.shape == (1, 9)
    np.testing.assert_allclose(out["t_full_sed.dat"][:, 0], rows[0, 1:], rtol=4e-15, atol=0)
    assert rows[0, 4] == 0 and rows[0, 5] == 0
    lines = open(tmp_path / "t_full_sed.dat").read().splitlines()
    assert lines[1].startswith("# column 2: total flux;") and lines[8].startswith("# column 9: 2-times scattered flux;")
    np.testing.assert_allclose([float(v) for v in lines[9].split()], rows[0], rtol=1e-8)
    with pytest.raises(sim.FatalError):
        sim.FullInstrument("x", 1.0, 0.0, pixelsX=2, fieldOfViewX=1.0, pixelsY=2, fieldOfViewY=1.0, scatteringLevels=-1)


@pytest.mark.skipif(not os.path.exists(RUN), reason="skirt_b200_run not built")
def test_cpp_full_instrument_write_matches_reference(tmp_path):
    """the C++ host layer's FullInstrument::write on the reference's raw channel arrays: same cubes, same SED columns"""
    import subprocess
    F = np.load(os.path.join(common.GOLDEN, "mc_full.npz"))
    cfg = common.cfg_full(); d = cfg["instruments"][0]
    prefix = str(tmp_path / "out")
    lines = ["sim oligo", f"wavelengths {cfg['wavelengths'][0]!r}", "units si neutral",
             f"instrument full full {d['distance']!r} {d['inclination']!r} 0 0 {d['Nxp']} {d['fovxp']!r} {d['Nyp']} {d['fovyp']!r} {d['scatteringLevels']}"]
    nf = d["Nxp"] * d["Nyp"]
    for c in range(5 + d["scatteringLevels"]):
        (F[f"last_frame{c}"] if f"last_frame{c}" in F.files else np.zeros(nf)).astype(np.float64).tofile(f"{prefix}_full_frame{c}.f64")
        (F[f"last_sed{c}"] if f"last_sed{c}" in F.files else np.zeros(1)).astype(np.float64).tofile(f"{prefix}_full_sed{c}.f64")
    f = tmp_path / "sim.txt"; f.write_text("\n".join(lines) + "\n")
    r = subprocess.run([RUN, "--write-only", str(f), prefix], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    names = ("total", "direct", "scattered", "transparent", "scatteringlevel1", "scatteringlevel2")
    assert sorted(p for p in os.listdir(tmp_path) if p.endswith(".fits")) == sorted(f"out_full_{n}.fits" for n in names)
    for n in names:
        hdr, cube = output.read_fits(f"{prefix}_full_{n}.fits")
        np.testing.assert_array_equal(cube.ravel(), F["written_" + n].astype(np.float32))
    lines = open(f"{prefix}_full_sed.dat").read().splitlines()
    assert len(lines) == 10 and lines[4] == "# column 5: total dust emission flux; lambda*F_lambda (W/m2)"
    np.testing.assert_allclose([float(v) for v in lines[9].split()], F["written_sed_rows"][0], rtol=1e-8)
