"""Output side of the instruments: calibration of the detector arrays and the two wire formats the reference writes.

Mirrors, on the host arrays the engine fills (skg_fetch_frame / skg_fetch_sed):
  SingleFrameInstrument::calibrateAndWriteDataCubes   SingleFrameInstrument.cpp:151-226
  DistantInstrument::calibrateAndWriteSEDs            DistantInstrument.cpp:131-183
  Units (SI / stellar / extragalactic, flux output styles)   Units.cpp:30-215,495-506,765-1040; SIUnits.cpp,
                                                      StellarUnits.cpp, ExtragalacticUnits.cpp
  FITSInOut::write (FLOAT_IMG data cube + keywords)   FITSInOut.cpp:32-90
  TextOutFile (column header lines, 'e' rows)         TextOutFile.cpp:45-85
"""
import datetime
import math
import os

import numpy as np

from .simulation import FatalError, INSTR_FRAME, INSTR_SED, INSTR_FULL, INSTR_MULTIFRAME, INSTR_PERSPECTIVE

_C = 2.99792458e8            # Units.cpp:17-22
_AU = 1.49597871e11
_PC = 3.08567758e16
_ARCSEC2 = (math.pi / (180. * 3600.)) ** 2

# conversion factors to SI of the units the instruments write in (Units.cpp:53-162)
_FACTOR = {
    "length m": 1., "length AU": _AU, "length pc": _PC,
    "distance m": 1., "distance pc": _PC, "distance Mpc": 1e6 * _PC,
    "wavelength m": 1., "wavelength micron": 1e-6,
    "neutralfluxdensity W/m2": 1.,
    "neutralsurfacebrightness W/m2/sr": 1., "neutralsurfacebrightness W/m2/arcsec2": 1. / _ARCSEC2,
    "wavelengthfluxdensity W/m3": 1., "wavelengthfluxdensity W/m2/micron": 1e6,
    "wavelengthsurfacebrightness W/m3/sr": 1., "wavelengthsurfacebrightness W/m2/micron/arcsec2": 1e6 / _ARCSEC2,
    "frequencyfluxdensity W/m2/Hz": 1., "frequencyfluxdensity Jy": 1e-26,
    "frequencysurfacebrightness W/m2/Hz/sr": 1., "frequencysurfacebrightness MJy/sr": 1e-20,
}


class Units:
    """Units (Units.hpp:66-74): a unit per physical quantity + the flux output style"""
    Neutral, Wavelength, Frequency = 0, 1, 2
    _unitForQty = {}

    def __init__(self, fluxOutputStyle=0):
        if fluxOutputStyle not in (0, 1, 2):
            raise FatalError("Unknown flux output style")
        self.fluxOutputStyle = fluxOutputStyle

    def unit(self, qty):
        if qty == "fluxdensity":
            qty = ("neutral", "wavelength", "frequency")[self.fluxOutputStyle] + "fluxdensity"
        if qty == "surfacebrightness":
            qty = ("neutral", "wavelength", "frequency")[self.fluxOutputStyle] + "surfacebrightness"
        if qty not in self._unitForQty:
            raise FatalError("Unknown quantity " + qty)
        return self._unitForQty[qty]

    def _c(self, qty):
        return _FACTOR[qty + " " + self._unitForQty[qty]]

    def out(self, qty, value):
        """Units::out: SI value -> value in the unit of this system"""
        return value / self._c(qty)

    def owavelength(self, lam):
        return lam / self._c("wavelength")

    def sfluxdensity(self):
        return ("lambda*F_lambda", "F_lambda", "F_nu")[self.fluxOutputStyle]

    def ufluxdensity(self):
        return self.unit("fluxdensity")

    def uwavelength(self):
        return self._unitForQty["wavelength"]

    def ofluxdensity(self, lam, Flambda):
        """Units::ofluxdensity, Units.cpp:995-1004"""
        if self.fluxOutputStyle == self.Wavelength:
            return Flambda / self._c("wavelengthfluxdensity")
        if self.fluxOutputStyle == self.Frequency:
            return (lam * lam * Flambda / _C) / self._c("frequencyfluxdensity")
        return (lam * Flambda) / self._c("neutralfluxdensity")

    def osurfacebrightness(self, lam, flambda):
        """Units::osurfacebrightness, Units.cpp:1033-1041"""
        if self.fluxOutputStyle == self.Wavelength:
            return flambda / self._c("wavelengthsurfacebrightness")
        if self.fluxOutputStyle == self.Frequency:
            return (lam * lam * flambda / _C) / self._c("frequencysurfacebrightness")
        return (lam * flambda) / self._c("neutralsurfacebrightness")


class SIUnits(Units):
    _unitForQty = {"length": "m", "distance": "m", "wavelength": "m", "neutralfluxdensity": "W/m2",
                   "neutralsurfacebrightness": "W/m2/sr", "wavelengthfluxdensity": "W/m3",
                   "wavelengthsurfacebrightness": "W/m3/sr", "frequencyfluxdensity": "W/m2/Hz",
                   "frequencysurfacebrightness": "W/m2/Hz/sr"}


class StellarUnits(Units):
    _unitForQty = {"length": "AU", "distance": "pc", "wavelength": "micron", "neutralfluxdensity": "W/m2",
                   "neutralsurfacebrightness": "W/m2/arcsec2", "wavelengthfluxdensity": "W/m2/micron",
                   "wavelengthsurfacebrightness": "W/m2/micron/arcsec2", "frequencyfluxdensity": "Jy",
                   "frequencysurfacebrightness": "MJy/sr"}


class ExtragalacticUnits(StellarUnits):
    _unitForQty = dict(StellarUnits._unitForQty, length="pc", distance="Mpc")


# ---- calibration -----------------------------------------------------------------------------------------------
def calibrate_frames(frames, lambdagrid, d, units):
    """SingleFrameInstrument::calibrateAndWriteDataCubes (:151-212) for one raw data cube [Nlambda, Nyp, Nxp]
    (bolometric luminosity per pixel, W) -> surface brightness in the output units; same order of operations"""
    f = np.array(frames, dtype=np.float64).reshape(lambdagrid.Nlambda, d["Nyp"], d["Nxp"])
    # step 1: W -> W/m
    f = f / lambdagrid.dlambdav[:, None, None]
    # step 2: per steradian (the area of a pixel on the sky)
    xpsiz = d["fovxp"] / d["Nxp"]; ypsiz = d["fovyp"] / d["Nyp"]
    area = (2.0 * math.atan(xpsiz / (2.0 * d["distance"]))) * (2.0 * math.atan(ypsiz / (2.0 * d["distance"])))
    f = f / area
    # step 3: flux density at the distance of the observer
    f = f / (4.0 * math.pi * d["distance"] * d["distance"])
    # output units
    lam = lambdagrid.lambdav[:, None, None]
    return units.osurfacebrightness(lam, f)


def calibrate_sed(sed, lambdagrid, d, units):
    """DistantInstrument::calibrateAndWriteSEDs (:139-156, :176-178): raw SED (W per bin) -> flux density in output units"""
    F = np.array(sed, dtype=np.float64) / lambdagrid.dlambdav
    F = F / (4.0 * math.pi * d["distance"] * d["distance"])
    return units.ofluxdensity(lambdagrid.lambdav, F)


# ---- wire formats ------------------------------------------------------------------------------------------------
def _card(key, value=None, comment=""):
    """one 80-character FITS header card in cfitsio's fixed format (ffpky / ffpkys)"""
    if value is None:
        s = key
    elif isinstance(value, bool):
        s = f"{key:<8}= {'T' if value else 'F':>20}"
    elif isinstance(value, int):
        s = f"{key:<8}= {value:>20d}"
    elif isinstance(value, float):
        # cfitsio writes TDOUBLE keywords with 15 decimals in %E style (ffd2e), shortened when exact
        txt = f"{value:.15G}"
        if "E" not in txt and "." not in txt:
            txt += "."
        s = f"{key:<8}= {txt:>20}"
    else:
        q = "'" + str(value).replace("'", "''").ljust(8) + "'"
        s = f"{key:<8}= {q:<20}"
    if comment and value is not None:
        s += " / " + comment
    return s[:80].ljust(80)


def write_fits(path, data, nx, ny, nz, incx, incy, xc, yc, dataunits, xyunits, stamp=None):
    """FITSInOut::write (FITSInOut.cpp:32-90): primary HDU with a FLOAT_IMG (BITPIX -32, big-endian) cube of
    nx x ny x nz pixels (2-D when nz == 1) and the reference's keywords, in 2880-byte blocks"""
    data = np.asarray(data, dtype=np.float64).ravel()
    if data.size != nx * ny * nz:
        raise FatalError("Inconsistent data size when creating FITS file " + path)
    if stamp is None:
        stamp = datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S")
    cards = [_card("SIMPLE", True, "file does conform to FITS standard"),
             _card("BITPIX", -32, "number of bits per data pixel"),
             _card("NAXIS", 2 if nz == 1 else 3, "number of data axes"),
             _card("NAXIS1", int(nx), "length of data axis 1"), _card("NAXIS2", int(ny), "length of data axis 2")]
    if nz != 1:
        cards.append(_card("NAXIS3", int(nz), "length of data axis 3"))
    cards += [_card("EXTEND", True, "FITS dataset may contain extensions"),
              _card("COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy"),
              _card("COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H"),
              _card("BSCALE", 1.0), _card("BZERO", 0.0),
              _card("DATE", stamp, "Date and time of creation (UTC)"),
              _card("ORIGIN", "SKIRT simulation", "Astronomical Observatory, Ghent University"),
              _card("BUNIT", dataunits, "Physical unit of the array values"),
              _card("CRPIX1", (nx + 1.0) / 2.0, "X-axis coordinate system reference pixel"),
              _card("CRVAL1", float(xc), "Coordinate system value at X-axis reference pixel"),
              _card("CDELT1", float(incx), "Coordinate increment along X-axis"),
              _card("CTYPE1", xyunits, "Physical units of the X-axis increment"),
              _card("CRPIX2", (ny + 1.0) / 2.0, "Y-axis coordinate system reference pixel"),
              _card("CRVAL2", float(yc), "Coordinate system value at Y-axis reference pixel"),
              _card("CDELT2", float(incy), "Coordinate increment along Y-axis"),
              _card("CTYPE2", xyunits, "Physical units of the Y-axis increment"),
              "END".ljust(80)]
    header = "".join(cards)
    header += " " * (-len(header) % 2880)
    body = data.astype(">f4").tobytes()
    body += b"\0" * (-len(body) % 2880)
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii")); fh.write(body)


def read_fits(path):
    """minimal reader for the files write_fits / the reference produce: (header dict, float32 cube [nz, ny, nx])"""
    raw = open(path, "rb").read()
    hdr = {}; pos = 0
    while True:
        card = raw[pos:pos + 80].decode("ascii"); pos += 80
        key = card[:8].strip()
        if key == "END":
            break
        if card[8:10] == "= ":
            v = card[10:].split(" / ")[0].strip()
            if v.startswith("'"):
                hdr[key] = v.strip("'").rstrip()
            elif v in ("T", "F"):
                hdr[key] = v == "T"
            else:
                hdr[key] = float(v) if any(c in v for c in ".E") else int(v)
    pos += -pos % 2880
    nx, ny, nz = hdr["NAXIS1"], hdr["NAXIS2"], hdr.get("NAXIS3", 1)
    cube = np.frombuffer(raw, dtype=">f4", count=nx * ny * nz, offset=pos).reshape(nz, ny, nx)
    return hdr, cube


def _qnum_e(v, prec):
    """QString::number(v, 'e', prec): C printf %.<prec>e"""
    return f"{v:.{prec}e}"


def full_instrument_arrays(ins, results, dustsystem=True, dustemission=False):
    """the lists FullInstrument::write assembles (FullInstrument.cpp:176-236): [(file name, column name, cube, sed)] with
    total = direct + scattered (+ dust), empty arrays left out like the reference's `if (farr->size())`"""
    get = lambda c, k: np.asarray(results[f"{ins.name}_{c}_{k}"], dtype=np.float64)
    fr = {c: get(c, "frame") for c in ins.channel_names()}; se = {c: get(c, "sed") for c in ins.channel_names()}
    rows = []
    if dustemission:
        rows.append(("total", "total flux", fr["direct"] + fr["scattered"] + fr["dustdirect"] + fr["dustscattered"],
                     se["direct"] + se["scattered"] + se["dustdirect"] + se["dustscattered"]))
    elif dustsystem:
        rows.append(("total", "total flux", fr["direct"] + fr["scattered"], se["direct"] + se["scattered"]))
    else:
        rows.append(("total", "total flux", fr["transparent"], se["transparent"]))
    if dustsystem:
        rows.append(("direct", "direct stellar flux", fr["direct"], se["direct"]))
        rows.append(("scattered", "scattered stellar flux", fr["scattered"], se["scattered"]))
    else:
        rows.append(("direct", "direct stellar flux", None, se["transparent"]))
        rows.append(("scattered", "scattered stellar flux", None, None))
    if dustemission:
        rows.append(("dust", "total dust emission flux", fr["dustdirect"] + fr["dustscattered"], se["dustdirect"] + se["dustscattered"]))
        rows.append(("dustscattered", "dust emission scattered flux", fr["dustscattered"], se["dustscattered"]))
    else:
        rows.append(("dust", "total dust emission flux", None, None))
        rows.append(("dustscattered", "dust emission scattered flux", None, None))
    rows.append(("transparent", "transparent flux", fr["transparent"] if dustsystem else None, se["transparent"]))
    for n in range(ins.d["scatteringLevels"]):
        c = f"scatteringlevel{n + 1}"
        rows.append((c, f"{n + 1}-times scattered flux", fr[c] if dustsystem else None, se[c] if dustsystem else None))
    return rows


def write_sed(path, lambdagrid, columns, names, units):
    """the <instrument>_sed.dat text file (DistantInstrument.cpp:160-182 through TextOutFile.cpp:45-85)"""
    lines = [f"# column 1: lambda ({units.uwavelength()})"]
    for q, nm in enumerate(names):
        lines.append(f"# column {q + 2}: {nm}; {units.sfluxdensity()} ({units.ufluxdensity()})")
    for ell in range(lambdagrid.Nlambda):
        vals = [units.owavelength(lambdagrid.lambdav[ell])] + [float(c[ell]) for c in columns]
        lines.append(" ".join(_qnum_e(v, 8) for v in vals))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def write_instruments(sim, results, outdir, prefix="", units=None, stamp=None):
    """Instrument::write() for every instrument of a finished simulation: calibrates the (already reduced) detector
    arrays and writes <prefix><name>_total.fits and <prefix><name>_sed.dat like FrameInstrument::write
    (FrameInstrument.cpp:51-66), SEDInstrument::write (SEDInstrument.cpp:46-61) and SimpleInstrument::write
    (SimpleInstrument.cpp:53-74).  Returns {file name: calibrated array}."""
    units = units or SIUnits()
    os.makedirs(outdir, exist_ok=True)
    out = {}
    lg = sim.lambdagrid
    for ins in sim.isys.instruments:
        d = ins.d
        if ins.kind == INSTR_FULL:
            rows = full_instrument_arrays(ins, results, dustsystem=sim.ds is not None, dustemission=bool(getattr(sim, "dustemission", False)))
            xpsiz = d["fovxp"] / d["Nxp"]; ypsiz = d["fovyp"] / d["Nyp"]
            cols, names = [], []
            for fname, cname, cube, sed in rows:
                if cube is not None:
                    cal = calibrate_frames(cube, lg, d, units)
                    name = f"{prefix}{ins.name}_{fname}.fits"
                    write_fits(os.path.join(outdir, name), cal, d["Nxp"], d["Nyp"], lg.Nlambda, units.out("length", xpsiz),
                               units.out("length", ypsiz), d["xpc"], d["ypc"], units.unit("surfacebrightness"), units.unit("length"), stamp)
                    out[name] = cal
                # an empty F-array still gets a column of zeros (DistantInstrument.cpp:178)
                cols.append(calibrate_sed(sed, lg, d, units) if sed is not None else np.zeros(lg.Nlambda)); names.append(cname)
            name = f"{prefix}{ins.name}_sed.dat"
            write_sed(os.path.join(outdir, name), lg, cols, names, units)
            out[name] = np.array(cols)
            continue
        if ins.kind == INSTR_MULTIFRAME:
            # InstrumentFrame::calibrateAndWriteDataFrames (InstrumentFrame.cpp:216-262): one FITS file per array and wavelength
            for ell, arrays in enumerate(results[ins.name + "_frames"]):
                fd = d["frames"][ell]
                xpsiz = fd["fovxp"] / fd["Nxp"]; ypsiz = fd["fovyp"] / fd["Nyp"]
                area = (2.0 * math.atan(xpsiz / (2.0 * d["distance"]))) * (2.0 * math.atan(ypsiz / (2.0 * d["distance"])))
                fourpid2 = 4.0 * math.pi * d["distance"] * d["distance"]
                unitfactor = units.osurfacebrightness(lg.lambdav[ell], 1.0)
                for fname, raw in arrays.items():
                    cal = np.asarray(raw, dtype=np.float64) * (unitfactor / (lg.dlambdav[ell] * area * fourpid2))
                    name = f"{prefix}{ins.name}_{fname}_{ell}.fits"
                    write_fits(os.path.join(outdir, name), cal, fd["Nxp"], fd["Nyp"], 1, units.out("length", xpsiz), units.out("length", ypsiz),
                               fd["xpc"], fd["ypc"], units.unit("surfacebrightness"), units.unit("length"), stamp)
                    out[name] = cal
            continue
        if ins.kind == INSTR_PERSPECTIVE:
            # PerspectiveInstrument::write (PerspectiveInstrument.cpp:354-397): every sample times 1/(4 pi s^2) / dlambda, to output units
            sp = d["fovxp"] / d["Nxp"]
            raw = np.array(results[ins.name + "_frame"], dtype=np.float64).reshape(lg.Nlambda, d["Nyp"], d["Nxp"])
            front = 1.0 / (4.0 * math.pi * sp * sp)
            cube = units.osurfacebrightness(lg.lambdav[:, None, None], raw * front / lg.dlambdav[:, None, None])
            name = f"{prefix}{ins.name}_total.fits"
            write_fits(os.path.join(outdir, name), cube, d["Nxp"], d["Nyp"], lg.Nlambda, units.out("length", sp), units.out("length", sp), 0.0, 0.0,
                       units.unit("surfacebrightness"), units.unit("length"), stamp)
            out[name] = cube
            continue
        if ins.kind != INSTR_SED:
            cube = calibrate_frames(results[ins.name + "_frame"], lg, d, units)
            xpsiz = d["fovxp"] / d["Nxp"]; ypsiz = d["fovyp"] / d["Nyp"]
            name = f"{prefix}{ins.name}_total.fits"
            write_fits(os.path.join(outdir, name), cube, d["Nxp"], d["Nyp"], lg.Nlambda, units.out("length", xpsiz),
                       units.out("length", ypsiz), d["xpc"], d["ypc"], units.unit("surfacebrightness"), units.unit("length"), stamp)
            out[name] = cube
        if ins.kind != INSTR_FRAME:
            F = calibrate_sed(results[ins.name + "_sed"], lg, d, units)
            name = f"{prefix}{ins.name}_sed.dat"
            write_sed(os.path.join(outdir, name), lg, [F], ["total flux"], units)
            out[name] = F
    return out
