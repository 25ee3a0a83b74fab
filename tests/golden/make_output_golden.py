"""Golden vectors for the output side (calibration + wire formats), generated from the reference's OWN
Instrument::write() (SingleFrameInstrument::calibrateAndWriteDataCubes, DistantInstrument::calibrateAndWriteSEDs)
through oracle/_ref, whose Image / TextOutFile stand-ins capture the calibrated data cubes and SED rows.
Run in the build container only:   python tests/golden/make_output_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                   # noqa: E402

PC = common.PC


def case(sim, units):
    ins = [dict(kind=1, name="fr", distance=3.5e6 * PC, inclination=float(np.radians(60)), azimuth=0.3, positionAngle=0.2,
                Nxp=12, fovxp=40000 * PC, Nyp=7, fovyp=30000 * PC),
           dict(kind=2, name="sd", distance=2.0e6 * PC, inclination=float(np.radians(20)), azimuth=0.0, positionAngle=0.0),
           dict(kind=3, name="sm", distance=8.0e6 * PC, inclination=float(np.radians(88)), azimuth=0.0, positionAngle=0.0,
                Nxp=9, fovxp=50000 * PC, Nyp=5, fovyp=20000 * PC)]
    cfg = common.cfg_c1(n=12, packages=2e4, instruments=ins, threads=os.cpu_count() or 1)
    cfg["units"] = units
    if sim == "pan":
        cfg["sim"] = "pan"; cfg["loggrid"] = (0.2e-6, 50e-6, 5); del cfg["wavelengths"]
        cfg["sources"] = [dict(cfg["sources"][0], L=[3.0, 5.0, 2.0, 1.0, 0.5])]
        m = common.mix_v()[0]
        cfg["dust"] = [dict(cfg["dust"][0], mix=tuple(np.repeat(np.asarray(a), 5) for a in m))]
    S = common.make_ref(cfg).setup()
    S.run_stellar()
    raw = S.instruments()
    S.write_instruments()
    out = dict(wavelengths=S.wavelengths(), units=np.array(units))
    for i, d in zip(raw, ins):
        for k in ("kind", "distance", "Nxp", "fovxp", "Nyp", "fovyp"):
            if k in d:
                out[f"{d['name']}_{k}"] = np.array([d[k]])
        if d["kind"] != 2:
            out[d["name"] + "_raw_frame"] = i["frame"]; out[d["name"] + "_cal_frame"] = S.saved_image(d["name"] + "_total")
        if d["kind"] != 1:
            out[d["name"] + "_raw_sed"] = i["sed"]; out[d["name"] + "_cal_sed"] = S.saved_table(d["name"] + "_sed")
    return out


if __name__ == "__main__":
    allc = {}
    for sim in ("oligo", "pan"):
        for us in (0, 1, 2):
            for style in (0, 1, 2):
                for k, v in case(sim, (us, style)).items():
                    allc[f"{sim}_{us}{style}_{k}"] = v
    path = os.path.join(HERE, "output_units.npz")
    np.savez_compressed(path, **allc)
    print(f"{len(allc)} arrays -> {os.path.getsize(path)/1024:.0f} KiB")
