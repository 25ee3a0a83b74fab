"""Golden batch statistics for FullInstrument (per-channel data cubes and SEDs) from the reference's own code
(oracle/_ref).  Run in the build container only:   python tests/golden/make_full_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common                                   # noqa: E402


if __name__ == "__main__":
    batches = 16
    S = common.make_ref(common.cfg_full(threads=os.cpu_count() or 1)).setup()
    t = S.grid_tables(); med = S.medium()
    nframe = 40 * 16 * S.Nlambda
    chans = {c: ([], []) for c in (0, 1, 2, 5, 6)}
    tot = []
    for b in range(batches):
        S.reset(7000 + 17 * b); S.run_stellar()
        for c in chans:
            f, s = S.full_channel(0, c, nframe); chans[c][0].append(f); chans[c][1].append(s)
        tot.append(S.instruments()[1]["sed"].copy())
    out = dict(Npp=np.array([S.packages_per_lambda()]), L=S.luminosities(), batches=np.array([batches]))
    for c, (fr, se) in chans.items():
        fr = np.array(fr); se = np.array(se)
        out[f"frame{c}_mean"] = fr.mean(0); out[f"frame{c}_sem"] = fr.std(0, ddof=1) / np.sqrt(batches)
        out[f"sed{c}_mean"] = se.mean(0); out[f"sed{c}_sem"] = se.std(0, ddof=1) / np.sqrt(batches)
    tot = np.array(tot); out["sedtotal_mean"] = tot.mean(0); out["sedtotal_sem"] = tot.std(0, ddof=1) / np.sqrt(batches)
    for key, v in t.items():
        out["grid_" + key] = np.asarray(v)
    for key, v in med.items():
        out["med_" + key] = v
    # the calibrated files of the last batch (FullInstrument::write): names and SED columns
    S.write_instruments()
    out["written_sed_rows"] = S.saved_table("full_sed")
    for nm in ("total", "direct", "scattered", "transparent", "scatteringlevel1", "scatteringlevel2"):
        out["written_" + nm] = S.saved_image("full_" + nm)
    raw = {c: S.full_channel(0, c, nframe) for c in chans}      # write() calibrated these in place: keep the last raw batch instead
    for c in chans:
        out[f"last_frame{c}"] = chans[c][0][-1]; out[f"last_sed{c}"] = chans[c][1][-1]
    path = os.path.join(HERE, "mc_full.npz")
    np.savez_compressed(path, **out)
    print({k: float(np.sum(v)) for k, v in out.items() if k.startswith("sed") and k.endswith("_mean")})
    print(f"mc_full: {os.path.getsize(path)/1024:.0f} KiB")
