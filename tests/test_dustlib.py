"""The host-side dust emission spectra (GreyBodyDustLib: AllCellsDustLib + GreyBodyDustEmissivity restated in numpy)
against the reference's own DustLib, fed with the same absorbed luminosities."""
import numpy as np
import pytest

import common
from oracle import skirtref as sr


@pytest.mark.skipif(not sr.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_cell_luminosities_match_the_reference_dustlib():
    from oracle import refspec
    from skirt_b200 import configs, simulation as sim
    p = configs.c2_params(n=12, nlambda=30, packages=3e3)
    spec, L, mixes = refspec.reference_spec(p, threads=4, dustsamples=5)
    S = sr.RefSim(spec + "selfabs 1\n", luminosities=L, mixes=mixes).setup()
    S.reset(7); S.run_stellar()
    labs = S.labs()
    Lv_ref = S.prepare_dust(True)                       # [Nlambda, Ncells] = Labsbol * DustLib::luminosity
    lg = configs.wavelength_grid(p)
    lib = sim.GreyBodyDustLib(lg, [mixes[0][0]], S.medium()["rho"], S.volumes())
    Lv = (labs.sum(1)[:, None] * lib.luminosities(labs)).T
    assert Lv_ref.sum() > 0
    np.testing.assert_allclose(Lv.sum(0), Lv_ref.sum(0), rtol=1e-9)           # energy per cell is re-emitted
    big = Lv_ref > 1e-12 * Lv_ref.max()
    np.testing.assert_allclose(Lv[big], Lv_ref[big], rtol=1e-7)
