"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py with the oracle): the oracle must still
reproduce them bit for bit (CPU), and the CUDA path must match them through the C ABI like it matches the live oracle (GPU)."""
import os

import numpy as np
import pytest

from tests import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POS_TOL = 1e-9     # hit points / directions, relative (north_star)
FIELD_TOL = 1e-8   # detector field, relative L2 (north_star)


def test_oracle_reproduces_c2_golden(orc):
    g = np.load(os.path.join(GOLD, "c2_doublet_64.npz"))
    osc = scenes.doublet_spot_oracle()
    ref = orc.bulk_trace_rays(osc["system"], g["pos"], g["dir"], 707e-9, max_seg=8, spot=osc["spot"])
    assert ref["interactions"] == int(g["interactions"]) == 4 * 64
    assert np.array_equal(ref["nseg"], g["nseg"])
    assert np.array_equal(ref["seg"], g["seg"], equal_nan=True)
    assert np.array_equal(ref["spot"], g["spot"], equal_nan=True)


def test_oracle_reproduces_c1_golden(orc):
    g = np.load(os.path.join(GOLD, "c1_michelson_32.npz"))
    om = scenes.michelson_oracle(pd_n=32)
    B = scenes.MICHELSON_BEAM
    og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    orc.solve_system_(om["system"], og)
    f = om["pd"].pd_field(32)
    assert np.linalg.norm((f - g["field"]).ravel()) <= 1e-13 * np.linalg.norm(g["field"].ravel())   # libm sincos / exp may differ in the last bit between hosts
    assert abs(om["pd"].pd_power() - float(g["power"])) <= 1e-12 * float(g["power"])


@pytest.mark.gpu
def test_gpu_matches_c2_golden(bmo):
    g = np.load(os.path.join(GOLD, "c2_doublet_64.npz"))
    sc = scenes.doublet_spot(bmo)
    res = bmo.solve_system_(sc["system"], bmo.RayBundle(g["pos"], g["dir"], 707e-9), r_max=100)
    beams, seg = res.beams(), res.segments()
    assert res.interactions == int(g["interactions"])
    assert np.array_equal(beams["nseg"], g["nseg"])
    for i in range(64):
        f, k = int(beams["first"][i]), int(beams["nseg"][i])
        assert np.array_equal(seg["pos"][f:f + k], g["seg"][i, :k, 0:3])      # plain rays: bit for bit (see test_gpu_doublet.py)
        assert np.array_equal(seg["dir"][f:f + k], g["seg"][i, :k, 3:6])
        assert np.array_equal(seg["t"][f:f + k], g["seg"][i, :k, 7])
    assert np.array_equal(sc["spot"].data, g["spot"])


@pytest.mark.gpu
def test_gpu_matches_c1_golden(bmo):
    g = np.load(os.path.join(GOLD, "c1_michelson_32.npz"))
    sc = scenes.michelson(bmo, pd_n=32)
    B = scenes.MICHELSON_BEAM
    bmo.solve_system_(sc["system"], bmo.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"]))
    f = sc["pd"].field
    assert np.linalg.norm((f - g["field"]).ravel()) <= FIELD_TOL * np.linalg.norm(g["field"].ravel())
    assert abs(sc["pd"].optical_power() - float(g["power"])) <= FIELD_TOL * float(g["power"])
