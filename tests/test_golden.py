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


# ---- fixtures of the real Julia reference (baseline/dump_fixtures.jl), used the day they exist --------------------------------
def test_reference_fixture_reader_round_trip(tmp_path):
    """The binary format baseline/dump_fixtures.jl writes (Int64 rank, dims, Float64 column-major data)."""
    from tests.golden import load_reference_fixture as lf
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    with open(tmp_path / "a.bin", "wb") as fh:
        np.array([3, 2, 3, 4], np.int64).tofile(fh); a.ravel(order="F").tofile(fh)
    assert np.array_equal(lf.load(str(tmp_path / "a.bin")), a)
    z = (np.arange(6) + 1j * np.arange(6)[::-1]).reshape(2, 3)
    with open(tmp_path / "z.bin", "wb") as fh:
        np.array([2, 2, 3], np.int64).tofile(fh); np.stack([z.real, z.imag], -1).reshape(2, 3, 2).transpose(1, 0, 2).ravel().tofile(fh)
    assert np.array_equal(lf.load(str(tmp_path / "z.bin"), complex=True), z)


def _ref_c2():
    from tests.golden import load_reference_fixture as lf
    if not lf.available():
        pytest.skip("no fixtures of the Julia reference (tests/golden/ref/, written by baseline/dump_fixtures.jl)")
    seg = lf.load(os.path.join(GOLD, "ref", "c2_segments.bin"))          # (11, 8, n)
    return np.transpose(seg, (2, 1, 0))                                    # (n, 8, 11): pos, dir, n, t, normal


def test_oracle_matches_reference_fixture_c2(orc):
    ref = _ref_c2()
    n = ref.shape[0]
    osc = scenes.doublet_spot_oracle()
    out = orc.bulk_trace_rays(osc["system"], np.ascontiguousarray(ref[:, 0, 0:3]), np.ascontiguousarray(ref[:, 0, 3:6]), 707e-9, max_seg=8, spot=osc["spot"])
    live = np.isfinite(ref[:, :, 0])
    assert np.array_equal(out["nseg"], live.sum(axis=1))
    a, b = out["seg"][:, :, 0:6][live], ref[:, :, 0:6][live]
    assert np.abs(a - b).max() <= POS_TOL * np.abs(b).max()


@pytest.mark.gpu
def test_gpu_matches_reference_fixture_c2(bmo):
    ref = _ref_c2()
    sc = scenes.doublet_spot(bmo)
    res = bmo.solve_system_(sc["system"], bmo.RayBundle(np.ascontiguousarray(ref[:, 0, 0:3]), np.ascontiguousarray(ref[:, 0, 3:6]), 707e-9), r_max=100)
    beams, seg = res.beams(), res.segments()
    for i in range(ref.shape[0]):
        f, k = int(beams["first"][i]), int(beams["nseg"][i])
        assert k == int(np.isfinite(ref[i, :, 0]).sum())
        got = np.concatenate([seg["pos"][f:f + k], seg["dir"][f:f + k]], axis=1)
        assert np.abs(got - ref[i, :k, 0:6]).max() <= POS_TOL * np.abs(ref[i, :k, 0:6]).max()
