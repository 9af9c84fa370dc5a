"""C1 parity on the GPU: cube-beamsplitter Michelson (src/Workloads/michelson_wl.jl:8-67) with one
GaussianBeamlet (support fixed).  Detector field within 1e-8 relative L2 of the oracle."""
import numpy as np
import pytest

from tests import scenes


def _tree(beams):
    return [(b["parent"], len(b["chief"]["t"])) for b in beams]


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0.0, 5e-9])
def test_michelson_field_matches_oracle(bmo, orc, shift):
    n = 200
    sc, osc = scenes.michelson(bmo, pd_n=n, m1_shift=shift), scenes.michelson_oracle(pd_n=n, m1_shift=shift)
    B = scenes.MICHELSON_BEAM
    g = bmo.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    res = bmo.solve_system_(sc["system"], g)
    orc.solve_system_(osc["system"], og)
    ref = orc.gauss_export(osc["system"], og)
    # beam tree: same shape in BFS order
    order = res.bfs_order()
    b = res.beams()
    assert len(order) == len(ref)
    assert [int(b["nseg"][i]) for i in order] == [len(r["chief"]["t"]) for r in ref]
    seg = res.segments()
    worst = 0.0
    for i, r in zip(order, ref):
        f, k = int(b["first"][i]), int(b["nseg"][i])
        for lane, key in enumerate(("chief", "waist", "div")):
            rows = (f + np.arange(k)) * 3 + lane
            for name in ("pos", "dir"):
                a, c = seg[name][rows], r[key][name]
                worst = max(worst, float(np.abs(a - c).max() / np.abs(c).max()))
            assert np.array_equal(seg["n"][rows], r[key]["n"])
            assert np.array_equal(np.isinf(seg["t"][rows]), np.isinf(r[key]["t"]))
        assert abs(b["w0"][i] - r["w0"]) <= 1e-12 * r["w0"]
        assert abs(b["E0"][i] - r["E0"]) <= 1e-12 * abs(r["E0"])
    assert worst <= 1e-9, worst
    field = sc["pd"].field
    ofield = osc["pd"].pd_field(n)
    assert np.abs(ofield).max() > 0
    rel = np.linalg.norm((field - ofield).ravel()) / np.linalg.norm(ofield.ravel())
    assert rel <= 1e-8, rel
    I, Io = np.abs(field) ** 2, np.abs(ofield) ** 2
    assert np.linalg.norm((I - Io).ravel()) / np.linalg.norm(Io.ravel()) <= 1e-8
    assert abs(sc["pd"].optical_power() - osc["pd"].pd_power()) <= 1e-8 * abs(osc["pd"].pd_power())
