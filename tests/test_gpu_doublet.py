"""C2 parity on the GPU: AC254-150-AB doublet spot diagram through the C ABI vs the CPU oracle.
Tolerance (north_star): hit points and directions within 1e-9 relative."""
import numpy as np
import pytest

from tests import scenes

TOL = 1e-9


def _compare_segments(res, ref, n):
    beams, seg = res.beams(), res.segments()
    assert np.array_equal(beams["nseg"][:n], ref["nseg"])
    worst = 0.0
    for i in range(n):
        f, k = int(beams["first"][i]), int(beams["nseg"][i])
        for name, sl in (("pos", slice(0, 3)), ("dir", slice(3, 6))):
            a, b = seg[name][f:f + k], ref["seg"][i, :k, sl]
            worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
        assert np.array_equal(seg["n"][f:f + k], ref["seg"][i, :k, 6])
        tt, rt = seg["t"][f:f + k], ref["seg"][i, :k, 7]
        assert np.array_equal(np.isinf(tt), np.isinf(rt))
        fin = np.isfinite(rt)
        if fin.any():
            worst = max(worst, float(np.abs(tt[fin] - rt[fin]).max() / np.abs(rt[fin]).max()))
            nn = np.abs(seg["nrm"][f:f + k][fin] - ref["seg"][i, :k, 8:11][fin]).max()
            worst = max(worst, float(nn))
            assert np.array_equal(seg["obj"][f:f + k][fin], ref["seg"][i, :k, 11][fin].astype(int))
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("rotate", [False, True])
def test_doublet_segments_match_oracle(bmo, orc, rotate):
    n = 2048
    sc, osc = scenes.doublet_spot(bmo, rotate), scenes.doublet_spot_oracle(rotate)
    pos, d = scenes.fibonacci_disc(n)
    if rotate:   # same rigid motion for the rays as for the group
        from bmo_b200 import linalg as la
        R = la.matmul(la.rotate3d((0, 0, 1), np.radians(45)), la.rotate3d((1, 0, 0), np.radians(-60)))
        Rm = np.array(R)
        pos = pos @ Rm.T + np.array([0.05, 0.05, 0.05]) @ Rm.T * 0 + np.array(la.matvec(R, (0.0, 0.0, 0.0)))
        # group pivot is its centre (0.05,0.05,0.05 after the translation): rotate about it
        c = np.array([0.05, 0.05, 0.05])
        p0, d0 = scenes.fibonacci_disc(n)
        pos = (p0 @ Rm.T) + c
        d = d0 @ Rm.T
    src = bmo.RayBundle(pos, d, 707e-9)
    res = bmo.solve_system_(sc["system"], src, r_max=100)
    ref = orc.bulk_trace_rays(osc["system"], src.pos, src.dir, 707e-9, max_seg=8, spot=osc["spot"])
    assert res.interactions == ref["interactions"]
    worst = _compare_segments(res, ref, n)
    assert worst <= TOL, worst
    # spot diagram (Spotdetector.data) in ray order
    hit = ~np.isnan(ref["spot"][:, 0])
    assert hit.sum() == sc["spot"].data.shape[0]
    scale = np.abs(ref["spot"][hit]).max()
    assert np.abs(sc["spot"].data - ref["spot"][hit]).max() <= TOL * max(scale, 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("rotate", [False, True])
def test_doublet_bitwise_equal_to_oracle(bmo, orc, rotate):
    """Plain-Ray tracing only uses +, -, *, /, sqrt (IEEE-exact on both sides, no FMA contraction on
    either), so the CUDA path must reproduce the oracle bit for bit -- this pins the "bit-identical"
    claims of the fast SDF evaluation (bmo_geom.cuh) much harder than the 1e-9 tolerance does."""
    n = 4096
    sc, osc = scenes.doublet_spot(bmo, rotate), scenes.doublet_spot_oracle(rotate)
    pos, d = scenes.fibonacci_disc(n, diameter=25.0e-3)      # includes rays at / beyond the lens edge
    if rotate:
        from bmo_b200 import linalg as la
        Rm = np.array(la.matmul(la.rotate3d((0, 0, 1), np.radians(45)), la.rotate3d((1, 0, 0), np.radians(-60))))
        pos, d = pos @ Rm.T + np.array([0.05, 0.05, 0.05]), d @ Rm.T
    src = bmo.RayBundle(pos, d, 707e-9)
    res = bmo.solve_system_(sc["system"], src, r_max=100)
    ref = orc.bulk_trace_rays(osc["system"], src.pos, src.dir, 707e-9, max_seg=8, spot=osc["spot"])
    beams, seg = res.beams(), res.segments()
    assert np.array_equal(beams["nseg"], ref["nseg"])
    rows = np.concatenate([beams["first"][i] + np.arange(beams["nseg"][i]) for i in range(n)])
    rsel = np.concatenate([ref["seg"][i, :beams["nseg"][i]] for i in range(n)])
    assert np.array_equal(seg["pos"][rows], rsel[:, 0:3])
    assert np.array_equal(seg["dir"][rows], rsel[:, 3:6])
    assert np.array_equal(seg["t"][rows], rsel[:, 7])
    fin = np.isfinite(rsel[:, 7])
    assert np.array_equal(seg["nrm"][rows][fin], rsel[:, 8:11][fin])


@pytest.mark.gpu
def test_doublet_single_beam_rebuild(bmo, orc):
    """solve_system!(system, ::Beam): 4 segments, n = [1, n1, n2, 1] (test/runtests.jl:1304-1306)."""
    sc = scenes.doublet_spot(bmo)
    beam = bmo.Beam((1e-3, -0.05, 2e-3), (0.0, 1.0, 0.0), 707e-9)
    bmo.solve_system_(sc["system"], beam)
    assert len(beam.rays) == 4
    assert [r.n for r in beam.rays] == [1.0, scenes.N_NLAK22_707, scenes.N_NSF10_707, 1.0]
    assert beam.rays[0].intersection.object is sc["doublet"]
    assert beam.rays[3].intersection.object is sc["spot"]
    assert sc["spot"].data.shape == (1, 2)
