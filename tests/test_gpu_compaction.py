"""Queue compaction (K3) on the GPU: bundles in which most units die on the first wave, so that the dead slots
become the majority of a >= 4096-slot queue and the order-preserving squeeze runs before the later waves.
The survivors must come out exactly as without it: bit for bit against the oracle for plain rays (live and dead
slots interleaved at random), within the field tolerance for Gaussian triples (3 rays per unit)."""
import numpy as np
import pytest

from tests import scenes
from tests import scenes2 as s2

FIELD_TOL = 1e-8   # detector field, relative L2 (north_star)


@pytest.mark.gpu
def test_compaction_plain_rays_bitwise(bmo, orc):
    n = 16384
    sc, osc = scenes.doublet_spot(bmo), scenes.doublet_spot_oracle()
    pos, d = scenes.fibonacci_disc(n, diameter=40e-3)       # 25.4 mm doublet: ~60 % of the rays miss everything
    perm = np.random.default_rng(7).permutation(n)           # dead and live slots interleaved
    pos = np.ascontiguousarray(pos[perm])
    src = bmo.RayBundle(pos, d, 707e-9)
    res = bmo.solve_system_(sc["system"], src, r_max=100)
    ref = orc.bulk_trace_rays(osc["system"], src.pos, src.dir, 707e-9, max_seg=8, spot=osc["spot"])
    beams, seg = res.beams(), res.segments()
    dead_first_wave = int((ref["nseg"] == 1).sum())
    assert 2 * (n - dead_first_wave) <= n                    # the compaction condition of finish_chunk holds
    assert res.interactions == ref["interactions"]
    assert np.array_equal(beams["nseg"], ref["nseg"])
    rows = np.concatenate([beams["first"][i] + np.arange(beams["nseg"][i]) for i in range(n)])
    rsel = np.concatenate([ref["seg"][i, :beams["nseg"][i]] for i in range(n)])
    assert np.array_equal(seg["pos"][rows], rsel[:, 0:3])
    assert np.array_equal(seg["dir"][rows], rsel[:, 3:6])
    assert np.array_equal(seg["t"][rows], rsel[:, 7])
    hit = ~np.isnan(ref["spot"][:, 0])                       # spot diagram stays in ray order
    assert hit.sum() == sc["spot"].data.shape[0]
    assert np.array_equal(sc["spot"].data, ref["spot"][hit])


@pytest.mark.gpu
def test_compaction_gaussian_triples_field(bmo, orc):
    k, npx = 66, 48                                           # 4356 beamlets on a 90 mm lattice
    sc, osc = s2.expander(bmo, npx), s2.expander_oracle(npx)
    lat = s2.beamlet_lattice(k, aperture=90e-3)
    bundle = bmo.BeamletBundle.from_params(lat["pos"], lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=lat["P0"], support=lat["support"])
    res = bmo.solve_system_(sc["system"], bundle)
    r = np.hypot(lat["pos"][:, 0], lat["pos"][:, 2])
    # non-sequential: a beamlet outside the first lens may still meet the 2" second lens or the 40 mm detector; beyond
    # r = 29 mm (detector corner) nothing is hit at all -- that is the majority of this lattice
    assert 2 * int((r < 29e-3 + 3 * lat["w0"]).sum()) <= k * k
    for p in lat["pos"]:
        g = orc.gaussian_beamlet(p, lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=lat["P0"], support=lat["support"])
        orc.solve_system_(osc["system"], g)
    ref = osc["pd"].pd_field(npx)
    assert np.abs(ref).max() > 0
    err = float(np.linalg.norm((sc["pd"].field - ref).ravel()) / np.linalg.norm(ref.ravel()))
    assert err <= FIELD_TOL, err


@pytest.mark.gpu
def test_late_look_after_a_full_bundle_is_result_identical(bmo, orc):
    """A splitter-free system that saw a bundle in which nobody died enqueues 4 waves before its first look at the device
    (bmo_sys::TraceHint::first_chunk, kept with the uploaded system).  A later bundle of the same size that does lose most
    of its rays on the first wave is then compacted late or not at all: the hit points must not depend on that."""
    n = 8192
    sc, osc = scenes.doublet_spot(bmo), scenes.doublet_spot_oracle()
    dsys = bmo.upload_system(sc["system"], [707e-9])
    lam = np.zeros(n, dtype=np.int32)
    pos, d = scenes.fibonacci_disc(n)                          # fills 20 of the 25.4 mm: every ray reaches the detector
    res0 = bmo.trace_rays(dsys, pos, d, lam)
    assert res0.interactions == 4 * n
    res0.free()
    pos2, _ = scenes.fibonacci_disc(n, diameter=40e-3)
    pos2 = np.ascontiguousarray(pos2[np.random.default_rng(11).permutation(n)])
    ref = orc.bulk_trace_rays(osc["system"], pos2, d, 707e-9, max_seg=8, spot=osc["spot"])
    rsel = np.concatenate([ref["seg"][i, :ref["nseg"][i]] for i in range(n)])
    for _ in range(2):                                         # with the hint left by res0, then with the hint this call leaves
        res = bmo.trace_rays(dsys, pos2, d, lam)
        beams, seg = res.beams(), res.segments()
        assert res.interactions == ref["interactions"]
        assert np.array_equal(beams["nseg"], ref["nseg"])
        rows = np.concatenate([beams["first"][i] + np.arange(beams["nseg"][i]) for i in range(n)])
        assert np.array_equal(seg["pos"][rows], rsel[:, 0:3])
        assert np.array_equal(seg["t"][rows], rsel[:, 7])
        res.free()
