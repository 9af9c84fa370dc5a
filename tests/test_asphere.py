"""Even-asphere lenses (reference: src/SDFs/AsphericalLensSDF.jl, Lens(::AbstractRotationallySymmetricSurface...)).

CPU: the oracle against the reference's known-answer tests (test/runtests.jl:1529-1696): ray-marched hit
points on the Thorlabs AL50100J surface within 1e-10 of the sag equation, its working distance, and the
three-lens aspherical imaging system (thicknesses, on-axis focus within 1e-7).
GPU: ray bundles through the same lenses against the oracle, segment by segment.
"""
import math

import numpy as np
import pytest

from tests import scenes

POS_TOL = 1e-9
INF = math.inf

AL50100J = dict(R=50.3583e-3, k=-0.789119, A=[0, 2.10405e-7 * (1e3) ** 3, 1.76468e-11 * (1e3) ** 5, 1.02641e-15 * (1e3) ** 7],
                ct=10.2e-3, d=50e-3, n=1.5036)                         # runtests.jl:1533-1544


def _asph(orc, s1, s2, ct, n):
    """s = (r, d, k, coeffs) for an EvenAsphericalSurface, (r, d) for a spherical / flat surface."""
    def part(s):
        if len(s) == 2:
            return [s[0], s[1], s[1], 0.0, -1.0], []
        return [s[0], s[1], s[1], s[2], float(len(s[3]))], list(s[3])
    a, ca = part(s1)
    b, cb = part(s2)
    return orc.new("AsphericLens", a + b + [ct] + ca + cb, [orc.refindex(n)])


def _sag(r, R, k, A):
    c = 1 / R
    return c * r * r / (1 + math.sqrt(1 - (1 + k) * c * c * r * r)) + sum(a * (r * r) ** (i + 1) for i, a in enumerate(A))


def test_oracle_al50100j_surface_and_working_distance(orc):
    L = AL50100J
    lens = _asph(orc, (L["R"], L["d"], L["k"], L["A"]), (INF, L["d"]), L["ct"], L["n"])
    sys_ = orc.system([lens])
    errs = []
    for z in np.linspace(-0.02, 0.02, 100):                                    # runtests.jl:1556-1563
        b = orc.beam([0.0, -0.1, z], [0.0, 1.0, 0.0], 1e-6)
        orc.solve_system_(sys_, b, r_max=40)
        r = orc.beam_export(sys_, b)[0]["rays"]
        hit = r["pos"][0] + r["t"][0] * r["dir"][0]
        errs.append(hit[1] - _sag(z, L["R"], L["k"], L["A"]))
    assert np.abs(errs).max() <= 1e-10                                          # :1567
    b = orc.beam([0.0, -0.1, 0.02], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sys_, b, r_max=40)
    r = orc.beam_export(sys_, b)[0]["rays"]
    pos, d = r["pos"][-1], r["dir"][-1]
    dist = -pos[2] / d[2]
    wd = math.cos(math.asin(d[2])) * dist
    assert abs(wd - 93.2e-3) <= 1e-4                                            # :1578


def _imaging_system(F):
    """runtests.jl:1581-1660"""
    e = 1e3
    L1 = F((1.054e-3, 1.333024e-3, -0.14294, [0, 0.038162 * e ** 3, 0.06317 * e ** 5, -0.020792 * e ** 7, 0.18432 * e ** 9, -0.04827 * e ** 11, 0.094529 * e ** 13]),
           (2.027e-3, 1.216472e-3, 8.0226, [0, 0.0074974 * e ** 3, 0.064686 * e ** 5, 0.19354 * e ** 7, -0.50703 * e ** 9, -0.34529 * e ** 11, 5.9938 * e ** 13]),
           0.72e-3, 1.580200)
    L2 = F((-3.116e-3, 1.4e-3, -49.984, [0, -0.31608 * e ** 3, 0.34755 * e ** 5, -0.17102 * e ** 7, -0.41506 * e ** 9, -1.342 * e ** 11, 5.0594 * e ** 13, -2.7483 * e ** 15]),
           (-4.835e-3, 1.9e-3, 1.6674, [0, -0.079727 * e ** 3, 0.13899 * e ** 5, -0.044057 * e ** 7, -0.019369 * e ** 9, 0.016993 * e ** 11, 0.093716 * e ** 13, -0.080329 * e ** 15]),
           0.55e-3, 1.804700)
    L3 = F((3.618e-3, 3.04e-3, -44.874, [0, -0.14756 * e ** 3, 0.035194 * e ** 5, -0.0032262 * e ** 7, 0.0018592 * e ** 9, 0.00036658 * e ** 11, -0.00016039 * e ** 13, -3.1846e-5 * e ** 15]),
           (2.161e-3, 3.7e-3, -10.719, [0, -0.096568 * e ** 3, 0.026771 * e ** 5, -0.011261 * e ** 7, 0.0019879 * e ** 9, 0.00015579 * e ** 11, -0.00012433 * e ** 13, 1.5264e-5 * e ** 15]),
           0.7e-3, 1.580200)
    Filt = F((INF, 4.2e-3), (INF, 4.2e-3), 0.15e-3, 1.516800)
    Cover = F((INF, 4.9e-3), (INF, 4.9e-3), 0.5e-3, 1.469200)
    return L1, L2, L3, Filt, Cover


def _place(lenses, thickness, position):
    L1, L2, L3, Filt, Cover = lenses
    L2.translate3d_([0, thickness(L1) + 0.39e-3, 0])
    L3.translate_to3d_(position(L2)); L3.translate3d_([0, thickness(L2) + 0.63e-3, 0])
    Filt.translate_to3d_(position(L3)); Filt.translate3d_([0, thickness(L3) + 0.19e-3, 0])
    Cover.translate_to3d_(position(Filt)); Cover.translate3d_([0, thickness(Filt) + 0.18e-3, 0])


def test_oracle_aspherical_imaging_system(orc):
    lenses = _imaging_system(lambda s1, s2, ct, n: _asph(orc, s1, s2, ct, n))
    th = lambda l: float(l.eval("thickness_object", nout=1)[0])
    _place(lenses, th, lambda l: list(l.position()))
    for l, t in zip(lenses, (0.72e-3, 0.55e-3, 0.7e-3, 0.15e-3, 0.5e-3)):       # runtests.jl:1663-1667
        assert abs(th(l) - t) <= 1.5e-8 * t
    sys_ = orc.system(list(lenses))
    for z in (-1.3e-3 / 2, 0.0, 1.3e-3 / 2):                                    # :1672-1683
        b = orc.beam([0.0, -0.5e-3, z], [0.0, 1.0, 0.0], 0.5876e-6)
        orc.solve_system_(sys_, b, r_max=50)
        r = orc.beam_export(sys_, b)[0]["rays"]
        f_pos = r["pos"][-1] + 0.12e-3 * r["dir"][-1]
        assert abs(f_pos[2]) <= 1e-7


# ---- GPU ------------------------------------------------------------------------------------------
def _surf(bmo, s):
    return bmo.SphericalSurface(s[0], s[1]) if len(s) == 2 else bmo.EvenAsphericalSurface(s[0], s[1], s[2], s[3])


def _compare_bundle(bmo, orc, sys_, osys, pos, d, lam, r_max):
    res = bmo.solve_system_(sys_, bmo.RayBundle(pos, d, lam), r_max=r_max)
    b, seg = res.beams(), res.segments()
    ref = orc.bulk_trace_rays(osys, pos, d, lam, r_max=r_max, max_seg=64)
    assert np.array_equal(b["nseg"], ref["nseg"])
    worst = 0.0
    for i in range(pos.shape[0]):
        f0, k = int(b["first"][i]), int(b["nseg"][i])
        got = np.concatenate([seg["pos"][f0:f0 + k], seg["dir"][f0:f0 + k]], axis=1)
        worst = max(worst, float(np.abs(got - ref["seg"][i, :k, 0:6]).max()))
        assert np.array_equal(seg["n"][f0:f0 + k], ref["seg"][i, :k, 6])
    return worst, b


@pytest.mark.gpu
def test_gpu_al50100j_matches_oracle(bmo, orc):
    L = AL50100J
    s1, s2 = (L["R"], L["d"], L["k"], L["A"]), (INF, L["d"])
    lens = bmo.LensFromSurfaces(_surf(bmo, s1), L["ct"], L["n"])
    olens = _asph(orc, s1, s2, L["ct"], L["n"])
    for x in (lens, olens):
        x.xrotate3d_(0.05); x.zrotate3d_(-0.08); x.translate3d_([1e-3, 0.0, -2e-3])
    rng = np.random.default_rng(2)
    n = 256
    pos = np.zeros((n, 3)); pos[:, 0] = rng.uniform(-0.03, 0.03, n); pos[:, 2] = rng.uniform(-0.03, 0.03, n); pos[:, 1] = -0.1
    d = np.tile([0.0, 1.0, 0.0], (n, 1)) + 0.03 * rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    worst, b = _compare_bundle(bmo, orc, bmo.System([lens]), orc.system([olens]), pos, d, 1.31e-6, 40)
    assert (b["nseg"] == 3).sum() > n // 3          # most rays go through both faces
    assert worst <= POS_TOL, worst


@pytest.mark.gpu
def test_gpu_aspherical_imaging_system_matches_oracle(bmo, orc):
    olenses = _imaging_system(lambda s1, s2, ct, n: _asph(orc, s1, s2, ct, n))
    lenses = _imaging_system(lambda s1, s2, ct, n: bmo.LensFromSurfaces(_surf(bmo, s1), _surf(bmo, s2), ct, n))
    _place(olenses, lambda l: float(l.eval("thickness_object", nout=1)[0]), lambda l: list(l.position()))
    _place(lenses, lambda l: l.thickness(), lambda l: list(l.position()))
    for a, t in zip(lenses, (0.72e-3, 0.55e-3, 0.7e-3, 0.15e-3, 0.5e-3)):
        assert abs(a.thickness() - t) <= 1.5e-8 * t
    sys_, osys = bmo.System(list(lenses)), orc.system(list(olenses))
    # the reference's three on-axis-field rays focus on the axis (runtests.jl:1672-1683)
    for z in (-1.3e-3 / 2, 0.0, 1.3e-3 / 2):
        beam = bmo.Beam(bmo.Ray((0.0, -0.5e-3, z), (0.0, 1.0, 0.0), 0.5876e-6))
        bmo.solve_system_(sys_, beam, r_max=50)
        last = beam.rays[-1]
        assert abs(last.pos[2] + 0.12e-3 * last.dir[2]) <= 1e-7
    rng = np.random.default_rng(4)
    n = 128
    pos = np.zeros((n, 3)); pos[:, 0] = rng.uniform(-0.6e-3, 0.6e-3, n); pos[:, 2] = rng.uniform(-0.6e-3, 0.6e-3, n); pos[:, 1] = -0.5e-3
    d = np.tile([0.0, 1.0, 0.0], (n, 1)) + 0.05 * rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    worst, b = _compare_bundle(bmo, orc, sys_, osys, pos, d, 0.5876e-6, 50)
    assert (b["nseg"] >= 11).sum() > n // 4         # through all five elements
    assert worst <= POS_TOL, worst
