"""Cylindrical lenses (reference: src/SDFs/CylindricalSDF.jl, Lens(::AbstractCylindricalSurface, ...) in
src/OpticalComponents/Lenses.jl:331-414).

CPU: the oracle against the reference's known-answer test for the Thorlabs LJ1878L2 / LK1900L1 lenses
(test/runtests.jl:1698-1743): centre thickness, edge thickness, working distance.
GPU: ray bundles through cylindrical lenses against the oracle, segment by segment.
"""
import math

import numpy as np
import pytest

from tests import scenes

POS_TOL = 1e-9


class _OF(scenes._OracleFactory):
    def CylindricalLens(self, r1, d1, h1, ct, n, r2=math.inf, d2=None, h2=None, md1=None, md2=None):
        d2 = d1 if d2 is None else d2
        h2 = h1 if h2 is None else h2
        return self.orc.new("CylindricalLens", [r1, d1, h1, d1 if md1 is None else md1, r2, d2, h2, d2 if md2 is None else md2, ct],
                            [self.orc.refindex(n)])


class _PF:
    def __init__(self, m): self.m = m
    def CylindricalLens(self, r1, d1, h1, ct, n, r2=math.inf, d2=None, h2=None, md1=None, md2=None):
        m = self.m
        front = m.CylindricalSurface(r1, d1, h1, md1)
        if math.isinf(r2):
            return m.CylindricalLens(front, ct, n)
        return m.CylindricalLens(front, m.CylindricalSurface(r2, d1 if d2 is None else d2, h1 if h2 is None else h2, md2), ct, n)
    def System(self, objs): return self.m.System(objs)


def _working_distance(orc, lens, offset_z):   # runtests.jl:1699-1710
    sys_ = orc.system([lens])
    b = orc.beam([0.0, -1.0, offset_z], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sys_, b)
    rays = orc.beam_export(sys_, b)[0]["rays"]
    pos, d = rays["pos"][-1], rays["dir"][-1]
    dist = -pos[2] / d[2]
    alpha = math.degrees(math.asin(d[2]))
    return math.cos(math.radians(alpha)) * dist


def test_oracle_thorlabs_cylinder_lenses(orc):
    F = _OF()
    r, d, h, ct = 5.2e-3, 10e-3, 20e-3, 5.9e-3             # LJ1878L2, plano-convex
    lens = F.CylindricalLens(r, d, h, ct, 1.517)
    assert abs(lens.eval("thickness_object", nout=1)[0] - ct) <= 1e-15        # :1727
    edge = ct - abs(r - math.sqrt(r * r - 0.25 * d * d))
    assert abs(edge - 2.12e-3) <= 1e-4                                         # :1729 (thickness of the box section)
    assert abs(_working_distance(orc, lens, 0.05 * d / 2) - 6.1e-3) <= 1e-4    # :1731
    r, d, h, ct = -13.1e-3, 16e-3, 18e-3, 2.0e-3           # LK1900L1, plano-concave
    lens = F.CylindricalLens(r, d, h, ct, 1.517)
    assert abs(lens.eval("thickness_object", nout=1)[0] - ct) <= 1e-15        # :1743
    wd = _working_distance(orc, lens, 0.05 * d / 2)
    f = abs(r) / (1.517 - 1)                                # thin-lens estimate of the (virtual) focus
    assert wd < 0 and abs(abs(wd) - f) < 0.1 * f


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plano_convex", "plano_concave", "biconvex_md", "concave_back"])
def test_gpu_cylindrical_lens_matches_oracle(bmo, orc, case):
    kw = dict(plano_convex=dict(r1=5.2e-3, d1=10e-3, h1=20e-3, ct=5.9e-3, n=1.517),
              plano_concave=dict(r1=-13.1e-3, d1=16e-3, h1=18e-3, ct=2.0e-3, n=1.517),
              biconvex_md=dict(r1=20e-3, d1=12e-3, h1=15e-3, ct=6e-3, n=1.6, r2=-25e-3, md1=16e-3, md2=16e-3),
              concave_back=dict(r1=30e-3, d1=12e-3, h1=15e-3, ct=5e-3, n=1.5, r2=18e-3))[case]
    lens, olens = _PF(bmo).CylindricalLens(**kw), _OF().CylindricalLens(**kw)
    for x in (lens, olens):
        x.zrotate3d_(0.2); x.xrotate3d_(-0.1); x.translate3d_([1e-3, 0.02, -2e-3])
    sys_, osys = bmo.System([lens]), orc.system([olens])
    rng = np.random.default_rng(11)
    n = 256
    pos = np.zeros((n, 3)); pos[:, 0] = rng.uniform(-7e-3, 7e-3, n); pos[:, 2] = rng.uniform(-5e-3, 5e-3, n); pos[:, 1] = -0.05
    d = np.tile([0.0, 1.0, 0.0], (n, 1)) + 0.02 * rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    res = bmo.solve_system_(sys_, bmo.RayBundle(pos, d, 1e-6))
    b, seg = res.beams(), res.segments()
    ref = orc.bulk_trace_rays(osys, pos, d, 1e-6, max_seg=128)
    assert np.array_equal(b["nseg"], ref["nseg"])
    assert (b["nseg"] >= 3).sum() > n // 4                       # a good part of the bundle goes through the lens
    worst = 0.0
    for i in range(n):
        f0, k = int(b["first"][i]), int(b["nseg"][i])
        got = np.concatenate([seg["pos"][f0:f0 + k], seg["dir"][f0:f0 + k]], axis=1)
        worst = max(worst, float(np.abs(got - ref["seg"][i, :k, 0:6]).max()))
        assert np.array_equal(seg["n"][f0:f0 + k], ref["seg"][i, :k, 6])
    assert worst <= POS_TOL, worst


# ---- acylindrical lenses (src/SDFs/AcylindricalSDF.jl; test/runtests.jl:1745-1791) -------------------------------
AYL2520 = dict(radius=15.538e-3, diameter=25e-3, height=50e-3, k=-1.0, ct=7.5e-3, n=1.777,
               A=[0, 1.1926075e-5 * (1e3) ** 3, -2.9323497e-9 * (1e3) ** 5, -1.8718889e-11 * (1e3) ** 7, -1.7009961e-14 * (1e3) ** 9,
                  3.5481542e-17 * (1e3) ** 11, 6.5241296e-20 * (1e3) ** 13])


def _oacyl(orc, r, d, h, k, A, ct, n, r2=math.inf):
    return orc.new("AcylindricalLens", [r, d, h, d, k, float(len(A)), r2, d, h, d, 0.0, -1.0, ct] + list(A), [orc.refindex(n)])


def test_oracle_thorlabs_acylinder_lens(orc):
    L = AYL2520
    lens = _oacyl(orc, L["radius"], L["diameter"], L["height"], L["k"], L["A"], L["ct"], L["n"])
    assert abs(lens.eval("thickness_object", nout=1)[0] - L["ct"]) <= 1.5e-8 * L["ct"]       # runtests.jl:1765
    assert abs(_working_distance(orc, lens, 0.05 * L["diameter"] / 2) - 15.8e-3) <= 1e-4       # :1768
    inv = _oacyl(orc, -L["radius"], L["diameter"], L["height"], L["k"], L["A"], L["ct"], L["n"])
    assert abs(inv.eval("thickness_object", nout=1)[0] - L["ct"]) <= 1.5e-8 * L["ct"]         # :1790


@pytest.mark.gpu
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_gpu_acylindrical_lens_matches_oracle(bmo, orc, sign):
    L = AYL2520
    r = sign * L["radius"]
    lens = bmo.CylindricalLens(bmo.AcylindricalSurface(r, L["diameter"], L["height"], L["k"], L["A"]), L["ct"], L["n"])
    olens = _oacyl(orc, r, L["diameter"], L["height"], L["k"], L["A"], L["ct"], L["n"])
    assert abs(lens.thickness() - L["ct"]) <= 1.5e-8 * L["ct"]
    for x in (lens, olens):
        x.zrotate3d_(-0.15); x.xrotate3d_(0.07); x.translate3d_([2e-3, 0.01, 1e-3])
    rng = np.random.default_rng(21)
    n = 256
    pos = np.zeros((n, 3)); pos[:, 0] = rng.uniform(-0.02, 0.02, n); pos[:, 2] = rng.uniform(-0.014, 0.014, n); pos[:, 1] = -0.05
    d = np.tile([0.0, 1.0, 0.0], (n, 1)) + 0.02 * rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    res = bmo.solve_system_(bmo.System([lens]), bmo.RayBundle(pos, d, 1e-6))
    b, seg = res.beams(), res.segments()
    ref = orc.bulk_trace_rays(orc.system([olens]), pos, d, 1e-6, max_seg=128)
    assert np.array_equal(b["nseg"], ref["nseg"])
    assert (b["nseg"] >= 3).sum() > n // 4
    worst = 0.0
    for i in range(n):
        f0, k = int(b["first"][i]), int(b["nseg"][i])
        got = np.concatenate([seg["pos"][f0:f0 + k], seg["dir"][f0:f0 + k]], axis=1)
        worst = max(worst, float(np.abs(got - ref["seg"][i, :k, 0:6]).max()))
    assert worst <= POS_TOL, worst
