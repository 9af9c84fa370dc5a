"""PolarizationFilter (reference: src/OpticalComponents/Polarizers/PolarizationFilter.jl, JonesCalculus.jl).

CPU: the oracle against the reference's known-answer tests (test/runtests.jl:2469-2535): Malus' law and the
projected polarisation direction for a filter rotated about the ray, and I = 0.75 cos^2 + 0.25 for the filter
tilted by 45 degrees.  GPU: the same loops (they re-solve the same beam, i.e. the retrace path) through the
C ABI against the oracle.
"""
import math

import numpy as np
import pytest

from tests import scenes

MM = 1e-3
THETAS = np.arange(1, 360, 10)       # thetas = 1:10:360


class _OF(scenes._OracleFactory):
    def PolarizationFilter(self, edge): return self.orc.new("PolarizationFilter", [edge])


def _rot(axis, th):
    """rotate3d (LinearAlgebraUtils.jl:55-65) for the reference vector of the test."""
    a = np.asarray(axis, dtype=np.float64); a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def _angle(a, b):
    return math.degrees(math.acos(np.clip(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)), -1, 1)))


def _linpol_angles():
    out = np.zeros(len(THETAS))      # runtests.jl:2444-2456
    for i, t in enumerate(THETAS - 1):
        out[i] = 180 if 90 < t <= 270 else 0
    return out


def _oracle_pose(f):
    pos, R = f.pose()
    return pos, R


def test_oracle_malus_law_normal_incidence(orc):
    f = _OF().PolarizationFilter(5 * MM)
    sys_ = orc.system([f])
    pos, R = _oracle_pose(f)
    ray_dir, pol = R[:, 1].copy(), R[:, 0].copy()
    ray_pos = pos - 10 * MM * ray_dir
    beam = orc.polarized_beam(ray_pos, ray_dir, 1e-6, pol.astype(np.complex128))
    Rm = _rot(ray_dir, math.radians(10))
    i_n, ang = [], []
    for th in THETAS:
        orc.solve_system_(sys_, beam, retrace=True)
        rays = orc.beam_export(sys_, beam)[0]["rays"]
        assert len(rays["t"]) == 2
        E1 = rays["E0"][1]
        assert np.abs(E1.imag).max() == 0.0                       # linear polarisation stays linear
        i_n.append(float(np.linalg.norm(E1) ** 2))
        ang.append(_angle(E1.real, pol))
        f.rotate3d_(ray_dir, math.radians(10))
        pol = Rm @ pol
    assert np.allclose(ang, _linpol_angles(), rtol=1.5e-8, atol=1e-6)      # runtests.jl:2505 (angles_n ≈ angles_a)
    assert np.allclose(i_n, np.cos(np.radians(THETAS - 1)) ** 2, rtol=1e-9, atol=1e-12)   # Malus' law


def test_oracle_tilted_filter(orc):
    f = _OF().PolarizationFilter(5 * MM)
    sys_ = orc.system([f])
    f.translate3d_([0, 10 * MM, 0]); f.xrotate3d_(math.radians(45)); f.zrotate3d_(math.radians(30))   # runtests.jl:2509-2511
    pos, R = _oracle_pose(f)
    ray_dir, local_x = R[:, 1].copy(), R[:, 0].copy()
    beam = orc.polarized_beam(pos - 10 * MM * ray_dir, ray_dir, 1e-6, local_x.astype(np.complex128))
    f.rotate3d_(local_x, math.radians(45))
    i_n = []
    for th in THETAS:
        orc.solve_system_(sys_, beam, retrace=True)
        rays = orc.beam_export(sys_, beam)[0]["rays"]
        i_n.append(float(np.linalg.norm(rays["E0"][1]) ** 2))
        f.rotate3d_(ray_dir, math.radians(10))
    assert np.allclose(i_n, np.cos(np.radians(THETAS - 1)) ** 2 * 0.75 + 0.25, rtol=1.5e-8)   # runtests.jl:2532-2533


@pytest.mark.gpu
@pytest.mark.parametrize("tilted", [False, True])
def test_gpu_polfilter_matches_oracle(bmo, orc, tilted):
    f, of = bmo.PolarizationFilter(5 * MM), _OF().PolarizationFilter(5 * MM)
    sys_, osys = bmo.System([f]), orc.system([of])
    if tilted:
        for x in (f, of):
            x.translate3d_([0, 10 * MM, 0]); x.xrotate3d_(math.radians(45)); x.zrotate3d_(math.radians(30))
    pos, R = _oracle_pose(of)
    assert np.allclose(np.array(f.shape.dir), R, atol=1e-15)
    ray_dir, local_x = R[:, 1].copy(), R[:, 0].copy()
    ray_pos = pos - 10 * MM * ray_dir
    if tilted:
        f.rotate3d_(local_x, math.radians(45)); of.rotate3d_(local_x, math.radians(45))
    beam = bmo.Beam(bmo.PolarizedRay(ray_pos, ray_dir, 1e-6, E0=local_x.astype(np.complex128)))
    obeam = orc.polarized_beam(ray_pos, ray_dir, 1e-6, local_x.astype(np.complex128))
    for th in THETAS[:12]:
        bmo.solve_system_(sys_, beam)                      # second and later solves retrace the stored path
        orc.solve_system_(osys, obeam, retrace=True)
        ref = orc.beam_export(osys, obeam)[0]["rays"]
        assert len(beam.rays) == len(ref["t"]) == 2
        got = np.array(beam.rays[1].E0)
        assert np.abs(got - ref["E0"][1]).max() <= 1e-12
        assert np.abs(np.array(beam.rays[1].pos) - ref["pos"][1]).max() <= 1e-12
        f.rotate3d_(ray_dir, math.radians(10)); of.rotate3d_(ray_dir, math.radians(10))
    # a bundle of oblique rays through the same filter, fresh trace
    rng = np.random.default_rng(5)
    n = 64
    d = np.tile(ray_dir, (n, 1)) + 0.05 * rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = np.tile(ray_pos, (n, 1)) + 0.2 * MM * rng.standard_normal((n, 3))
    E = np.cross(d, rng.standard_normal((n, 3))); E /= np.linalg.norm(E, axis=1, keepdims=True)
    # E0 is held orthogonal to dir at 1e-14 by the PolarizedRay constructor: re-orthogonalise in double
    E -= np.sum(E * d, axis=1, keepdims=True) * d
    res = bmo.solve_system_(sys_, bmo.RayBundle(p, d, 1e-6, E0=E.astype(np.complex128)))
    seg, b = res.segments(), res.beams()
    hits = 0
    for i in range(n):
        ob = orc.polarized_beam(p[i], d[i], 1e-6, E[i].astype(np.complex128))
        try:
            orc.solve_system_(osys, ob)
        except orc.OracleError:
            continue                                        # the reference throws (E0 not orthogonal at 1e-14): flagged on the GPU
        ref = orc.beam_export(osys, ob)[0]["rays"]
        f0, ns = int(b["first"][i]), int(b["nseg"][i])
        assert ns == len(ref["t"])
        if ns == 2:
            hits += 1
            assert np.abs(seg["E0"][f0 + 1] - ref["E0"][1]).max() <= 1e-12
    assert hits > n // 2
