"""PSFDetector (reference: src/OpticalComponents/Detectors/PSFDetector.jl).

CPU: the oracle's restatement against the reference's own known-answer test, the Airy-disc radius of an
almost-thin lens (test/runtests.jl:2765-2812).  GPU: bmo_psf_* through the C ABI against the oracle:
hit records, calc_local_lims, and the intensity map (1e-8 relative L2).
"""
import math

import numpy as np
import pytest

from tests import scenes
from tests.scenes import INCH

L_, R1, D_LENS, N_, LAM, D_BEAM, F_ = 1e-3, 100e-3, 25.4e-3, 1.5, 1e-6, 15e-3, 200e-3   # runtests.jl:2767-2775


def _airy(F, shift=0.0, group=False):
    lens = F.SphericalLens(R1, math.inf, L_, D_LENS, N_)
    psfd = F.PSFDetector(10e-3)
    psfd.translate3d_([0.0, F_ + 0.13e-3 + shift, 0.0])
    objs = [lens, psfd]
    g = None
    if group:
        g = F.ObjectGroup(objs)
        g.xrotate3d_(math.radians(20)); g.zrotate3d_(math.radians(-35)); g.translate3d_([0.01, -0.02, 0.03])
        objs = [g]
    return dict(system=F.System(objs), lens=lens, psf=psfd, group=g)


class _OF(scenes._OracleFactory):
    def PSFDetector(self, w): return self.orc.new("PSFDetector", [w])


def _disc(n):
    pos, d = scenes.fibonacci_disc(n, diameter=D_BEAM, y0=-10e-3)
    return pos, d


def test_oracle_airy_disc_first_zero(orc):
    """runtests.jl:2786-2802: first zero of the PSF through the centre column = 1.22 lambda f / D within 1 %."""
    sc = _airy(_OF())
    pos, d = _disc(1000)
    for p in pos:
        b = orc.beam(p, [0.0, 1.0, 0.0], LAM)
        orc.solve_system_(sc["system"], b)
    data = sc["psf"].psf_data()
    assert data.shape == (1000, 9)                                   # runtests.jl:2806
    x, z, I = sc["psf"].psf_intensity(n=500, crop_factor=5, center="bbox")
    ci = np.unravel_index(np.argmax(I), I.shape)
    num_min = x[np.argmin(I[:, ci[1]])]
    airy_min = 1.22 * LAM * F_ / D_BEAM
    assert abs(abs(num_min) - airy_min) <= 1e-2 * airy_min
    sc["psf"].psf_empty()
    assert sc["psf"].psf_data().shape[0] == 0                        # runtests.jl:2809-2810


def test_oracle_psf_opl_through_parents(orc):
    """optical_path_length(beam) includes the parents (Beam.jl:137-149): a ray that reaches the detector
    through a beamsplitter carries the OPL of the stem."""
    F = _OF()
    bs = F.ThinBeamsplitter(INCH, reflectance=0.5)
    bs.zrotate3d_(math.radians(45))
    psfd = F.PSFDetector(10e-3); psfd.translate3d_([0.0, 0.1, 0.0])
    sys_ = F.System([bs, psfd])
    b = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], LAM)
    orc.solve_system_(sys_, b)
    data = psfd.psf_data()
    assert data.shape[0] == 1 and abs(data[0, 6] - 0.2) < 1e-12 and abs(data[0, 7] - 1.0) < 1e-12
    assert abs(data[0, 8] - 2 * math.pi / LAM) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("group", [False, True])
def test_gpu_psf_matches_oracle(bmo, orc, group):
    n_rays, n = 1000, 160
    sc, osc = _airy(scenes._ProductFactory(bmo), group=group), _airy(_OF(), group=group)
    pos, d = _disc(n_rays)
    if group:   # rotate the bundle with the group so it still goes through the lens
        R = np.array(sc["group"].dir) if hasattr(sc["group"], "dir") else None
        src = bmo.UniformDiscSource(tuple(np.array(sc["lens"].shape.pos) - 10e-3 * np.array(sc["lens"].shape.dir)[:, 1]),
                                    tuple(np.array(sc["lens"].shape.dir)[:, 1]), D_BEAM, LAM, num_rays=n_rays,
                                    e1=tuple(np.array(sc["lens"].shape.dir)[:, 0]))
        pos, d = src.pos, src.dir
    bundle = bmo.RayBundle(pos, d, LAM)
    bmo.solve_system_(sc["system"], bundle)
    dd = np.broadcast_to(np.asarray(d, dtype=np.float64), pos.shape)
    for p, q in zip(pos, dd):
        orc.solve_system_(osc["system"], orc.beam(p, q, LAM))
    got, ref = sc["psf"].data, osc["psf"].psf_data()
    assert got.shape == ref.shape == (n_rays, 9)
    assert np.abs(got[:, 0:6] - ref[:, 0:6]).max() <= 1e-9 * max(1.0, np.abs(ref[:, 0:3]).max())
    assert np.abs(got[:, 6] - ref[:, 6]).max() <= 1e-12 * np.abs(ref[:, 6]).max()        # optical path length
    assert np.abs(got[:, 7] - ref[:, 7]).max() <= 1e-12
    assert np.array_equal(got[:, 8], ref[:, 8])
    for center in ("centroid", "bbox"):
        a, b = np.array(sc["psf"].calc_local_lims(2.0, center)), osc["psf"].psf_lims(2.0, center)
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max() + 1e-15      # local coordinates are differences of O(0.1 m) world coordinates
    x, z, I = sc["psf"].intensity(n=n, crop_factor=5, center="bbox")
    xo, zo, Io = osc["psf"].psf_intensity(n=n, crop_factor=5, center="bbox")
    assert np.abs(x - xo).max() <= 1e-15 and np.abs(z - zo).max() <= 1e-15
    assert Io.max() > 0.5 * n_rays ** 2                                                   # a focused spot
    assert np.linalg.norm((I - Io).ravel()) / np.linalg.norm(Io.ravel()) <= 1e-8
    # explicit window + shifts (PSFDetector.jl:202-211)
    kw = dict(n=64, x_min=-2e-5, x_max=3e-5, z_min=-1e-5, z_max=1e-5, x0_shift=2e-6, z0_shift=-1e-6)
    x, z, I = sc["psf"].intensity(**kw)
    xo, zo, Io = osc["psf"].psf_intensity(**kw)
    assert np.abs(x - xo).max() <= 1e-15
    assert np.linalg.norm((I - Io).ravel()) / np.linalg.norm(Io.ravel()) <= 1e-8
    # the detector accumulates until empty! (PSFDetector.jl:37-40)
    bmo.solve_system_(sc["system"], bmo.RayBundle(pos[:10], d, LAM))
    assert len(sc["psf"]) == n_rays + 10
    sc["psf"].empty_()
    assert len(sc["psf"]) == 0


@pytest.mark.gpu
def test_gpu_psf_opl_through_beamsplitter_children(bmo, orc):
    def build(F):
        bs = F.ThinBeamsplitter(INCH, reflectance=0.5); bs.zrotate3d_(math.radians(45))
        m = F.SquarePlanoMirror2D(INCH); m.zrotate3d_(math.radians(90)); m.translate3d_([-0.05, 0.0, 0.0])
        lens = F.SphericalLens(0.1, math.inf, 2e-3, INCH, 1.5); lens.translate3d_([0.0, 0.03, 0.0])
        psfd = F.PSFDetector(10e-3); psfd.translate3d_([0.0, 0.2, 0.0])
        return dict(system=F.System([bs, m, lens, psfd]), psf=psfd)
    class PF(scenes._ProductFactory):
        pass
    sc, osc = build(scenes._ProductFactory(bmo)), build(_OF())
    rng = np.random.default_rng(1)
    pos = np.zeros((50, 3)); pos[:, 0] = rng.uniform(-2e-3, 2e-3, 50); pos[:, 2] = rng.uniform(-2e-3, 2e-3, 50); pos[:, 1] = -0.1
    bmo.solve_system_(sc["system"], bmo.RayBundle(pos, np.array([0.0, 1.0, 0.0]), LAM))
    for p in pos:
        orc.solve_system_(osc["system"], orc.beam(p, [0.0, 1.0, 0.0], LAM))
    got, ref = sc["psf"].data, osc["psf"].psf_data()
    assert got.shape == ref.shape and got.shape[0] >= 50
    # push! order differs (wavefront vs depth of the beam tree): compare as sets keyed by the hit point
    ka, kb = np.lexsort(np.round(got[:, [0, 2, 6]], 9).T), np.lexsort(np.round(ref[:, [0, 2, 6]], 9).T)
    assert np.abs(got[ka] - ref[kb]).max() <= 1e-9 * np.abs(ref).max()
    x, z, I = sc["psf"].intensity(n=48, crop_factor=2)
    xo, zo, Io = osc["psf"].psf_intensity(n=48, crop_factor=2)
    assert np.linalg.norm((I - Io).ravel()) / np.linalg.norm(Io.ravel()) <= 1e-8
