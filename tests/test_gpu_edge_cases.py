"""Edge cases of the trace path on the GPU, against the oracle where the reference defines a result and against
the documented error codes where it throws: empty / degenerate bundles, rays that miss everything, r_max limits
(`while length(rays) < r_max`, System.jl:133), rays that start inside or exactly on a lens, grazing incidence and
total internal reflection, coincident objects (trace_all keeps the first object on equal t, System.jl:62-67),
ragged bundles (every ray a different number of segments), non-finite inputs."""
import ctypes as C
import math

import numpy as np
import pytest

from tests import scenes
from tests.scenes import INCH


def _bundle_vs_oracle(bmo, orc, sys_, osys, pos, d, lam=1e-6, r_max=100, tol=1e-9):
    res = bmo.solve_system_(sys_, bmo.RayBundle(pos, d, lam), r_max=r_max)
    b, seg = res.beams(), res.segments()
    ref = orc.bulk_trace_rays(osys, pos, np.broadcast_to(d, pos.shape), lam, r_max=r_max, max_seg=128)
    assert np.array_equal(b["nseg"], ref["nseg"])
    for i in range(pos.shape[0]):
        f0, k = int(b["first"][i]), int(b["nseg"][i])
        got = np.concatenate([seg["pos"][f0:f0 + k], seg["dir"][f0:f0 + k]], axis=1)
        assert np.abs(got - ref["seg"][i, :k, 0:6]).max() <= tol
        fin = np.isfinite(ref["seg"][i, :k, 7])
        assert np.array_equal(np.isfinite(seg["t"][f0:f0 + k]), fin)
        assert np.array_equal(seg["obj"][f0:f0 + k][fin], ref["seg"][i, :k, 11][fin].astype(int))     # same object hit
    return res, b, seg


def _lens_pair(F):
    l1 = F.SphericalLens(0.1, -0.1, 6e-3, INCH, 1.5)
    l2 = F.SphericalLens(-0.08, math.inf, 4e-3, INCH, 1.7)
    l2.translate3d_([0.0, 0.03, 0.0])
    return F.System([l1, l2]), l1, l2


@pytest.mark.gpu
def test_empty_bundle_and_bad_arguments_are_errors(bmo):
    from bmo_b200 import _lib as L
    sys_, _, _ = _lens_pair(scenes._ProductFactory(bmo))
    dsys = bmo.upload_system(sys_, [1e-6])
    h = C.c_void_p()
    z = np.zeros((1, 3))
    lam = np.zeros(1, np.int32)
    assert L.lib().bmo_trace_rays(dsys.h, 0, L.ptr(z), L.ptr(z), L.ptr(lam), None, None, 100, 0, C.byref(h)) == -1      # BMO_EINVAL: n must be > 0
    assert b"n must be" in L.lib().bmo_last_error()
    assert L.lib().bmo_trace_rays(dsys.h, 1, None, L.ptr(z), L.ptr(lam), None, None, 100, 0, C.byref(h)) == -1
    assert L.lib().bmo_trace_rays(dsys.h, 1, L.ptr(z), L.ptr(z), L.ptr(lam), None, None, 0, 0, C.byref(h)) == -1        # r_max >= 1
    assert L.lib().bmo_retrace(dsys.h, None, 100, 0, C.byref(h)) == -1


@pytest.mark.gpu
def test_ids_out_of_range_are_an_error_not_a_fault(bmo):
    """lambda_id / pose_id index device tables (n_table, lambdas, prims per pose): a bad one must come back as BMO_EINVAL
    (the header's promise), never as an illegal address -- and the context must stay usable afterwards."""
    from bmo_b200 import _lib as L
    sys_, _, _ = _lens_pair(scenes._ProductFactory(bmo))
    dsys = bmo.upload_system(sys_, [1e-6])
    n = 64
    pos = np.zeros((n, 3)); pos[:, 1] = -0.1; pos[:, 0] = np.linspace(-1e-3, 1e-3, n)
    d = np.tile(np.array([0.0, 1.0, 0.0]), (n, 1))
    for lam_ids, pose_ids in ((np.full(n, 1, np.int32), None), (np.full(n, -1, np.int32), None),
                              (np.zeros(n, np.int32), np.full(n, 3, np.int32)), (np.zeros(n, np.int32), np.full(n, -2, np.int32))):
        h = C.c_void_p()
        rc = L.lib().bmo_trace_rays(dsys.h, n, L.ptr(pos), L.ptr(d), L.ptr(lam_ids), None, L.ptr(pose_ids), 100, 0, C.byref(h))
        assert rc == -1 and b"outside" in L.lib().bmo_last_error()
    res = bmo.trace_rays(dsys, pos, d, np.zeros(n, np.int32), r_max=100)          # the same context still traces
    assert res.interactions > 0


@pytest.mark.gpu
def test_rays_that_miss_everything(bmo, orc):
    sys_, _, _ = _lens_pair(scenes._ProductFactory(bmo))
    osys, _, _ = _lens_pair(scenes._OracleFactory())
    n = 200
    pos = np.zeros((n, 3)); pos[:, 0] = np.linspace(0.05, 0.5, n); pos[:, 1] = -0.1
    res, b, seg = _bundle_vs_oracle(bmo, orc, sys_, osys, pos, np.array([0.0, 1.0, 0.0]))
    assert (b["nseg"] == 1).all() and (b["status"] == 1).all() and res.interactions == 0      # BMO_ST_MISS
    assert np.isinf(seg["t"]).all() and (seg["obj"] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("r_max", [1, 2, 3, 5])
def test_r_max_limits_the_number_of_rays_per_beam(bmo, orc, r_max):
    sys_, _, _ = _lens_pair(scenes._ProductFactory(bmo))
    osys, _, _ = _lens_pair(scenes._OracleFactory())
    n = 64
    pos = np.zeros((n, 3)); pos[:, 0] = np.linspace(-8e-3, 8e-3, n); pos[:, 1] = -0.1
    res, b, seg = _bundle_vs_oracle(bmo, orc, sys_, osys, pos, np.array([0.0, 1.0, 0.0]), r_max=r_max)
    assert (b["nseg"] == min(r_max, 5)).all()          # 4 surfaces -> 5 rays when unlimited
    if r_max < 5:
        assert (b["status"] == 3).all()                 # BMO_ST_RMAX: the last ray was not traced


@pytest.mark.gpu
def test_rays_starting_inside_on_and_grazing_a_lens(bmo, orc):
    sys_, l1, _ = _lens_pair(scenes._ProductFactory(bmo))
    osys, _, _ = _lens_pair(scenes._OracleFactory())
    pos, d = [], []
    for x in np.linspace(-5e-3, 5e-3, 11):
        pos.append([x, 3e-3, 0.0]); d.append([0.0, 1.0, 0.0])             # starts inside the first lens
        pos.append([x, 3e-3, 0.0]); d.append([0.3, -1.0, 0.1])            # inside, going backwards and sideways
        pos.append([x, 0.0 + (0.1 - math.sqrt(0.01 - x * x)), 0.0]); d.append([0.0, 1.0, 0.0])   # (almost) exactly on the front surface
    for a in np.linspace(1.2, 1.55, 12):                                    # towards grazing incidence on the rim
        pos.append([-0.05, -0.01, 0.0]); d.append([math.sin(a), math.cos(a), 0.0])
    pos, d = np.array(pos), np.array(d)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    res, b, seg = _bundle_vs_oracle(bmo, orc, sys_, osys, pos, d)
    assert len(set(b["nseg"].tolist())) >= 3            # a ragged bundle: different path lengths in one call
    assert (b["status"] != 7).all()                      # no BMO_ST_ERROR


@pytest.mark.gpu
def test_total_internal_reflection_in_a_prism(bmo, orc):
    def build(F):
        pr = F.RightAnglePrism(20e-3, 20e-3, 1.5)      # legs along -x / -y, hypotenuse x + y = 0: 45 degrees > the critical angle
        return F.System([pr])
    sys_, osys = build(scenes._ProductFactory(bmo)), build(scenes._OracleFactory())
    n = 48
    pos = np.zeros((n, 3)); pos[:, 0] = np.linspace(-8e-3, -1e-3, n); pos[:, 1] = -0.05; pos[:, 2] = np.linspace(-4e-3, 4e-3, n)
    res, b, seg = _bundle_vs_oracle(bmo, orc, sys_, osys, pos, np.array([0.0, 1.0, 0.0]))
    inside = [int(b["first"][i]) + 1 for i in range(n) if b["nseg"][i] >= 4]
    assert len(inside) > n // 2                          # most rays bounce inside the glass (n stays 1.5 across the reflection)
    assert np.all(seg["n"][np.array(inside) + 1] == 1.5)


@pytest.mark.gpu
def test_coincident_objects_first_one_wins(bmo, orc):
    def build(F):
        m1, m2, m3 = F.SquarePlanoMirror2D(INCH), F.SquarePlanoMirror2D(INCH), F.SquarePlanoMirror2D(INCH)
        m3.zrotate3d_(math.radians(10))                 # m1 and m2 coincide exactly; m3 shares their centre with another tilt
        return F.System([m1, m2, m3])
    sys_, osys = build(scenes._ProductFactory(bmo)), build(scenes._OracleFactory())
    pos = np.array([[0.0, -0.1, 0.0], [1e-3, -0.1, 0.0], [-1e-3, -0.1, 0.0], [0.0, -0.1, 2e-3]])
    res, b, seg = _bundle_vs_oracle(bmo, orc, sys_, osys, pos, np.array([0.0, 1.0, 0.0]))
    first_hits = seg["obj"][b["first"]]
    assert set(first_hits.tolist()) <= {0, 2}            # the exact duplicate (object 1) never wins: strict `<` keeps the first object
    assert first_hits[1] != first_hits[2]                # left and right of the axis the tilted mirror is in front / behind


@pytest.mark.gpu
def test_non_finite_inputs_do_not_hang_or_crash(bmo):
    sys_, _, _ = _lens_pair(scenes._ProductFactory(bmo))
    pos = np.array([[0.0, -0.1, 0.0], [np.nan, -0.1, 0.0], [0.0, -0.1, 0.0], [1e300, -0.1, 0.0], [0.0, -0.1, 0.0]])
    d = np.array([[0.0, 1.0, 0.0], [0.0, 1.0, 0.0], [np.nan, 1.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
    res = bmo.solve_system_(sys_, bmo.RayBundle(pos, d, 1e-6, normalize=False), r_max=20)
    b = res.beams()
    assert int(b["nseg"][0]) == 5                        # the healthy ray is unaffected by its neighbours
    assert (b["nseg"][1:] >= 1).all() and (b["nseg"][1:] <= 20).all()
