"""retrace_system! (reference: src/System.jl:188-255 for Beam, :326-428 for GaussianBeamlet, driven by
solve_system!(...; retrace=true), :444-461).

CPU part: the oracle's restatement against the properties the reference's own test relies on
(test/runtests.jl:1082-1092: retracing a solved multipass cell reproduces the solution) and against the
semantics the source spells out (tail dropped where the path breaks, heads of stored children modified,
retraced rays blind to objects that moved into their path).
GPU part: bmo_retrace through the C ABI against the oracle on the same scenes.
"""
import math

import numpy as np
import pytest

from tests import scenes, scenes2 as s2
from tests.scenes import INCH

POS_TOL = 1e-9
FIELD_TOL = 1e-8


# ---- scenes -------------------------------------------------------------------------------------
def _ring(F, n_mirrors=21, radius=1.0):
    """Circular multipass cell of flat mirrors (test/runtests.jl:1009-1030), fewer mirrors."""
    L = 6 * radius / n_mirrors
    dth = 360 / (n_mirrors + 1)
    mirrors = [F.SquarePlanoMirror2D(L) for _ in range(n_mirrors)]
    th = dth
    for m in mirrors:
        m.zrotate3d_(math.radians(th))
        m.translate3d_([radius * math.cos(math.radians(th)), radius * math.sin(math.radians(th)), 0.0])
        th += dth
    for m in mirrors:
        m.zrotate3d_(math.radians(90))
    c, s = math.cos(math.radians(dth)), math.sin(math.radians(dth))
    d = [-c, -s, 0.0]
    origin = [radius + d[0] * -1, d[1] * -1, 0.0]
    return dict(system=F.System(mirrors), mirrors=mirrors, origin=origin, dir=d)


def _fold(F):
    """lens -> fold mirror -> thin beamsplitter -> two end mirrors; used for path-break / occluder cases."""
    lens = F.SphericalLens(0.2, 0.2, 5e-3, INCH, 1.5)
    fold = F.SquarePlanoMirror2D(INCH)
    fold.zrotate3d_(math.radians(-45)); fold.translate3d_([0.0, 0.1, 0.0])
    bs = F.ThinBeamsplitter(INCH, reflectance=0.5)
    bs.zrotate3d_(math.radians(45)); bs.translate3d_([-0.1, 0.1, 0.0])
    e1 = F.SquarePlanoMirror2D(INCH)
    e1.zrotate3d_(math.radians(90)); e1.translate3d_([-0.2, 0.1, 0.0])
    e2 = F.SquarePlanoMirror2D(INCH)
    e2.translate3d_([-0.1, 0.0, 0.0])
    blocker = F.SquarePlanoMirror2D(INCH)
    blocker.translate3d_([0.0, 0.05, 1.0])      # parked out of the beam
    return dict(system=F.System([lens, fold, bs, e1, e2, blocker]), lens=lens, fold=fold, bs=bs, e1=e1, e2=e2, blocker=blocker)


RETRACE_ULP = 1e-13     # normalize() of an already (almost) unit direction moves it by an ulp or two; hit points follow


def _tree_equal(a, b, tol=0.0):
    assert [(x["parent"], len(x["rays"]["t"])) for x in a] == [(x["parent"], len(x["rays"]["t"])) for x in b]
    for x, y in zip(a, b):
        for k in ("pos", "dir", "n", "t"):
            u, v = x["rays"][k], y["rays"][k]
            fin = np.isfinite(v)
            assert np.array_equal(np.isfinite(u), fin)
            if tol == 0.0:
                assert np.array_equal(u[fin], v[fin]), k
            elif fin.any():
                assert np.abs(u[fin] - v[fin]).max() <= tol * max(1.0, np.abs(v[fin]).max()), k
        assert np.array_equal(x["rays"]["obj"], y["rays"]["obj"])


# ---- oracle (CPU) -------------------------------------------------------------------------------
def test_oracle_retrace_reproduces_solved_ring(orc):
    sc = _ring(scenes._OracleFactory())
    b = orc.beam(sc["origin"], sc["dir"], 1e-6)
    orc.solve_system_(sc["system"], b, r_max=1000)
    first = orc.beam_export(sc["system"], b)
    assert len(first[0]["rays"]["t"]) == 21 + 1            # runtests.jl:1055: n_mirrors + 1 rays
    orc.solve_system_(sc["system"], b, r_max=1000, retrace=True)
    # replace! writes the re-validated successor rays through direction!, which normalises (Beam.jl:81-95, AbstractRay.jl:83-86):
    # a retraced path equals the solved one up to that rounding, not to the bit
    _tree_equal(orc.beam_export(sc["system"], b), first, tol=RETRACE_ULP)


def test_oracle_retrace_follows_moved_mirror_and_matches_fresh_trace(orc):
    F = scenes._OracleFactory()
    sc = _fold(F)
    b = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], b)
    t0 = orc.beam_export(sc["system"], b)
    assert len(t0) >= 3                                      # splitter: children exist
    sc["e1"].translate3d_([-1e-3, 0.0, 0.0])
    sc["lens"].translate3d_([2e-4, 0.0, 0.0])
    orc.solve_system_(sc["system"], b, retrace=True)
    fresh = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], fresh)
    _tree_equal(orc.beam_export(sc["system"], b), orc.beam_export(sc["system"], fresh), tol=RETRACE_ULP)


def test_oracle_retrace_drops_tail_when_path_breaks(orc):
    F = scenes._OracleFactory()
    sc = _fold(F)
    b = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], b)
    sc["fold"].translate3d_([0.0, 0.0, 0.5])                 # the fold mirror leaves the beam
    orc.solve_system_(sc["system"], b, retrace=True)
    fresh = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], fresh)
    t = orc.beam_export(sc["system"], b)
    assert len(t) == 1                                       # children dropped (System.jl:249)
    _tree_equal(t, orc.beam_export(sc["system"], fresh), tol=RETRACE_ULP)


def test_oracle_retrace_is_blind_to_new_occluder(orc):
    """System.jl:209-211: a stored ray is intersected with the object it hit before, nothing else."""
    F = scenes._OracleFactory()
    sc = _fold(F)
    b = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], b)
    t0 = orc.beam_export(sc["system"], b)
    sc["blocker"].translate3d_([0.0, 0.0, -1.0])             # now between lens and fold mirror
    orc.solve_system_(sc["system"], b, retrace=True)
    _tree_equal(orc.beam_export(sc["system"], b), t0, tol=RETRACE_ULP)        # unchanged: the blocker is not seen
    fresh = orc.beam([0.0, -0.1, 0.0], [0.0, 1.0, 0.0], 1e-6)
    orc.solve_system_(sc["system"], fresh)
    assert len(orc.beam_export(sc["system"], fresh)) == 1    # a fresh trace is stopped... reflected by it


def test_oracle_retrace_gaussian_michelson_scan(orc):
    """Michelson arm scan with retrace=true (the loop of test/runtests.jl:2105-2113): the retraced
    detector field equals a fresh solve except for the stale child waists (Gaussian.jl:154-162)."""
    n = 32
    osc = scenes.michelson_oracle(pd_n=n)
    B = scenes.MICHELSON_BEAM
    g = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    orc.solve_system_(osc["system"], g)
    osc["pd"].pd_empty()
    osc["m1"].translate3d_([40e-9, 0.0, 0.0])
    orc.solve_system_(osc["system"], g, retrace=True)
    f_re = osc["pd"].pd_field(n)
    osc2 = scenes.michelson_oracle(pd_n=n, m1_shift=40e-9)
    g2 = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    orc.solve_system_(osc2["system"], g2)
    f_fr = osc2["pd"].pd_field(n)
    assert np.abs(f_fr).max() > 0
    assert np.linalg.norm(f_re - f_fr) / np.linalg.norm(f_fr) <= 1e-6
    a, b = orc.gauss_export(osc["system"], g), orc.gauss_export(osc2["system"], g2)
    assert [(x["parent"], len(x["chief"]["t"])) for x in a] == [(x["parent"], len(x["chief"]["t"])) for x in b]


# ---- GPU ------------------------------------------------------------------------------------------
def _compare_beam_tree(res, tree, tol=POS_TOL):
    order = res.bfs_order()
    b, seg = res.beams(), res.segments()
    assert len(order) == len(tree)
    for bi, t in zip(order, tree):
        f, ns = int(b["first"][bi]), int(b["nseg"][bi])
        r = t["rays"]
        assert ns == len(r["t"]), (ns, len(r["t"]))
        sl = slice(f, f + ns)
        assert np.array_equal(seg["obj"][sl][np.isfinite(r["t"])], r["obj"][np.isfinite(r["t"])])
        assert np.array_equal(np.isfinite(seg["t"][sl]), np.isfinite(r["t"]))
        assert np.array_equal(seg["n"][sl], r["n"])
        assert np.abs(seg["pos"][sl] - r["pos"]).max() <= tol * max(1.0, np.abs(r["pos"]).max())
        assert np.abs(seg["dir"][sl] - r["dir"]).max() <= tol


@pytest.mark.gpu
def test_gpu_retrace_ring(bmo, orc):
    sc, osc = _ring(scenes._ProductFactory(bmo)), _ring(scenes._OracleFactory())
    beam = bmo.Beam(bmo.Ray(sc["origin"], sc["dir"], 1e-6))
    ob = orc.beam(osc["origin"], osc["dir"], 1e-6)
    cb = bmo.counters()["tri_tests"]
    r1 = bmo.solve_system_(sc["system"], beam, r_max=1000)
    orc.solve_system_(osc["system"], ob, r_max=1000)
    c0 = bmo.counters()["tri_tests"]
    r2 = bmo.solve_system_(sc["system"], beam, r_max=1000)          # retrace=True is the default
    c1 = bmo.counters()["tri_tests"]
    orc.solve_system_(osc["system"], ob, r_max=1000, retrace=True)
    _compare_beam_tree(r2, orc.beam_export(osc["system"], ob))
    assert len(beam.rays) == 22
    # the sequential path tests only the mirror a stored ray hit before (2 triangles); the last ray, which
    # had no intersection, is traced against the whole system again (System.jl:200-206 + solve_leaf!)
    assert c1 - c0 < c0 - cb
    assert r2.interactions == r1.interactions


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["move", "break", "occluder"])
def test_gpu_retrace_fold_scene(bmo, orc, case):
    sc, osc = _fold(scenes._ProductFactory(bmo)), _fold(scenes._OracleFactory())
    n = 64
    rng = np.random.default_rng(3)
    pos = np.zeros((n, 3)); pos[:, 0] = rng.uniform(-3e-3, 3e-3, n); pos[:, 2] = rng.uniform(-3e-3, 3e-3, n); pos[:, 1] = -0.1
    bundle = bmo.RayBundle(pos, np.array([0.0, 1.0, 0.0]), 1e-6)
    obs = [orc.beam(p, [0.0, 1.0, 0.0], 1e-6) for p in pos]
    bmo.solve_system_(sc["system"], bundle)
    for ob in obs:
        orc.solve_system_(osc["system"], ob)
    for s in (sc, osc):
        if case == "move":
            s["e1"].translate3d_([-1e-3, 0.0, 0.0]); s["lens"].translate3d_([2e-4, 0.0, 0.0]); s["e2"].xrotate3d_(1e-3)
        elif case == "break":
            s["fold"].translate3d_([0.0, 0.0, INCH])
        else:
            s["blocker"].translate3d_([0.0, 0.0, -1.0])
    res = bmo.solve_system_(sc["system"], bundle)
    tree = []
    for ob in obs:
        orc.solve_system_(osc["system"], ob, retrace=True)
        tree += orc.beam_export(osc["system"], ob)
    _compare_beam_tree(res, tree)
    if case == "occluder":
        fresh = bmo.solve_system_(sc["system"], bmo.RayBundle(pos, np.array([0.0, 1.0, 0.0]), 1e-6))
        assert fresh.n_beams == n and res.n_beams > n          # the fresh trace sees the blocker, the retrace does not


@pytest.mark.gpu
def test_gpu_retrace_polarized_branching(bmo, orc):
    n = 32
    sc, osc = s2.mesh_scene(bmo, nu=24, nv=24), s2.mesh_scene_oracle(nu=24, nv=24)
    pos, d, E0 = s2.jittered_lattice(n)
    bundle = bmo.RayBundle(pos, d, 1e-6, E0=E0)
    obs = [orc.polarized_beam(p, d, 1e-6, E0) for p in pos]
    bmo.solve_system_(sc["system"], bundle)
    for ob in obs:
        orc.solve_system_(osc["system"], ob)
    for s in (sc, osc):
        s["retro"].translate3d_([0.0, 1e-3, 0.0]); s["ball"].translate3d_([1e-3, 0.0, 0.0])
    res = bmo.solve_system_(sc["system"], bundle)
    b, seg = res.beams(), res.segments()
    order = res.bfs_order()
    k = 0
    for ob in obs:
        orc.solve_system_(osc["system"], ob, retrace=True)
        for t in orc.beam_export(osc["system"], ob):
            bi = order[k]; k += 1
            f, ns = int(b["first"][bi]), int(b["nseg"][bi])
            r = t["rays"]
            assert ns == len(r["t"])
            sl = slice(f, f + ns)
            assert np.abs(seg["pos"][sl] - r["pos"]).max() <= POS_TOL and np.abs(seg["dir"][sl] - r["dir"]).max() <= POS_TOL
            # (a degenerate s/p basis at the splitter gives NaN polarisation in the reference arithmetic for two
            #  of these rays, fresh trace or retrace alike: same rays, same NaNs on both sides)
            nan = np.isnan(r["E0"])
            assert np.array_equal(np.isnan(seg["E0"][sl]), nan)
            if (~nan).any():
                assert np.abs(seg["E0"][sl][~nan] - r["E0"][~nan]).max() <= 1e-9 * np.abs(r["E0"][~nan]).max()
    assert k == res.n_beams


@pytest.mark.gpu
def test_gpu_retrace_michelson_scan(bmo, orc):
    """C1 with the reference's scan loop: solve, move the arm mirror, empty!(pd), solve again (retrace)."""
    n = 64
    sc, osc = scenes.michelson(bmo, pd_n=n), scenes.michelson_oracle(pd_n=n)
    B = scenes.MICHELSON_BEAM
    g = bmo.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], support=B["support"])
    bmo.solve_system_(sc["system"], g)
    orc.solve_system_(osc["system"], og)
    for step in range(3):
        sc["m1"].translate3d_([25e-9, 0.0, 0.0]); osc["m1"].translate3d_([25e-9, 0.0, 0.0])
        sc["pd"].empty_(); osc["pd"].pd_empty()
        res = bmo.solve_system_(sc["system"], g)
        orc.solve_system_(osc["system"], og, retrace=True)
        ref = orc.gauss_export(osc["system"], og)
        order, b = res.bfs_order(), res.beams()
        assert [int(b["nseg"][i]) for i in order] == [len(r["chief"]["t"]) for r in ref]
        for i, r in zip(order, ref):
            assert abs(b["w0"][i] - r["w0"]) <= 1e-12 * r["w0"]            # children keep their stored w0
            assert abs(b["E0"][i] - r["E0"]) <= 1e-10 * abs(r["E0"])
        field, ofield = sc["pd"].field, osc["pd"].pd_field(n)
        assert np.abs(ofield).max() > 0
        assert np.linalg.norm((field - ofield).ravel()) / np.linalg.norm(ofield.ravel()) <= FIELD_TOL
