"""GPU-vs-oracle parity of the device code paths that had no GPU test in round 1 (VERDICT "missing" #1/#2):
plate beamsplitters (Rectangular / Round) for Ray, PolarizedRay and GaussianBeamlet including the
coating-preference tie of `intersect3d(::AbstractPlateBeamsplitter, ray)` (PlateBeamsplitter.jl:160-187) and its
`interact3d` (:189-275); meniscus lenses (MeniscusLensSDF.jl:42-46, 122-189); the primitive SDFs used directly as
Prism / Mirror shapes, rotated -- Box, Cylinder, CutSphere, Ring (PrimitiveSDF.jl:41-46, 71-76, 112-124, 162-166) and
Sphere (SphericalLensSDF.jl:86-89, orientation pinned to I); ConcaveSphericalMirror (Mirrors.jl:186-195);
IntersectableObject (Intersectable.jl:15); RectangularCompensatorPlate (Compensators.jl:15-24); StaticSystem
(System.jl:38-45); the binary-STL loader (Mesh.jl:48-70) on the reference's own asset docs/src/assets/Mirror_Post.stl
(committed as tests/fixtures/Mirror_Post.stl) and BASELINE config 4 with that mesh and the unscaled Fresnel rhomb of
test/runtests.jl:2339-2349.  Tolerances: hit points / directions 1e-9 relative (north_star); plain rays through SDFs
are also compared bit for bit."""
import math
import os

import numpy as np
import pytest

from tests import scenes
from tests import scenes2 as s2
from tests.scenes import INCH

POS_TOL = 1e-9
FIELD_TOL = 1e-8
STL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "Mirror_Post.stl")


# ---- helpers -----------------------------------------------------------------------------------------
def _compare_ray_trees(bmo, orc, sys_, osys, pos, d, lam=1e-6, E0=None, r_max=100, bitwise=False, tol=POS_TOL):
    """Trace every ray of the bundle through both implementations and compare the complete beam trees (BFS order):
    segment counts, hit objects, refractive indices, positions / directions (and E0)."""
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    n = pos.shape[0]
    bundle = bmo.RayBundle(pos, d, lam, E0=E0)
    res = bmo.solve_system_(sys_, bundle, r_max=r_max)
    b, seg = res.beams(), res.segments()
    order = res.bfs_order()
    k = inter = 0
    worst = worst_e = 0.0
    dirs = np.broadcast_to(np.asarray(d, dtype=np.float64), pos.shape)
    for i in range(n):
        ob = orc.polarized_beam(pos[i], dirs[i], lam, E0) if E0 is not None else orc.beam(pos[i], dirs[i], lam)
        orc.solve_system_(osys, ob, r_max=r_max)
        for t in orc.beam_export(osys, ob):
            bi = order[k]; k += 1
            f, ns = int(b["first"][bi]), int(b["nseg"][bi])
            r = t["rays"]
            assert ns == len(r["t"]), (i, bi, ns, len(r["t"]))
            sl = slice(f, f + ns)
            assert np.array_equal(seg["obj"][sl], r["obj"]), (i, bi, seg["obj"][sl], r["obj"])
            assert np.array_equal(seg["n"][sl], r["n"]), (i, bi)
            assert np.array_equal(np.isinf(seg["t"][sl]), np.isinf(r["t"]))
            if bitwise:
                assert np.array_equal(seg["pos"][sl], r["pos"]) and np.array_equal(seg["dir"][sl], r["dir"]), (i, bi)
                fin = np.isfinite(r["t"])
                assert np.array_equal(seg["t"][sl][fin], r["t"][fin]) and np.array_equal(seg["nrm"][sl][fin], r["nrm"][fin]), (i, bi)
            worst = max(worst, np.abs(seg["pos"][sl] - r["pos"]).max() / max(np.abs(r["pos"]).max(), 1e-300), np.abs(seg["dir"][sl] - r["dir"]).max())
            if E0 is not None:
                worst_e = max(worst_e, np.abs(seg["E0"][sl] - r["E0"]).max() / np.abs(r["E0"]).max())
            inter += int(np.isfinite(r["t"]).sum())
    assert k == res.n_beams, (k, res.n_beams)
    assert res.interactions == inter
    assert worst <= tol, worst
    assert worst_e <= tol, worst_e
    return res, b, seg


def _compare_gauss_tree(bmo, orc, sys_, osys, B, pd=None, opd=None, r_max=100):
    g = bmo.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B.get("M2", 1.0), P0=B.get("P0", 1e-3), support=B["support"])
    og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B.get("M2", 1.0), P0=B.get("P0", 1e-3), support=B["support"])
    res = bmo.solve_system_(sys_, g, r_max=r_max)
    orc.solve_system_(osys, og, r_max=r_max)
    ref = orc.gauss_export(osys, og)
    order, b, seg = res.bfs_order(), res.beams(), res.segments()
    assert len(order) == len(ref)
    assert [int(b["nseg"][i]) for i in order] == [len(r["chief"]["t"]) for r in ref]
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
    assert worst <= POS_TOL, worst
    if pd is not None:
        ofield = opd.pd_field(pd.n)
        assert np.abs(ofield).max() > 0
        assert np.linalg.norm((pd.field - ofield).ravel()) / np.linalg.norm(ofield.ravel()) <= FIELD_TOL
    return res, ref


def _fan(n, half=8e-3, y0=-0.1, tilt=0.0):
    """n rays on a line across the aperture (x) with a second coordinate in z, all along +y (optionally tilted in x)."""
    pos = np.zeros((n, 3))
    pos[:, 0] = np.linspace(-half, half, n)
    pos[:, 2] = np.linspace(-half / 3, half / 2, n)[::-1]
    pos[:, 1] = y0
    d = np.array([math.sin(tilt), math.cos(tilt), 0.0])
    return pos, d


# ---- plate beamsplitters ---------------------------------------------------------------------------------
def _plate_scene(F, kind, tilt_deg=45.0, n=1.5, with_pd=None):
    """A plate splitter tilted about z; the transmitted child meets a fold mirror (or the detector), the reflected child
    another one.  Both mirrors steer their beams past the plate: a closed path through a splitter has no end
    (solve_system! limits the rays per beam, not the depth of the beam tree)."""
    if kind == "rect":
        bs = F.RectangularPlateBeamsplitter(36e-3, 25e-3, 3e-3, n, reflectance=0.4)
    else:
        bs = F.RoundPlateBeamsplitter(INCH, 5e-3, n, reflectance=0.6)
    bs.zrotate3d_(math.radians(tilt_deg))
    side = -1.0 if tilt_deg >= 0 else 1.0
    m_r = F.SquarePlanoMirror2D(2 * INCH)          # catches the reflected child
    m_r.zrotate3d_(math.radians(90 + 20)); m_r.translate3d_([-0.07 * side, 0.0, 0.0])
    objs = [bs, m_r]
    pd = None
    if with_pd:
        pd = F.Photodetector(20e-3, with_pd)
        pd.zrotate3d_(math.radians(3)); pd.translate3d_([0.0, 0.08, 0.0])
        objs.append(pd)
    else:
        m_t = F.SquarePlanoMirror2D(2 * INCH)      # catches the transmitted child
        m_t.zrotate3d_(math.radians(25)); m_t.translate3d_([0.0, 0.08, 0.0])
        objs.append(m_t)
    return dict(system=F.System(objs), bs=bs, pd=pd)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rect", "round"])
@pytest.mark.parametrize("tilt", [45.0, -30.0])
def test_plate_beamsplitter_rays(bmo, orc, kind, tilt):
    sc = _plate_scene(scenes._ProductFactory(bmo), kind, tilt)
    osc = _plate_scene(scenes._OracleFactory(), kind, tilt)
    pos, d = _fan(24, half=6e-3)
    res, b, seg = _compare_ray_trees(bmo, orc, sc["system"], osc["system"], pos, d, r_max=12)
    assert res.n_beams >= 3 * len(pos)                      # every ray splits at the coating
    assert (b["status"] == 4).sum() >= len(pos)             # BMO_ST_SPLIT


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rect", "round"])
def test_plate_beamsplitter_polarized_rays(bmo, orc, kind):
    """PolarizedRay through a plate splitter.  The reference overwrites the transmitted child's direction with the refracted
    one (`direction!(first(rays(beam.children[1])), n_d)`, PlateBeamsplitter.jl:222) without touching its E0, so at oblique
    incidence only s-polarised light (E0 perpendicular to the plane of incidence) keeps `dir . E0 = 0`; any p content makes
    the next PolarizedRay constructor throw (PolarizedRays.jl:54-56).  Compared here: s-polarised at 45 deg, a general
    (elliptical) state at normal incidence; for p content at 45 deg the oracle throws like the reference and the GPU
    reports it through the e0_warn status bit instead."""
    pos, d = _fan(12, half=5e-3)
    sc, osc = _plate_scene(scenes._ProductFactory(bmo), kind, 45.0), _plate_scene(scenes._OracleFactory(), kind, 45.0)
    res, b, seg = _compare_ray_trees(bmo, orc, sc["system"], osc["system"], pos, d, E0=np.array([0.0, 0.0, 1.0]), r_max=10)
    assert res.polarized and res.n_beams >= 3 * len(pos) and not b["e0_warn"].any()
    sc0, osc0 = _plate_scene(scenes._ProductFactory(bmo), kind, 0.0), _plate_scene(scenes._OracleFactory(), kind, 0.0)
    res, b, seg = _compare_ray_trees(bmo, orc, sc0["system"], osc0["system"], pos, d, E0=np.array([0.6, 0.0, 0.8j]), r_max=10)
    assert res.n_beams >= 3 * len(pos)
    pos2 = pos.copy(); pos2[:, 1] = 0.05        # from the substrate side: refraction, hint to the coating, split inside the glass
    res, b, seg = _compare_ray_trees(bmo, orc, sc["system"], osc["system"], pos2, -d, E0=np.array([0.0, 0.0, 1.0]), r_max=10)
    assert (b["nseg"][:len(pos)] == 2).all()
    # p content at oblique incidence: the reference throws
    ob = orc.polarized_beam(pos[0], d, 1e-6, np.array([0.6, 0.0, 0.8j]))
    with pytest.raises(orc.OracleError):
        orc.solve_system_(osc["system"], ob, r_max=10)
    bad = bmo.solve_system_(sc["system"], bmo.RayBundle(pos[:1], d, 1e-6, E0=np.array([0.6, 0.0, 0.8j])), r_max=10)
    assert bad.beams()["e0_warn"].any()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rect", "round"])
def test_plate_beamsplitter_gaussian(bmo, orc, kind):
    n = 48
    sc = _plate_scene(scenes._ProductFactory(bmo), kind, 45.0, with_pd=n)
    osc = _plate_scene(scenes._OracleFactory(), kind, 45.0, with_pd=n)
    B = dict(pos=(0.0, -0.1, 0.0), dir=(0.0, 1.0, 0.0), lam=1e-6, w0=1e-3, M2=1.0, P0=1e-3, support=(1.0, 0.0, 0.0))
    res, ref = _compare_gauss_tree(bmo, orc, sc["system"], osc["system"], B, sc["pd"], osc["pd"], r_max=10)
    assert res.n_beams >= 3


@pytest.mark.gpu
def test_plate_beamsplitter_coating_tie(bmo, orc):
    """Rays that reach the coated face from outside see the substrate surface and the coating at the same distance:
    the coating must win whenever `t_coating ~ t_substrate` (isapprox, PlateBeamsplitter.jl:176-179), from both sides of
    the plate and at normal as well as oblique incidence."""
    for kind, tilt in (("rect", 0.0), ("round", 0.0), ("rect", 20.0), ("round", -35.0)):
        sc = _plate_scene(scenes._ProductFactory(bmo), kind, tilt)
        osc = _plate_scene(scenes._OracleFactory(), kind, tilt)
        pos, d = _fan(9, half=4e-3)
        pos2 = pos.copy(); pos2[:, 1] = 0.05                 # second bundle from the other side
        for p, dd in ((pos, d), (pos2, -d)):
            res, b, seg = _compare_ray_trees(bmo, orc, sc["system"], osc["system"], p, dd, r_max=6)
            flat = bmo.upload_system(sc["system"], [1e-6]).flat
            roots_first = b["first"][:len(p)]
            first_part = seg["part"][roots_first]
            coat = [i for i, o in enumerate(flat.part_owner) if o is sc["bs"].coating][0]
            sub = [i for i, o in enumerate(flat.part_owner) if o is sc["bs"].substrate][0]
            assert set(first_part.tolist()) <= {coat, sub}
            # exactly one of the two directions meets the coated face first: there the coating wins the tie for every ray
            assert (first_part == coat).all() or (first_part == sub).all()


# ---- lenses with a meniscus ------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("radii", [(0.03, 0.05, 2e-3), (-0.06, -0.035, 2e-3)])
def test_meniscus_lens(bmo, orc, radii):
    r1, r2, ct = radii
    def build(F):
        lens = F.SphericalLens(r1, r2, ct, INCH, 1.6)
        lens.xrotate3d_(math.radians(7)); lens.translate3d_([1e-3, 0.0, -0.5e-3])
        back = F.SquarePlanoMirror2D(2 * INCH); back.translate3d_([0.0, 0.06, 0.0])
        return F.System([lens, back]), lens
    sys_, lens = build(scenes._ProductFactory(bmo))
    osys, _ = build(scenes._OracleFactory())
    assert isinstance(lens.shape, bmo.MeniscusLensSDF) or any(isinstance(s, bmo.MeniscusLensSDF) for s in getattr(lens.shape, "sdfs", []))
    pos, d = _fan(48, half=11e-3)
    res, b, seg = _compare_ray_trees(bmo, orc, sys_, osys, pos, d, r_max=12, bitwise=True)
    assert (b["nseg"] >= 4).sum() > 30       # through the lens, off the mirror, through the lens again


# ---- primitive SDFs as optical shapes --------------------------------------------------------------------
PRIMS = {
    "box": ("BoxSDF", (12e-3, 8e-3, 16e-3)),
    "cylinder": ("CylinderSDF", (7e-3, 5e-3)),
    "cutsphere": ("CutSphereSDF", (10e-3, 4e-3)),
    "ring": ("RingSDF", (4e-3, 5e-3, 6e-3)),
    "sphere": ("SphereSDF", (8e-3,)),
}


def _prim_scene(F, bmo, orc, name, as_mirror):
    kind, args = PRIMS[name]
    if isinstance(F, scenes._OracleFactory):
        shape = orc.new(kind, list(args))
        obj = orc.new("Mirror", ih=[shape]) if as_mirror else orc.new("Prism", ih=[shape, orc.refindex(1.7)])
    else:
        shape = getattr(bmo, kind)(*args)
        obj = bmo.Mirror(shape) if as_mirror else bmo.Prism(shape, 1.7)
    obj.xrotate3d_(math.radians(25)); obj.zrotate3d_(math.radians(-40)); obj.yrotate3d_(math.radians(10))
    obj.translate3d_([0.5e-3, 0.0, -0.3e-3])
    screen = F.SquarePlanoMirror2D(0.2); screen.translate3d_([0.0, 0.1, 0.0])
    return F.System([obj, screen]), obj


@pytest.mark.gpu
@pytest.mark.parametrize("as_mirror", [False, True])
@pytest.mark.parametrize("name", sorted(PRIMS))
def test_primitive_sdf_shapes(bmo, orc, name, as_mirror):
    sys_, obj = _prim_scene(scenes._ProductFactory(bmo), bmo, orc, name, as_mirror)
    osys, oobj = _prim_scene(scenes._OracleFactory(), bmo, orc, name, as_mirror)
    # same pose on both sides (SphereSDF: orientation stays I, SphericalLensSDF.jl:82-84)
    po, do = oobj.pose()
    assert np.array_equal(np.array(obj.position()), po) and np.array_equal(np.array(obj.orientation()), do)
    pos, d = _fan(64, half=11e-3)
    pos[:, 2] *= 1.7
    res, b, seg = _compare_ray_trees(bmo, orc, sys_, osys, pos, d, r_max=16, bitwise=True)
    flat = bmo.upload_system(sys_, [1e-6]).flat
    hit_prim = (seg["obj"] == flat.object_index(obj)).sum()
    assert hit_prim >= 20, hit_prim                     # the fan really probes the primitive (edges, faces, the hole of the ring)
    assert (b["nseg"] == 1).sum() < len(pos)


@pytest.mark.gpu
def test_concave_spherical_mirror(bmo, orc):
    def build(F):
        m = F.ConcaveSphericalMirror(0.1, 6e-3, INCH)
        m.zrotate3d_(math.radians(180 + 4)); m.translate3d_([0.0, 0.05, 0.0])   # concave side towards the source, slightly tilted
        catch = F.SquarePlanoMirror2D(0.1); catch.translate3d_([0.0, -0.08, 0.0])
        return F.System([m, catch])
    sys_, osys = build(scenes._ProductFactory(bmo)), build(scenes._OracleFactory())
    pos, d = _fan(48, half=10e-3, y0=-0.02)
    res, b, seg = _compare_ray_trees(bmo, orc, sys_, osys, pos, d, r_max=8, bitwise=True)
    assert (b["nseg"] >= 3).sum() > 40
    # Gaussian beamlet off the curved mirror onto a detector
    n = 40
    def build_pd(F):
        m = F.ConcaveSphericalMirror(0.1, 6e-3, INCH)
        m.zrotate3d_(math.radians(180 + 8)); m.translate3d_([0.0, 0.05, 0.0])
        pd = F.Photodetector(6e-3, n)
        pd.zrotate3d_(math.radians(16)); pd.translate3d_([0.05 * math.tan(math.radians(16)), 0.0, 0.0])
        return F.System([m, pd]), pd
    (sysg, pd), (osysg, opd) = build_pd(scenes._ProductFactory(bmo)), build_pd(scenes._OracleFactory())
    B = dict(pos=(0.0, -0.02, 0.0), dir=(0.0, 1.0, 0.0), lam=1e-6, w0=8e-4, support=(1.0, 0.0, 0.0))
    _compare_gauss_tree(bmo, orc, sysg, osysg, B, pd, opd)


@pytest.mark.gpu
def test_intersectable_object_absorbs(bmo, orc):
    """IntersectableObject (Intersectable.jl:10-15): intersect3d only, interact3d returns nothing -> the beam ends there."""
    def build(F, prod):
        if prod:
            stop = bmo.IntersectableObject(bmo.RingSDF(3e-3, 20e-3, 2e-3))      # an aperture stop: a washer with a 6 mm hole
        else:
            stop = orc.new("IntersectableObject", ih=[orc.new("RingSDF", [3e-3, 20e-3, 2e-3])])
        stop.xrotate3d_(math.radians(3))
        lens = F.SphericalLens(0.05, -0.05, 5e-3, INCH, 1.5); lens.translate3d_([0.0, 0.02, 0.0])
        return F.System([stop, lens]), stop
    (sys_, stop), (osys, _) = build(scenes._ProductFactory(bmo), True), build(scenes._OracleFactory(), False)
    pos, d = _fan(40, half=9e-3)
    res, b, seg = _compare_ray_trees(bmo, orc, sys_, osys, pos, d, r_max=10, bitwise=True)
    flat = bmo.upload_system(sys_, [1e-6]).flat
    si = flat.object_index(stop)
    last = b["first"][:len(pos)] + b["nseg"][:len(pos)] - 1
    stopped = seg["obj"][last] == si
    assert stopped.sum() >= 10 and (~stopped).sum() >= 5          # rim absorbed, centre passes through the hole
    assert (b["status"][:len(pos)][stopped] == 2).all()           # BMO_ST_ABSORBED
    assert (b["nseg"][:len(pos)][stopped] == 1).all()


@pytest.mark.gpu
def test_compensator_plate_in_michelson_arm(bmo, orc):
    """RectangularCompensatorPlate (Compensators.jl:15-24): a cuboid-mesh Prism; rays, polarised rays and a beamlet."""
    def build(F, prod, pd_n=None):
        cp = bmo.RectangularCompensatorPlate(30e-3, 20e-3, 4e-3, 1.45) if prod else orc.new("RectangularCompensatorPlate", [30e-3, 20e-3, 4e-3], [orc.refindex(1.45)])
        cp.zrotate3d_(math.radians(30))
        m = F.SquarePlanoMirror2D(2 * INCH); m.translate3d_([0.0, 0.06, 0.0])
        objs = [cp, m]
        pd = None
        if pd_n:
            pd = F.Photodetector(10e-3, pd_n); pd.translate3d_([0.0, -0.15, 0.0])
            objs.append(pd)
        return F.System(objs), pd
    (sys_, _), (osys, _) = build(scenes._ProductFactory(bmo), True), build(scenes._OracleFactory(), False)
    pos, d = _fan(24, half=6e-3)
    _compare_ray_trees(bmo, orc, sys_, osys, pos, d, r_max=10)
    _compare_ray_trees(bmo, orc, sys_, osys, pos[:8], d, E0=np.array([0.0, 0.0, 1.0]), r_max=10)
    (sysg, pd), (osysg, opd) = build(scenes._ProductFactory(bmo), True, 40), build(scenes._OracleFactory(), False, 40)
    B = dict(pos=(0.0, -0.1, 0.0), dir=(0.0, 1.0, 0.0), lam=1e-6, w0=1e-3, support=(1.0, 0.0, 0.0))
    res, ref = _compare_gauss_tree(bmo, orc, sysg, osysg, B, pd, opd)
    assert len(ref[0]["chief"]["t"]) == 6       # plate in, plate out, mirror, plate in, plate out, detector


@pytest.mark.gpu
def test_static_system_flattens_like_system(bmo, orc):
    """StaticSystem (System.jl:38-45) holds the same objects as a tuple: identical flattening, identical trace."""
    sc, osc = scenes.doublet_spot(bmo, rotate=True), scenes.doublet_spot_oracle(rotate=True)
    static = bmo.StaticSystem(tuple(sc["system"].objects))
    pos, d = scenes.fibonacci_disc(64)
    # the group was moved: bring the rays along (same transformation as scenes.doublet_spot applies to the group)
    g = sc["system"].objects[0]
    R = np.array(g.orientation()); c = np.array(g.position())
    pos_w, d_w = pos @ R.T + c, R @ d[0]
    res_a = bmo.solve_system_(sc["system"], bmo.RayBundle(pos_w, d_w, 707e-9))
    res_b = bmo.solve_system_(static, bmo.RayBundle(pos_w, d_w, 707e-9))
    sa, sb = res_a.segments(), res_b.segments()
    for key in ("pos", "dir", "t", "nrm", "n", "obj"):
        assert np.array_equal(sa[key], sb[key]), key
    _compare_ray_trees(bmo, orc, static, osc["system"], pos_w, d_w, lam=707e-9, bitwise=True)
    assert (res_b.beams()["nseg"] == 4).all()


# ---- STL meshes --------------------------------------------------------------------------------------------
def test_load_stl_matches_the_oracle_loader(bmo, orc):
    """Mesh(load(path)) (Mesh.jl:48-70): Float32 vertices scaled by Float32(1e-3), faces = consecutive triples."""
    pm = bmo.load_stl(STL)
    om = orc.load_stl(STL)
    nv, nf, f32, scale = om.eval("mesh_counts", nout=4)
    assert (int(nv), int(nf)) == (54588, 18196) == (pm.vertices.shape[0], pm.faces.shape[0])
    assert f32 == 1.0 and pm.f32 and scale == pm.scale == float(np.float32(1e-3))
    ov = om.eval("mesh_vertices", nout=3 * int(nv)).reshape(-1, 3)
    of = om.eval("mesh_faces", nout=3 * int(nf)).reshape(-1, 3).astype(np.int32)
    assert np.array_equal(pm.vertices, ov) and np.array_equal(pm.faces, of)
    assert np.array_equal(pm.vertices, pm.vertices.astype(np.float32).astype(np.float64))     # values are Float32
    # kinematics of a Float32 mesh round to Float32 on both sides
    pm.translate3d_([0.1, -0.02, 0.033]); pm.rotate3d_([0.0, 0.0, 1.0], 0.7)
    om.translate3d_([0.1, -0.02, 0.033]); om.rotate3d_([0.0, 0.0, 1.0], 0.7)
    ov = om.eval("mesh_vertices", nout=3 * int(nv)).reshape(-1, 3)
    assert np.array_equal(pm.vertices, ov)


@pytest.mark.gpu
def test_c4_with_the_reference_stl_and_unscaled_rhomb(bmo, orc):
    """BASELINE config 4 as specified: the Fresnel rhomb of test/runtests.jl:2339-2349 at full size, a thin beamsplitter,
    a Retroreflector and Mirror(Mesh(load("Mirror_Post.stl"))) (18 196 triangles behind the BVH), PolarizedRays."""
    sc, osc = s2.mesh_scene_c4(bmo, STL), s2.mesh_scene_c4_oracle(STL)
    n = 64
    pos, d, E0 = s2.jittered_lattice(n, half=s2.C4_HALF, y0=s2.C4_Y0)
    res, b, seg = _compare_ray_trees(bmo, orc, sc["system"], osc["system"], pos, d, E0=E0, r_max=100)
    flat = bmo.upload_system(sc["system"], [1e-6]).flat
    assert (seg["obj"] == flat.object_index(sc["post"])).sum() >= 8      # the STL mirror is really hit
    assert (seg["obj"] == flat.object_index(sc["rhomb"])).sum() >= 4 * n // 2
    assert bmo.counters()["tri_tests"] > 0
    # the BVH agrees with the reference's face loop bit for bit on this mesh: direct bundle at the post
    post = bmo.Mirror(bmo.load_stl(STL)); opost = orc.new("Mirror", ih=[orc.load_stl(STL)])
    rng = np.random.default_rng(3)
    m = 2048
    p2 = np.zeros((m, 3)); p2[:, 0] = (rng.random(m) - 0.5) * 0.07; p2[:, 2] = rng.random(m) * 0.11 - 0.026; p2[:, 1] = -0.2
    d2 = np.tile([0.0, 1.0, 0.0], (m, 1)); d2[:, [0, 2]] += (rng.random((m, 2)) - 0.5) * 0.04
    bundle = bmo.RayBundle(p2, d2, 1e-6)
    r2 = bmo.solve_system_(bmo.System([post]), bundle, r_max=2)
    ref = orc.bulk_trace_rays(orc.system([opost]), bundle.pos, bundle.dir, 1e-6, r_max=2, max_seg=3)
    b2, g2 = r2.beams(), r2.segments()
    assert np.array_equal(b2["nseg"], ref["nseg"])
    t_gpu, t_ref = g2["t"][b2["first"]], ref["seg"][:, 0, 7]
    hit = np.isfinite(t_ref)
    assert hit.sum() > m // 4 and np.array_equal(np.isfinite(t_gpu), hit)
    assert np.array_equal(t_gpu[hit], t_ref[hit])
    assert np.array_equal(g2["nrm"][b2["first"]][hit], ref["seg"][:, 0, 8:11][hit])
