"""More known-answer tests of the reference (test/runtests.jl) re-expressed against the CPU oracle -- and, for the ones
with the tightest tolerances, against the CUDA path as well (same scene through the product's host mirror and the C ABI).
Each test names the line range it follows; where a loop was shortened the physics is unchanged and the reduction is stated."""
import math

import numpy as np
import pytest

from tests import scenes, scenes2

INCH = 25.4e-3
MM = 1e-3


# ---- Issue #22 / #23 (:2831-2908): phase of the field through a substrate of varying refractive index -------------------
# CubeBeamsplitter(10 mm, n = 1), a refractive cylinder (r = 5 mm, length 10 mm) in the probe arm, Photodetector(10 mm, 250).
# Raising the index from 1 to 1 + lambda / L adds exactly one wavelength of optical path: the detector power must follow
# (cos(phi) + 1) / 2 * 2 mW to 1e-8 W -- the tightest pin the reference holds for electric_field + optical_power.
# (The reference draws a random support vector perpendicular to the beam; here it is fixed: x for the probe, z for the reference arm.)
def _issue22_scene(F, n_substrate, make_substrate):
    splitter = F.CubeBeamsplitter(10 * MM, 1.0)
    substrate = make_substrate(n_substrate)
    detector = F.Photodetector(10 * MM, 250)
    substrate.translate3d_([0.0, -25 * MM, 0.0])
    detector.translate3d_([0.0, 40 * MM, 0.0])
    return F.System([substrate, splitter, detector]), detector


def _ref_signal(phi, A):
    return (math.cos(phi) + 1) / 2 * A


N_STEPS = 13        # the reference scans 50 index values; 13 cover the same period


def test_issue22_index_phase_shift_oracle(orc):
    lam, L = 1e-6, 10 * MM
    F = scenes._OracleFactory()
    for nf in np.linspace(0, 1, N_STEPS):
        n_sub = 1 + lam / L * nf
        system, det = _issue22_scene(F, n_sub, lambda n: orc.new("Prism", ih=[orc.new("CylinderSDF", [5 * MM, L / 2]), orc.refindex(n)]))
        prb = orc.gaussian_beamlet([0, -50 * MM, 0], [0, 1, 0], lam, 0.5 * MM)
        ref = orc.gaussian_beamlet([50 * MM, 0, 0], [-1, 0, 0], lam, 0.5 * MM, support=(0.0, 0.0, 1.0))
        orc.solve_system_(system, prb)
        orc.solve_system_(system, ref)
        assert abs(det.pd_power() - _ref_signal(2 * math.pi * nf, 2e-3)) <= 1e-8
    # the optical path lengths of the two arms differ by exactly one wavelength at the end of the scan (:2889-2893)
    opl = lambda g: max(b["opl"] for b in orc.gauss_export(system, g))
    assert abs((opl(prb) - opl(ref)) / lam - 1) <= 1e-9


@pytest.mark.gpu
def test_issue22_index_phase_shift_gpu(bmo):
    lam, L = 1e-6, 10 * MM
    F = scenes._ProductFactory(bmo)
    for nf in np.linspace(0, 1, N_STEPS):
        n_sub = 1 + lam / L * nf
        system, det = _issue22_scene(F, n_sub, lambda n: bmo.Prism(bmo.CylinderSDF(5 * MM, L / 2), n))
        prb = bmo.GaussianBeamlet([0, -50 * MM, 0], [0, 1, 0], lam, 0.5 * MM)
        ref = bmo.GaussianBeamlet([50 * MM, 0, 0], [-1, 0, 0], lam, 0.5 * MM, support=(0.0, 0.0, 1.0))
        bmo.solve_system_(system, prb)
        bmo.solve_system_(system, ref)
        assert abs(det.optical_power() - _ref_signal(2 * math.pi * nf, 2e-3)) <= 1e-8


def test_issue22_visibility_is_one_for_any_substrate_index(orc):
    """:2863-2880 (4 indices x 30 phases -> 2 indices x 7 phases of the reference beam, set through its start position: a shift
    of the reference beamlet's origin by d along its axis shifts its phase by 2 pi d / lambda, which is what shift_phase does)."""
    lam, L = 1e-6, 10 * MM
    F = scenes._OracleFactory()
    for index in (1.0, 100.0):
        system, det = _issue22_scene(F, index, lambda n: orc.new("Prism", ih=[orc.new("CylinderSDF", [5 * MM, L / 2]), orc.refindex(n)]))
        pw = []
        for k in range(7):
            det.pd_empty()
            prb = orc.gaussian_beamlet([0, -50 * MM, 0], [0, 1, 0], lam, 0.5 * MM)
            ref = orc.gaussian_beamlet([50 * MM + lam * k / 6, 0, 0], [-1, 0, 0], lam, 0.5 * MM, support=(0.0, 0.0, 1.0))
            orc.solve_system_(system, prb)
            orc.solve_system_(system, ref)
            pw.append(det.pd_power())
        vis = (max(pw) - min(pw)) / (max(pw) + min(pw))
        assert abs(vis - 1) <= 1e-2


# ---- Issue #14 (:2814-2829): oblique Gaussian onto a rotated detector keeps its power ---------------------------------------
def test_issue14_power_on_rotated_detector(orc):
    pd = orc.new("Photodetector", [10e-3], [250])          # 1000 px in the reference; 250 resolve the 2.5 mm waist equally well
    pd.zrotate3d_(math.radians(90)); pd.translate3d_([0.46, 0.0, 0.0])
    g = orc.gaussian_beamlet([0, 0.2, 0], [0.46, -0.2, 0], 532e-9, 2.5e-3, P0=10e-3)
    orc.solve_system_(orc.system([pd]), g)
    assert abs(pd.pd_power() - 10e-3) <= 1e-5


# ---- Fresnel rhomb (:2339-2362): 45 deg linear -> circular after two total internal reflections -------------------------------
def _rhomb(F):
    s1 = F.CuboidMesh(0.5, 1.25, 0.5, math.radians(53.3))
    l1 = F.LensFromMesh(s1, 1.5)
    l1.translate3d_([-0.25, 0.0, -0.25])
    s1.set_new_origin3d_()
    l1.yrotate3d_(math.radians(135))
    return l1


def test_fresnel_rhomb_quarter_wave_oracle(orc):
    F = scenes2.OracleFactory2()
    system = F.System([_rhomb(F)])
    E = math.sqrt(2 * 1.0 * 376.730313668)                 # electric_field(1) = sqrt(2 I Z0), I = 1
    b = orc.polarized_beam([0, -1, 0], [0, 1, 0], 1000e-9, [0, 0, E])
    orc.solve_system_(system, b)
    rays = orc.beam_export(system, b)[0]["rays"]
    E0 = rays["E0"][-1]
    assert rays["pos"].shape[0] == 5                       # in, two total internal reflections, out
    phi = np.angle(E0[2]) - np.angle(E0[0])
    assert abs(phi - math.pi / 2) <= 1.5e-8 * math.pi      # Julia's `phi ≈ π/2`: rtol = sqrt(eps)
    assert abs(E0[1]) < 2e-14


@pytest.mark.gpu
def test_fresnel_rhomb_quarter_wave_gpu(bmo):
    F = scenes2.ProductFactory2(bmo)
    system = F.System([_rhomb(F)])
    E = math.sqrt(2 * 1.0 * 376.730313668)
    beam = bmo.Beam([0, -1, 0], [0, 1, 0], 1000e-9, E0=[0, 0, E])
    res = bmo.solve_system_(system, beam)
    E0 = res.segments()["E0"][-1]
    assert res.n_segments == 5
    assert abs(np.angle(E0[2]) - np.angle(E0[0]) - math.pi / 2) <= 1.5e-8 * math.pi
    assert abs(E0[1]) < 2e-14


# ---- Mach-Zehnder interferometer (:2364-2436): sign table of the Jones matrices of the thin splitters -------------------------
def _mzi(F):
    m1, m2 = F.SquarePlanoMirror2D(INCH), F.SquarePlanoMirror2D(INCH)
    b1, b2 = F.ThinBeamsplitter(INCH, reflectance=0.5), F.ThinBeamsplitter(INCH, reflectance=0.5)
    b1.translate3d_([0.0, 0.0, 0.0]); b2.translate3d_([2 * INCH, 2 * INCH, 0.0])
    m1.translate3d_([0.0, 2 * INCH, 0.0]); m2.translate3d_([2 * INCH, 0.0, 0.0])
    b1.zrotate3d_(math.radians(360 - 135)); b2.zrotate3d_(math.radians(45))
    m1.zrotate3d_(math.radians(360 - 135)); m2.zrotate3d_(math.radians(45))
    return F.System([m1, m2, b1, b2])


def _mzi_table(tree):
    """tree: beams in BFS order with parents -> E0 of t, r, tr, rr, trt, trr, rrt, rrr (children = [transmitted, reflected])."""
    kids = {}
    for i, b in enumerate(tree):
        kids.setdefault(b["parent"], []).append(i)
    root = kids[-1][0]
    t, r = kids[root]
    e = lambda i, k: tree[i]["rays"]["E0"][k]
    (trt, trr), (rrt, rrr) = kids[t], kids[r]
    return dict(t=e(t, 0), r=e(r, 0), tr=e(t, 1), rr=e(r, 1), trt=e(trt, 0), trr=e(trr, 0), rrt=e(rrt, 0), rrr=e(rrr, 0))


def _check_mzi(tab_z, tab_x):
    s = math.sqrt(2) / 2
    close = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) <= 1.5e-8 * max(np.linalg.norm(a), np.linalg.norm(b), 1e-300)
    assert abs(tab_z["t"][2] - s) <= 1e-8 and abs(tab_z["r"][2] + s) <= 1e-8
    assert abs(tab_z["tr"][2] + s) <= 1e-8 and abs(tab_z["rr"][2] - s) <= 1e-8
    assert close(tab_z["trt"], tab_z["rrr"]) and close(tab_z["trr"], tab_z["rrt"])
    assert abs(tab_x["t"][0] - s) <= 1e-8 and abs(tab_x["r"][1] + s) <= 1e-8
    assert close(tab_x["tr"], tab_x["r"]) and close(tab_x["rr"], tab_x["t"])
    assert close(tab_x["trt"], tab_x["rrr"]) and close(tab_x["trr"], tab_x["rrt"])


def test_mzi_sign_table_oracle(orc):
    F = scenes._OracleFactory()
    system = _mzi(F)
    tabs = []
    for E0 in ([0, 0, 1], [1, 0, 0]):
        b = orc.polarized_beam([0, -0.1, 0], [0, 1, 0], 1000e-9, E0)
        orc.solve_system_(system, b)
        tree = orc.beam_export(system, b)
        assert sum(1 for i, x in enumerate(tree) if not any(y["parent"] == i for y in tree)) == 4     # four leaves (:2414)
        tabs.append(_mzi_table(tree))
    _check_mzi(*tabs)


@pytest.mark.gpu
def test_mzi_sign_table_gpu(bmo):
    F = scenes._ProductFactory(bmo)
    system = _mzi(F)
    tabs = []
    for E0 in ([0, 0, 1], [1, 0, 0]):
        res = bmo.solve_system_(system, bmo.Beam([0, -0.1, 0], [0, 1, 0], 1000e-9, E0=E0))
        b, seg = res.beams(), res.segments()
        tree = []
        for i in res.bfs_order():
            f, k = int(b["first"][i]), int(b["nseg"][i])
            tree.append(dict(id=int(i), parent=int(b["parent"][i]), rays=dict(E0=seg["E0"][f:f + k])))
        ids = {t["id"]: j for j, t in enumerate(tree)}
        for t in tree:
            t["parent"] = ids.get(t["parent"], -1)
        tabs.append(_mzi_table(tree))
    _check_mzi(*tabs)


# ---- Issue #22 / #23, third part (:2895-2908): the start phase of a beamlet is changed between solves that retrace ------------
def test_issue22_phase_mutation_with_retrace_oracle(orc):
    lam, L = 1e-6, 10 * MM
    F = scenes._OracleFactory()
    system, det = _issue22_scene(F, 1 + lam / L, lambda n: orc.new("Prism", ih=[orc.new("CylinderSDF", [5 * MM, L / 2]), orc.refindex(n)]))
    prb = orc.gaussian_beamlet([0, -50 * MM, 0], [0, 1, 0], lam, 0.5 * MM)
    ref = orc.gaussian_beamlet([50 * MM, 0, 0], [-1, 0, 0], lam, 0.5 * MM, support=(0.0, 0.0, 1.0))
    phis = np.linspace(0, 2 * math.pi, N_STEPS)
    step = phis[1] - phis[0]
    for k, phi in enumerate(phis):
        det.pd_empty()
        orc.solve_system_(system, prb, retrace=True)
        orc.solve_system_(system, ref, retrace=True)
        assert abs(det.pd_power() - _ref_signal(phi, 2e-3)) <= 1e-8, k
        prb.eval("gauss_scale_E0", [math.cos(step), math.sin(step)])


@pytest.mark.gpu
def test_issue22_phase_mutation_with_retrace_gpu(bmo):
    lam, L = 1e-6, 10 * MM
    F = scenes._ProductFactory(bmo)
    system, det = _issue22_scene(F, 1 + lam / L, lambda n: bmo.Prism(bmo.CylinderSDF(5 * MM, L / 2), n))
    prb = bmo.GaussianBeamlet([0, -50 * MM, 0], [0, 1, 0], lam, 0.5 * MM)
    ref = bmo.GaussianBeamlet([50 * MM, 0, 0], [-1, 0, 0], lam, 0.5 * MM, support=(0.0, 0.0, 1.0))
    phis = np.linspace(0, 2 * math.pi, N_STEPS)
    step = phis[1] - phis[0]
    for k, phi in enumerate(phis):
        det.empty_()
        bmo.solve_system_(system, prb, retrace=True)
        bmo.solve_system_(system, ref, retrace=True)          # unchanged between the steps: re-validated on the device (bmo_retrace)
        assert abs(det.optical_power() - _ref_signal(phi, 2e-3)) <= 1e-8, k
        prb.E0 = prb.E0 * complex(math.cos(step), math.sin(step))


# ---- power conservation behind a thin splitter with misaligned detectors, retracing (:2168-2215) ---------------------------
def _power_scene(F):
    bs = F.ThinBeamsplitter(10e-3)
    pd1, pd2 = F.Photodetector(10e-3, 100), F.Photodetector(10e-3, 100)
    bs.zrotate3d_(math.radians(45))
    pd1.translate3d_([0.0, 0.1, 0.0]); pd1.zrotate3d_(math.radians(180))
    pd2.translate3d_([0.1, 0.0, 0.0]); pd2.zrotate3d_(math.radians(90))
    bs.zrotate3d_(math.radians(0.017)); pd1.zrotate3d_(math.radians(10)); pd1.xrotate3d_(math.radians(15))
    return F.System([bs, pd1, pd2]), pd1, pd2


def test_power_conservation_with_retrace_oracle(orc):
    P0, l0, w0, lam = 0.5, 0.1, 0.5e-3, 1064e-9
    system, pd1, pd2 = _power_scene(scenes._OracleFactory())
    l1 = orc.gaussian_beamlet([0, -l0, 0], [0, 1, 0], lam, w0, P0=P0)
    l2 = orc.gaussian_beamlet([-l0, 0, 0], [1, 0, 0], lam, w0, P0=P0, support=(0.0, 0.0, 1.0))
    phis = np.linspace(0, 2 * math.pi, 9)          # 25 phases in the reference
    step = phis[1] - phis[0]
    p1 = []
    for phi in phis:
        pd1.pd_empty(); pd2.pd_empty()
        orc.solve_system_(system, l1, retrace=True)
        orc.solve_system_(system, l2, retrace=True)
        p1.append(pd1.pd_power())
        assert abs(pd1.pd_power() + pd2.pd_power() - 2 * P0) < 1e-4
        l1.eval("gauss_scale_E0", [math.cos(step), math.sin(step)])
    assert max(p1) - min(p1) > 0.5 * P0               # the two outputs do trade power as the phase moves


@pytest.mark.gpu
def test_power_conservation_with_retrace_gpu(bmo):
    P0, l0, w0, lam = 0.5, 0.1, 0.5e-3, 1064e-9
    system, pd1, pd2 = _power_scene(scenes._ProductFactory(bmo))
    l1 = bmo.GaussianBeamlet([0, -l0, 0], [0, 1, 0], lam, w0, P0=P0)
    l2 = bmo.GaussianBeamlet([-l0, 0, 0], [1, 0, 0], lam, w0, P0=P0, support=(0.0, 0.0, 1.0))
    phis = np.linspace(0, 2 * math.pi, 9)
    step = phis[1] - phis[0]
    for phi in phis:
        pd1.empty_(); pd2.empty_()
        bmo.solve_system_(system, l1, retrace=True)
        bmo.solve_system_(system, l2, retrace=True)
        assert abs(pd1.optical_power() + pd2.optical_power() - 2 * P0) < 1e-4
        l1.E0 = l1.E0 * complex(math.cos(step), math.sin(step))


# ---- Double Gauss lens (:2693-2762): six elements in nested ObjectGroups, moved and rotated; spot radii at the back focus -----
def _double_gauss(F, moved=True):
    l1 = F.SphericalLens(48.88e-3, 182.96e-3, 8.89e-3, 52.3e-3, 1.62286)
    l23 = F.SphericalDoubletLens(36.92e-3, math.inf, 23.06e-3, 15.11e-3, 2.31e-3, 45.11e-3, 1.58565, 1.67764)
    l45 = F.SphericalDoubletLens(-23.91e-3, math.inf, -36.92e-3, 1.92e-3, 7.77e-3, 40.01e-3, 1.57046, 1.64128)
    l6 = F.SphericalLens(1063.24e-3, -48.88e-3, 6.73e-3, 45.11e-3, 1.62286)
    thick = lambda o: o.thickness() if hasattr(o, "thickness") else float(o.eval("thickness_object")[0])
    l_23 = thick(l1) + 0.38e-3
    l_45 = l_23 + thick(l23) + 9.14e-3 + 13.36e-3
    l_6 = l_45 + thick(l45) + 0.38e-3
    f_z = l_6 + thick(l6) + 58.21e-3 + 7e-4
    l23.translate3d_([0.0, l_23, 0.0]); l45.translate3d_([0.0, l_45, 0.0]); l6.translate3d_([0.0, l_6, 0.0])
    dg = F.ObjectGroup([l1, l23, l45, l6])
    det = F.Spotdetector(5e-3)
    det.translate3d_([0.0, f_z, 0.0])
    setup = F.ObjectGroup([dg, det])
    if moved:
        setup.translate3d_([0.05, 0.05, 0.05]); setup.xrotate3d_(math.radians(60)); setup.zrotate3d_(math.radians(45))
    else:
        det.translate_to3d_([0.0, 0.147, 0.0])
    return dict(system=F.System([setup]), det=det, l1=l1, dg=dg)


class _OracleFactoryDG(scenes._OracleFactory):
    def SphericalDoubletLens(self, r1, r2, r3, l1, l2, d, n1, n2):
        return self.orc.new("SphericalDoubletLens", [r1, r2, r3, l1, l2, d], [self.orc.refindex(n1), self.orc.refindex(n2)])


def _dg_sources(bmo, sc_product):
    """The three ray bundles of the reference test, generated once (product's source constructors, BeamGroups.jl)."""
    dirv = np.array(sc_product["dg"].orientation())[:, 1]
    pos = np.array(sc_product["l1"].position()) - 0.05 * dirv
    return bmo.CollimatedSource(pos, dirv, 0.04, 486.0e-9, num_rays=1000, num_rings=10)


def _radii(xz):
    return np.hypot(xz[:, 0], xz[:, 1])


def test_double_gauss_spot_radii_oracle(bmo, orc):
    src = _dg_sources(bmo, _double_gauss(scenes._ProductFactory(bmo)))
    osc = _double_gauss(_OracleFactoryDG())
    out = orc.bulk_trace_rays(osc["system"], src.pos, np.broadcast_to(src.dir, src.pos.shape).copy(), 486.0e-9, max_seg=16, spot=osc["det"])
    assert (out["nseg"] == 11).all()                      # 10 refracting surfaces (two singlets, two cemented doublets) + the detector
    assert _radii(out["spot"]).max() <= 2e-5
    # back at the origin: point sources, wide (2 deg) and narrow (5e-5 rad: regression of issue #11)
    osc = _double_gauss(_OracleFactoryDG(), moved=False)
    for theta, tol in ((math.radians(2), 6e-5), (5e-5, 2e-7)):
        ps = bmo.PointSource([0, -0.5, 0], [0, 1, 0], theta, 486.0e-9, num_rays=1000, num_rings=10)
        out = orc.bulk_trace_rays(osc["system"], ps.pos, ps.dir, 486.0e-9, max_seg=16, spot=osc["det"])
        assert np.isfinite(out["spot"]).all() and _radii(out["spot"]).max() <= tol


@pytest.mark.gpu
def test_double_gauss_spot_radii_gpu(bmo, orc):
    sc = _double_gauss(scenes._ProductFactory(bmo))
    src = _dg_sources(bmo, sc)
    res = bmo.solve_system_(sc["system"], src)
    assert (res.beams()["nseg"] == 11).all()
    assert _radii(sc["det"].data).max() <= 2e-5
    osc = _double_gauss(_OracleFactoryDG())
    ref = orc.bulk_trace_rays(osc["system"], src.pos, np.broadcast_to(src.dir, src.pos.shape).copy(), 486.0e-9, max_seg=16, spot=osc["det"])
    assert np.array_equal(sc["det"].data, ref["spot"])     # plain rays through unions and doublets, rotated: bit for bit
    sc = _double_gauss(scenes._ProductFactory(bmo), moved=False)
    for theta, tol in ((math.radians(2), 6e-5), (5e-5, 2e-7)):
        sc["det"].empty_()
        bmo.solve_system_(sc["system"], bmo.PointSource([0, -0.5, 0], [0, 1, 0], theta, 486.0e-9, num_rays=1000, num_rings=10))
        assert len(sc["det"].data) == 1000 and _radii(sc["det"].data).max() <= tol


# ---- Gaussian beamlet through a lens against the ABCD / complex-q formalism (:1870-1931) -------------------------------------
def test_gaussian_through_lens_vs_abcd(orc):
    lam, w0, nl, R1, R2, y_lens = 1000e-9, 1e-3, 1.5, 1.0, 1.0, 0.1
    zr = math.pi * w0 ** 2 / lam
    f = 1 / ((nl - 1) * (1 / R1 + 1 / R2))                 # lensmakers_eq(R1, -R2, nl), thin lens
    dy = 0.001
    ys = np.arange(0, 1.5 + dy / 2, dy)
    q = complex(0, zr)
    w_ana, R_ana = np.zeros(len(ys)), np.zeros(len(ys))
    for i in range(len(ys)):
        w_ana[i] = math.sqrt(-lam / (math.pi * (1 / q).imag))
        R_ana[i] = (1 / q).real
        if abs((i + 1) * dy - y_lens) < 1e-12:             # the reference applies the lens in place of one propagation step
            q = q / (-q / f + 1)
            continue
        q = q + dy
    lens = orc.new("Lens", ih=[orc.new("ThinLensSDF", [R1, R2, 0.025]), orc.refindex(nl)])
    lens.translate3d_([0, y_lens, 0])
    g = orc.gaussian_beamlet([0, 0, 0], [0, 1, 0], lam, w0, M2=1.0, support=(1.0, 0.0, 0.0))
    orc.solve_system_(orc.system([lens]), g)
    num = np.array([g.eval("gauss_parameters", [y]) for y in ys])          # w, R, psi, w0 along the beam
    assert np.abs(num[:, 0] - w_ana).max() <= 1e-6                        # beam radius within 1 um
    assert (np.abs(num[:, 1] - R_ana) <= 1e-2).mean() > 0.95 and not np.isnan(num[:, 1]).any()
    i_min = int(np.argmin(w_ana))
    assert abs(num[0, 2]) <= 1e-3 and abs(num[i_min, 2]) <= 1e-3          # Gouy phase zero at both waists
    assert abs(num[i_min, 3] - w_ana[i_min]) <= 1e-7                      # local waist after the lens


# ---- two co-propagating beamlets through a thin lens onto a small detector: power vs start offset (:2038-2068) --------------
def test_lambda_phase_scan_through_lens(orc):
    w0, lam, M2, P0 = 0.01e-3, 1000e-9, 1.0, 1e-3
    z, l, n = 0.1, 1e-2, 1000
    R1 = R2 = d = 0.01
    nl = 1.5
    f = 1 / ((nl - 1) * (1 / R1 + 1 / R2))
    pd_s = orc.new("Photodetector", [l / 10], [n // 10])
    ln = orc.new("ThinLens", [R1, R2, d], [orc.refindex(nl)])
    t_ln = float(ln.eval("thickness_object")[0])
    pd_s.translate3d_([0, z, 0])
    ln.translate3d_([0, z - f - t_ln / 2, 0])
    system = orc.system([pd_s, ln])
    dzs = np.linspace(0, lam, 13)                           # 50 offsets in the reference
    for dz in dzs:
        pd_s.pd_empty()
        g1 = orc.gaussian_beamlet([0, 0, 0], [0, 1, 0], lam, w0, M2=M2, P0=P0)
        g2 = orc.gaussian_beamlet([0, dz, 0], [0, 1, 0], lam, w0, M2=M2, P0=P0)
        orc.solve_system_(system, g1)
        orc.solve_system_(system, g2)
        len1, opl1 = g1.eval("gauss_length")
        len2, _ = g2.eval("gauss_length")
        assert math.isclose(len1, z, rel_tol=1.5e-8)
        assert math.isclose(len1, opl1 - t_ln * (nl - 1), rel_tol=1.5e-8)
        assert math.isclose(len2, len1 - dz, rel_tol=1.5e-8)
        p_ana = 4 * P0 * (math.cos(2 * math.pi * dz / lam) + 1) / 2
        assert abs(pd_s.pd_power() - p_ana) <= 1e-4


# ---- unequal-arm Michelson: real and imaginary part of the detector field against the closed form (:2122-2166) ---------------
def test_unequal_arm_michelson_field(orc):
    l_0, lam, w0, P0, M2 = 0.1, 635e-9, 1e-4, 1e-3, 1.0
    pd_size, pd_n = INCH / 5, 100
    m1, m2 = orc.new("SquarePlanoMirror2D", [INCH]), orc.new("SquarePlanoMirror2D", [INCH])
    bs = orc.new("ThinBeamsplitter", [INCH, INCH, 0.5]); pd = orc.new("Photodetector", [pd_size], [pd_n])
    m1.translate3d_([l_0, 0, 0]); m2.translate3d_([0, l_0, 0]); pd.translate3d_([-l_0, 0, 0])
    bs.zrotate3d_(math.radians(45)); m1.zrotate3d_(math.radians(90)); pd.zrotate3d_(math.radians(90))
    system = orc.system([m1, m2, bs, pd])
    dl = 1 * l_0
    m2.translate_to3d_([0, l_0 + dl, 0])
    g = orc.gaussian_beamlet([0, -l_0, 0], [0, 1, 0], lam, w0, M2=M2, P0=P0)
    orc.solve_system_(system, g)
    field = pd.pd_field(pd_n)
    # closed form: two Gaussians that travelled 4 l_0 and 4 l_0 + 2 dl, the longer one with a pi flip; amplitude E0 / sqrt(2)^2
    I0 = 2 * P0 / (math.pi * w0 ** 2)
    E0 = math.sqrt(2 * I0 * 376.730313668) / 2
    zr = math.pi * w0 ** 2 / lam / M2
    k = 2 * math.pi / lam

    def efield(r, zz):
        w = w0 * np.sqrt(1 + (zz / zr) ** 2)
        R = zz / (zz ** 2 + zr ** 2)
        psi = -math.atan(zz / zr)
        return E0 * w0 / w * np.exp(-r ** 2 / w ** 2) * np.exp(1j * (k * zz + psi + (k * r ** 2 * R) / 2))
    xs = np.linspace(-pd_size / 2, pd_size / 2, pd_n)
    r = np.hypot(xs[:, None], xs[None, :])
    screen = efield(r, 4 * l_0) + efield(r, 4 * l_0 + 2 * dl) * np.exp(1j * math.pi)
    assert np.abs(screen.real - field.real).max() <= 5e-2 and np.abs(screen.imag - field.imag).max() <= 5e-2
    assert np.abs(field).max() > 50.0                      # V/m: the comparison is not between two zeros
