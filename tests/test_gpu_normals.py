"""normal3d of the lens primitives (AbstractSDF.jl:79-95): the gradients written out for unrotated primitives
(identity_gradient, csrc/bmo_geom.cuh) against the generic dual-number evaluation with ForwardDiff's rules, bit for bit,
through bmo_debug_normals.  Random points, points on the axis (zero-vector norm rule), on faces, rims and edges."""
import ctypes as C
import math

import numpy as np
import pytest

from tests import scenes

INCH = 25.4e-3


def _system(bmo):
    """Unrotated lenses of every spherical kind, a cylinder mirror, a ball lens; and the same doublet rotated (generic path)."""
    dl = bmo.SphericalDoubletLens(*scenes.AC254, 1.6456, 1.7168)                       # plano + convex + convex(rotated pi) | plano + concave
    pcx = bmo.SphericalLens(0.05, math.inf, 5e-3, INCH, 1.5); pcx.translate3d_([0.04, 0.0, 0.0])
    bcc = bmo.SphericalLens(-0.06, 0.08, 3e-3, INCH, 1.5); bcc.translate3d_([-0.04, 0.01, 0.0])   # concave + plano + concave(rotated)
    mir = bmo.RoundPlanoMirror(INCH, 5e-3); mir.translate3d_([0.0, 0.08, 0.02])
    ball = bmo.Prism(bmo.SphereSDF(4e-3), 1.5); ball.translate3d_([0.02, -0.03, 0.01])
    rod = bmo.Mirror(bmo.CylinderSDF(7e-3, 5e-3)); rod.translate3d_([-0.03, -0.04, 0.02])
    rot = bmo.SphericalLens(0.05, -0.07, 6e-3, INCH, 1.5); rot.translate3d_([0.0, 0.15, 0.0]); rot.xrotate3d_(0.3)
    return bmo.System([dl, pcx, bcc, mir, ball, rod, rot])


def _points(rng, prim, n):
    """World points around one primitive record: random ones near the surface scale, the axis, faces, rims, edges."""
    pos = np.array(prim.pos[:])
    par = np.array(prim.par[:])
    scale = float(np.max(np.abs(par[np.isfinite(par) & (np.abs(par) < 1.0)]), initial=1e-2))
    radial = [0.0, scale / 2, scale / 4, par[1] / 2 if abs(par[1]) < 1 else scale, par[0] if abs(par[0]) < 1 else scale, 1e-300, 1e-9]
    axial = [0.0, par[0], par[0] / 2, -par[0], par[2] if abs(par[2]) < 1 else 0.0, -par[2] / 2 if abs(par[2]) < 1 else 0.0, par[3] if abs(par[3]) < 1 else 0.0, 1e-12, -1e-12]
    pts = [rng.normal(size=(n, 3)) * scale]
    special = []
    for r in radial:
        for y in axial:
            for ang in (0.0, 0.7, math.pi / 2, math.pi, 4.0):
                special.append([r * math.cos(ang), y, r * math.sin(ang)])
                special.append([r * math.cos(ang) * (1 + 1e-13), y * (1 - 1e-13), r * math.sin(ang)])
    pts.append(np.array(special))
    pts.append(rng.normal(size=(n // 4, 3)) * scale * np.array([0.0, 1.0, 0.0]))      # on the axis
    pts.append(rng.normal(size=(n // 4, 3)) * scale * np.array([1.0, 0.0, 1.0]))      # in the plane y = 0
    return np.concatenate(pts) + pos


@pytest.mark.gpu
def test_written_out_gradients_equal_the_dual_number_evaluation(bmo):
    from bmo_b200 import _lib as L
    dsys = bmo.upload_system(_system(bmo), [1e-6])
    flat = dsys.flat
    rng = np.random.default_rng(11)
    pts, idx = [], []
    for i in range(flat.n_prims):
        p = _points(rng, flat._prims[i], 4000)
        pts.append(p); idx.append(np.full(len(p), i, np.int32))
    pts, idx = np.ascontiguousarray(np.concatenate(pts)), np.ascontiguousarray(np.concatenate(idx))
    n = len(pts)
    fast, gen = np.zeros((n, 3)), np.zeros((n, 3))
    L.check(L.lib().bmo_debug_normals(dsys.h, n, L.ptr(pts), L.ptr(idx), L.ptr(fast), L.ptr(gen)))
    assert np.isfinite(gen).all(axis=1).mean() > 0.95
    bad = ~((fast == gen) | (np.isnan(fast) & np.isnan(gen))).all(axis=1)
    assert not bad.any(), (int(bad.sum()), pts[bad][:5], idx[bad][:5], fast[bad][:5], gen[bad][:5])
    types = {flat._prims[i].type for i in range(flat.n_prims)}
    assert {0, 1, 2, 3, 4} <= types                          # BMO_PRIM_PLANO, CYLINDER, SPHERE, CONVEX, CONCAVE
    # error behaviour of the entry point
    rc = L.lib().bmo_debug_normals(dsys.h, 1, L.ptr(pts), L.ptr(np.array([flat.n_prims], np.int32)), L.ptr(fast), L.ptr(gen))
    assert rc == -1
