"""Scenes of BASELINE.json's configs, built twice: through the product's host mirror (`m` =
beamletoptics.jl_b200) and through the oracle's constructor-level API -- the two share no code."""
import math

import numpy as np

INCH = 25.4e-3
N_NLAK22_707, N_NSF10_707 = 1.6456, 1.7168      # test/runtests.jl:1275-1277 at 707 nm
AC254 = (87.9e-3, -105.6e-3, math.inf, 6e-3, 3e-3, INCH)   # test/runtests.jl:1281-1282
BFL = 143.68e-3


# ---- C2: AC254-150-AB doublet + Spotdetector at the vendor back focus ---------------------------
def doublet_spot(m, rotate=False):
    dl = m.SphericalDoubletLens(*AC254, N_NLAK22_707, N_NSF10_707)
    sd = m.Spotdetector(5e-3)
    sd.translate3d_((0.0, dl.thickness() + BFL, 0.0))
    objs = [dl, sd]
    if rotate:
        g = m.ObjectGroup(objs)
        g.translate3d_((0.05, 0.05, 0.05)); g.xrotate3d_(math.radians(-60)); g.zrotate3d_(math.radians(45))
        objs = [g]
    return dict(system=m.System(objs), doublet=dl, spot=sd)


def doublet_spot_oracle(rotate=False):
    from oracle import oracle as orc
    n1, n2 = orc.refindex(N_NLAK22_707), orc.refindex(N_NSF10_707)
    dl = orc.new("SphericalDoubletLens", AC254, [n1, n2])
    sd = orc.new("Spotdetector", [5e-3])
    th = dl.eval("thickness_object", nout=1)[0]
    sd.translate3d_([0.0, th + BFL, 0.0])
    objs = [dl, sd]
    if rotate:
        g = orc.new("ObjectGroup", ih=objs)
        g.translate3d_([0.05, 0.05, 0.05]); g.xrotate3d_(math.radians(-60)); g.zrotate3d_(math.radians(45))
        objs = [g]
    return dict(system=orc.system(objs), doublet=dl, spot=sd)


def fibonacci_disc(n, diameter=20e-3, y0=-0.05):
    """UniformDiscSource formula (src/BeamGroups.jl:232-243) with the fixed basis e1 = x, e2 = z x ... """
    k = np.arange(n, dtype=np.float64)
    R = diameter / 2
    phi0 = 2 * math.pi / (1 + math.sqrt(5))
    r = R * np.sqrt((k + 0.5) / n)
    phi = k * phi0
    pos = np.zeros((n, 3))
    pos[:, 0] = r * np.cos(phi)
    pos[:, 1] = y0
    pos[:, 2] = -(r * np.sin(phi))   # e2 = normalize(cross(dir, e1)) = cross(y, x) = -z
    d = np.zeros((n, 3)); d[:, 1] = 1.0
    return pos, d
