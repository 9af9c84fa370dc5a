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


# ---- C1: Michelson interferometer of src/Workloads/michelson_wl.jl:8-67 ---------------------------
CM = 1e-2
N_NBK7_6328 = 1.51509


def _michelson_build(F, pd_n=200, m1_shift=0.0):
    """F: factory namespace with the same method names for the product and the oracle."""
    so = [18.81 * CM, 23.5 * CM, 0.0]
    rpm = F.RightAnglePrismMirror(25e-3, 25e-3)
    rpm.zrotate3d_(math.radians(45)); rpm.translate3d_([0.0, 33.5 * CM, 0.0])
    mirror_assembly = F.ObjectGroup([rpm])
    cbs = F.CubeBeamsplitter(INCH, N_NBK7_6328)
    cbs.zrotate3d_(math.radians(-90))
    splitter_assembly = F.ObjectGroup([cbs])
    m1 = F.RoundPlanoMirror(INCH, 5e-3)
    m1.zrotate3d_(math.radians(-90)); m1.translate3d_([22 * CM, 0.0, 0.0])
    m2 = F.RoundPlanoMirror(INCH, 5e-3)
    m2.zrotate3d_(math.radians(-90)); m2.translate3d_([12 * CM, 0.0, 0.0])
    arm_1, arm_2 = F.ObjectGroup([m1]), F.ObjectGroup([m2])
    pd = F.Photodetector(8e-3, pd_n)
    pd.translate3d_([0.0, -12 * CM, 0.0])
    pd_assembly = F.ObjectGroup([pd])
    system = F.System([mirror_assembly, splitter_assembly, arm_1, arm_2, pd_assembly])
    mirror_assembly.translate_to3d_([0.0, -10 * CM, 0.0])
    splitter_assembly.translate_to3d_(so)
    arm_1.translate_to3d_(so); arm_2.translate_to3d_(so)
    arm_1.translate3d_([3.81 * CM / 2, 0.0, 0.0]); arm_2.translate3d_([0.0, 3.81 * CM / 2, 0.0])
    arm_2.zrotate3d_(math.radians(90))
    pd_assembly.translate_to3d_(so); pd_assembly.translate3d_([0.0, -3.81 * CM / 2, 0.0])
    if m1_shift:
        m1.translate3d_([m1_shift, 0.0, 0.0])
    return dict(system=system, pd=pd, m1=m1, m2=m2, cbs=cbs, rpm=rpm)


class _ProductFactory:
    def __init__(self, m): self.m = m
    def __getattr__(self, k): return getattr(self.m, k)


class _OracleFactory:
    """Same constructor names on top of oracle.new(...)."""
    def __init__(self):
        from oracle import oracle as orc
        self.orc = orc
    def RightAnglePrismMirror(self, leg, h): return self.orc.new("RightAnglePrismMirror", [leg, h])
    def CubeBeamsplitter(self, leg, n, reflectance=0.5): return self.orc.new("CubeBeamsplitter", [leg, reflectance], [self.orc.refindex(n)])
    def RoundPlanoMirror(self, d, t): return self.orc.new("RoundPlanoMirror", [d, t])
    def Photodetector(self, w, n): return self.orc.new("Photodetector", [w], [n])
    def Spotdetector(self, w): return self.orc.new("Spotdetector", [w])
    def ObjectGroup(self, objs): return self.orc.new("ObjectGroup", ih=objs)
    def System(self, objs): return self.orc.system(objs)
    def ThinBeamsplitter(self, w, h=None, reflectance=0.5): return self.orc.new("ThinBeamsplitter", [w, w if h is None else h, reflectance])
    def SquarePlanoMirror2D(self, s): return self.orc.new("SquarePlanoMirror2D", [s])
    def RectangularPlateBeamsplitter(self, w, h, t, n, reflectance=0.5):
        return self.orc.new("RectangularPlateBeamsplitter", [w, h, t, reflectance], [self.orc.refindex(n)])
    def RoundPlateBeamsplitter(self, d, t, n, reflectance=0.5):
        return self.orc.new("RoundPlateBeamsplitter", [d, t, reflectance], [self.orc.refindex(n)])
    def SphericalLens(self, r1, r2, l, d, n): return self.orc.new("SphericalLens", [r1, r2, l, d], [self.orc.refindex(n)])
    def ThinLens(self, r1, r2, d, n): return self.orc.new("ThinLens", [r1, r2, d], [self.orc.refindex(n)])
    def Retroreflector(self, s): return self.orc.new("Retroreflector", [s])
    def RightAnglePrism(self, leg, h, n): return self.orc.new("RightAnglePrism", [leg, h], [self.orc.refindex(n)])
    def ConcaveSphericalMirror(self, r, t, d): return self.orc.new("ConcaveSphericalMirror", [r, t, d])
    def RectangularPlanoMirror(self, w, h, t): return self.orc.new("RectangularPlanoMirror", [w, h, t])


def michelson(m, **kw): return _michelson_build(_ProductFactory(m), **kw)
def michelson_oracle(**kw): return _michelson_build(_OracleFactory(), **kw)
MICHELSON_BEAM = dict(pos=(0.0, 0.0, 0.0), dir=(0.0, 1.0, 0.0), lam=632.8e-9, w0=5e-4, M2=2.0, support=(1.0, 0.0, 0.0))
