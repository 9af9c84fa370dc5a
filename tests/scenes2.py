"""Scenes of BASELINE.json configs 3-5 (expander + Photodetector, mesh non-sequential with branching
and PolarizedRays, Mach-Zehnder pose sweep), each built through the product mirror and the oracle."""
import math

import numpy as np

from tests.scenes import INCH, _OracleFactory, _ProductFactory


class OracleFactory2(_OracleFactory):
    def LensFromMesh(self, mesh, n): return self.orc.new("Lens", ih=[mesh, self.orc.refindex(n)])
    def MirrorFromMesh(self, mesh): return self.orc.new("Mirror", ih=[mesh])
    def CuboidMesh(self, x, y, z, th=math.pi / 2): return self.orc.new("CuboidMesh", [x, y, z, th])
    def MeshF32(self, v, f): return self.orc.mesh(v, f, f32=True)
    def LoadSTL(self, path): return self.orc.load_stl(path)


class ProductFactory2(_ProductFactory):
    def LensFromMesh(self, mesh, n): return self.m.Lens(mesh, n)
    def MirrorFromMesh(self, mesh): return self.m.Mirror(mesh)
    def MeshF32(self, v, f): return self.m.Mesh(v, f, scale=float(np.float32(1e-3)), f32=True)
    def LoadSTL(self, path): return self.m.load_stl(path)


# ---- C3: Keplerian expander (docs/src/tutorials/expander.md:25-66) + Photodetector ---------------------
def _expander(F, pd_n, pd_w=40e-3):
    R11, R12, R21, R22, nl = 20e-3, 60e-3, 60e-3, 90e-3, 1.5
    f1 = 1 / ((nl - 1) * (1 / R11 + 1 / R12))
    f2 = 1 / ((nl - 1) * (1 / R21 + 1 / R22))
    l1 = F.SphericalLens(R11, R12, 0, INCH, nl)
    l2 = F.SphericalLens(R21, R22, 0, 2 * INCH, nl)
    l1.zrotate3d_(math.radians(180)); l2.zrotate3d_(math.radians(180))
    l2.translate3d_([0, f1 + f2, 0])
    pd = F.Photodetector(pd_w, pd_n)
    pd.translate3d_([0, 0.25, 0])
    return dict(system=F.System([l1, l2, pd]), pd=pd, l1=l1, l2=l2)


def expander(m, pd_n, **kw): return _expander(ProductFactory2(m), pd_n, **kw)
def expander_oracle(pd_n, **kw): return _expander(OracleFactory2(), pd_n, **kw)


def beamlet_lattice(k, aperture=8e-3, y0=-0.05):
    """k x k beamlets across `aperture`: lambda = 1 um, w0 = 1.5 pitch, P0 = 1 mW / k^2, support = x."""
    pitch = aperture / k
    c = (np.arange(k) - (k - 1) / 2) * pitch
    X, Z = np.meshgrid(c, c, indexing="ij")
    pos = np.stack([X.ravel(), np.full(k * k, y0), Z.ravel()], axis=1)
    return dict(pos=pos, dir=(0.0, 1.0, 0.0), lam=1e-6, w0=1.5 * pitch, M2=1.0, P0=1e-3 / (k * k), support=(1.0, 0.0, 0.0))


# ---- C4: non-sequential mesh scene with beamsplitter branching and PolarizedRays ---------------------
def uv_sphere_f32(radius_mm, nu, nv):
    """A closed triangle mesh in STL style (3 private vertices per face, Float32 millimetres): stands in for
    the reference's binary STL assets (docs/src/assets/*.stl are not shipped to the GPU box)."""
    th = np.linspace(0, math.pi, nv + 1)
    ph = np.linspace(0, 2 * math.pi, nu + 1)
    P = lambda i, j: radius_mm * np.array([math.sin(th[j]) * math.cos(ph[i]), math.cos(th[j]), math.sin(th[j]) * math.sin(ph[i])])
    tris = []
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
            if j > 0: tris.append([a, c, b])
            if j < nv - 1: tris.append([a, d, c])
    v = np.array(tris, dtype=np.float32).reshape(-1, 3) * np.float32(1e-3)   # Mesh.jl:63-65: scaled in Float32
    f = np.arange(v.shape[0], dtype=np.int32).reshape(-1, 3)
    return v.astype(np.float64), f


def _mesh_scene(F, nu=96, nv=96):
    # Fresnel rhomb (test/runtests.jl:2339-2349), scaled to 1/10
    rhomb = F.LensFromMesh(F.CuboidMesh(0.05, 0.125, 0.05, math.radians(53.3)), 1.5)
    rhomb.translate3d_([-0.025, -0.5, -0.025])
    bs = F.ThinBeamsplitter(0.5, reflectance=0.5)
    bs.zrotate3d_(math.radians(45))
    bs.translate3d_([0.075, 0.0, 0.0])
    retro = F.Retroreflector(0.25)
    # turn the open corner (1,1,1)/sqrt(3) towards -y and park it behind the splitter
    a = np.array([1.0, 1.0, 1.0]) / math.sqrt(3)
    b = np.array([0.0, -1.0, 0.0])
    ax = np.cross(a, b); ax /= np.linalg.norm(ax)
    retro.rotate3d_([float(x) for x in ax], math.acos(float(a @ b)))
    retro.translate3d_([0.075, 0.45, 0.0])
    v, f = uv_sphere_f32(50.0, nu, nv)
    ball = F.MirrorFromMesh(F.MeshF32(v, f))
    ball.translate3d_([0.4, 0.0, 0.0])
    return dict(system=F.System([rhomb, bs, retro, ball]), rhomb=rhomb, bs=bs, retro=retro, ball=ball)


def mesh_scene(m, **kw): return _mesh_scene(ProductFactory2(m), **kw)
def mesh_scene_oracle(**kw): return _mesh_scene(OracleFactory2(), **kw)


# BASELINE config 4 as SURVEY 8(d) specifies it: the Fresnel rhomb of test/runtests.jl:2339-2349 at full size (CuboidMesh
# 0.5 x 1.25 x 0.5 m, 53.3 deg, n = 1.5, turned by 135 deg about y), a ThinBeamsplitter (R = 0.5), Retroreflector(25e-3)
# (Misc.jl:49) and the reference's own STL asset as a mirror: Mirror(Mesh(load("Mirror_Post.stl"))), 18 196 triangles.
# A ray entering at the origin along +y leaves the rhomb at x = z = RHOMB_EXIT, y = 1.25, still along +y.
RHOMB_EXIT = -0.52102107
C4_HALF, C4_Y0 = 0.006, -1.0      # the bundle: jittered lattice of half-width 6 mm (inside the retroreflector's aperture)


def _mesh_scene_c4(F, stl_path):
    s1 = F.CuboidMesh(0.5, 1.25, 0.5, math.radians(53.3))
    rhomb = F.LensFromMesh(s1, 1.5)
    rhomb.translate3d_([-0.25, 0.0, -0.25])
    s1.set_new_origin3d_()
    rhomb.yrotate3d_(math.radians(135))
    c = RHOMB_EXIT
    bs = F.ThinBeamsplitter(0.1, reflectance=0.5)
    bs.zrotate3d_(math.radians(45))
    bs.translate3d_([c, 1.5, c])
    retro = F.Retroreflector(25e-3)
    a = np.array([1.0, 1.0, 1.0]) / math.sqrt(3)
    b = np.array([0.0, -1.0, 0.0])
    ax = np.cross(a, b); ax /= np.linalg.norm(ax)
    retro.rotate3d_([float(x) for x in ax], math.acos(float(a @ b)))      # open corner towards -y
    retro.translate3d_([c, 1.8 + 25e-3 / math.sqrt(3), c])                 # apex on the beam axis, aperture plane at y = 1.8
    post = F.MirrorFromMesh(F.LoadSTL(stl_path))
    post.zrotate3d_(math.radians(20))
    post.translate3d_([c + 0.3, 1.49, c - 0.03])                            # in the reflected arm (+x from the splitter)
    return dict(system=F.System([rhomb, bs, retro, post]), rhomb=rhomb, bs=bs, retro=retro, post=post)


def mesh_scene_c4(m, stl_path): return _mesh_scene_c4(ProductFactory2(m), stl_path)
def mesh_scene_c4_oracle(stl_path): return _mesh_scene_c4(OracleFactory2(), stl_path)


def jittered_lattice(n, seed=0, half=0.012, y0=-0.6):
    k = int(math.ceil(math.sqrt(n)))
    rng = np.random.default_rng(seed)
    c = (np.arange(k) + 0.5) / k * 2 * half - half
    X, Z = np.meshgrid(c, c, indexing="ij")
    pos = np.stack([X.ravel(), np.full(k * k, y0), Z.ravel()], axis=1)[:n]
    pos[:, [0, 2]] += (rng.random((n, 2)) - 0.5) * (2 * half / k)
    return pos, np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0], dtype=np.complex128)   # |E0| = 1 keeps the 1e-14 orthogonality check away from rounding noise


# ---- C5: Mach-Zehnder (test/runtests.jl:2366-2382) with a GaussianBeamlet and a Photodetector ----------
def _mzi(F, pd_n=64):
    m1, m2 = F.SquarePlanoMirror2D(INCH), F.SquarePlanoMirror2D(INCH)
    b1, b2 = F.ThinBeamsplitter(INCH, reflectance=0.5), F.ThinBeamsplitter(INCH, reflectance=0.5)
    pd = F.Photodetector(5e-3, pd_n)
    system = F.System([m1, m2, b1, b2, pd])
    b2.translate3d_([2 * INCH, 2 * INCH, 0]); m1.translate3d_([0, 2 * INCH, 0]); m2.translate3d_([2 * INCH, 0, 0])
    b1.zrotate3d_(math.radians(360 - 135)); b2.zrotate3d_(math.radians(45))
    m1.zrotate3d_(math.radians(360 - 135)); m2.zrotate3d_(math.radians(45))
    pd.translate3d_([2 * INCH, 3 * INCH, 0])
    return dict(system=system, m1=m1, m2=m2, b1=b1, b2=b2, pd=pd)


def mzi(m, **kw): return _mzi(ProductFactory2(m), **kw)
def mzi_oracle(**kw): return _mzi(OracleFactory2(), **kw)
MZI_BEAM = dict(pos=(0.0, -0.1, 0.0), dir=(0.0, 1.0, 0.0), lam=1e-6, w0=5e-4, M2=1.0, P0=1e-3, support=(1.0, 0.0, 0.0))
MZI_M1_NORMAL = (-math.sin(math.radians(45)), math.sin(math.radians(45)), 0.0)


def mzi_shift(p, n_poses, lam=1e-6):
    """Displacement of m1 along its normal for pose p: uniform over [0, 2 lambda]."""
    s = 2 * lam * p / max(n_poses - 1, 1)
    return [MZI_M1_NORMAL[0] * s, MZI_M1_NORMAL[1] * s, 0.0]
