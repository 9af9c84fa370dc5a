"""Ray / beam data model of the reference (layer L2) on the host.

Single beams mirror the reference's objects (`Ray`, `Beam`, `GaussianBeamlet`); bundles
(`RayBundle`, `BeamletBundle`) are the structure-of-arrays form the GPU path consumes -- a
`CollimatedSource` of a million rays is one bundle, not a million Python objects.
Citations: /root/reference/src.
"""
import math

import numpy as np

from . import linalg as la

Z_VACUUM = 376.730313668   # Constants.jl:6


class Intersection:
    """AbstractTypes/AbstractRay.jl:13-18"""
    __slots__ = ("object", "shape", "t", "n")

    def __init__(self, t, n, obj=None, shape=None):
        self.t, self.n, self.object, self.shape = t, n, obj, shape


class Ray:
    """Rays.jl:14-42 (dir is normalised, n = 1)."""
    polarized = False

    def __init__(self, pos, dir, lam=1000e-9, n=1.0, normalize=True):
        self.pos = la.v3(pos)
        self.dir = la.normalize(la.v3(dir)) if normalize else la.v3(dir)
        self.intersection = None
        self.lam = float(lam)
        self.n = float(n)

    def length(self): return math.inf if self.intersection is None else self.intersection.t
    def optical_path_length(self): return math.inf if self.intersection is None else self.intersection.t * self.n


class PolarizedRay(Ray):
    """PolarizedRays.jl:37-95: E0 must be orthogonal to dir (atol 1e-14)."""
    polarized = True

    def __init__(self, pos, dir, lam=1000e-9, E0=None, n=1.0, normalize=True):
        super().__init__(pos, dir, lam, n, normalize)
        if E0 is None:
            E0 = (math.sqrt(2 * 1 * Z_VACUUM), 0.0, 0.0)   # [electric_field(1), 0, 0]
        self.E0 = tuple(complex(x) for x in E0)
        d = sum(self.dir[k] * self.E0[k] for k in range(3))
        if abs(d) > 1e-14:
            raise ValueError("Ray dir. and E0 must be orthogonal.")


class Beam:
    """Beam.jl:13-17"""

    def __init__(self, ray_or_pos, dir=None, lam=None, E0=None):
        if isinstance(ray_or_pos, Ray):
            ray = ray_or_pos
        elif E0 is not None:
            ray = PolarizedRay(ray_or_pos, dir, 1000e-9 if lam is None else lam, E0)
        else:
            ray = Ray(ray_or_pos, dir, 1000e-9 if lam is None else lam)
        self.rays = [ray]
        self.parent = None
        self.children = []

    def length(self):   # Beam.jl:125-169
        l = 0.0
        for r in self.rays:
            if r.intersection is None:
                break
            l += r.length()
        return l + (self.parent.length() if self.parent is not None else 0.0)

    def optical_path_length(self):   # :137-149
        l0 = self.parent.optical_path_length() if self.parent is not None else 0.0
        for r in self.rays:
            if r.intersection is None:
                break
            l0 += r.optical_path_length()
        return l0

    def leaves(self):
        out, q = [], [self]
        while q:
            b = q.pop(0)
            if not b.children:
                out.append(b)
            q.extend(b.children)
        return out


class GaussianBeamlet:
    """Gaussian.jl:33-42 / 215-256.  `support` must be given for reproducible results (the
    reference draws a random orthogonal vector when it is omitted)."""

    def __init__(self, position, direction, lam=1e-6, w0=1e-3, M2=1.0, P0=1e-3, z0=0.0, support=None):
        d = la.normalize(la.v3(direction))
        if support is None:
            rng = np.random.default_rng()
            nw = tuple(rng.random(3))
            nn = la.norm(d)
            nw = la.sub(nw, la.scale(1.0 / (nn * nn), la.scale(la.dot(nw, d), d)))
            support = nw
        s1 = la.normalize(la.v3(support))
        tant = math.tan(M2 * lam / (math.pi * w0))
        pos = la.v3(position)
        self.chief = Beam(Ray(pos, d, lam))
        self.waist = Beam(Ray(la.add(pos, la.scale(w0, s1)), d, lam))
        dz = -z0 * tant
        self.divergence = Beam(Ray(la.add(pos, la.scale(dz, s1)), la.normalize(la.add(d, la.scale(tant, s1))), lam))  # normalised twice, like the reference
        self.lam = float(lam)
        self.w0 = float(w0)
        I0 = 2 * P0 / (math.pi * (w0 * w0))
        self.E0 = complex(math.sqrt(2 * I0 * Z_VACUUM), 0.0)
        self.parent = None
        self.children = []

    @classmethod
    def _raw(cls, chief, waist, div, lam, w0, E0):
        g = cls.__new__(cls)
        g.chief, g.waist, g.divergence, g.lam, g.w0, g.E0 = chief, waist, div, lam, w0, E0
        g.parent, g.children = None, []
        return g

    def length(self): return self.chief.length()
    def optical_path_length(self): return self.chief.optical_path_length()

    def rays18(self):
        out = []
        for b in (self.chief, self.waist, self.divergence):
            r = b.rays[0]
            out.extend(r.pos + r.dir)
        return out

    def leaves(self):
        out, q = [], [self]
        while q:
            b = q.pop(0)
            if not b.children:
                out.append(b)
            q.extend(b.children)
        return out


# ---- bundles (SoA) -----------------------------------------------------------------------------
class RayBundle:
    """N independent root rays: pos/dir (N,3), lam (N,), optional E0 (N,3) complex -> PolarizedRays."""

    def __init__(self, pos, dir, lam=1000e-9, E0=None, normalize=True):
        self.pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dir, dtype=np.float64)
        self.uniform_dir = d.ndim == 1      # one direction for the whole bundle (collimated sources): shipped once (BMO_UNIFORM_DIR)
        if d.ndim == 1:
            d = np.broadcast_to(d, self.pos.shape)
        d = np.array(d, dtype=np.float64)
        if normalize:   # inv(norm) * d like Ray(pos, dir, lam)
            inv = 1.0 / np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
            d = d * inv[:, None]
        self.dir = np.ascontiguousarray(d)
        self.uniform_lam = np.ndim(lam) == 0
        self.lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lam, dtype=np.float64), (self.pos.shape[0],)))
        self.E0 = None
        if E0 is not None:
            e = np.asarray(E0, dtype=np.complex128)
            if e.ndim == 1:
                e = np.broadcast_to(e, self.pos.shape)
            self.E0 = np.ascontiguousarray(e)
        self.result = None

    def __len__(self): return self.pos.shape[0]


class BeamletBundle:
    """N root Gaussian beamlets in SoA form: rays (N,3,6) = chief/waist/divergence pos+dir."""

    def __init__(self, rays, lam, w0, E0):
        self.rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3, 6)
        n = self.rays.shape[0]
        self.lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lam, dtype=np.float64), (n,)))
        self.w0 = np.ascontiguousarray(np.broadcast_to(np.asarray(w0, dtype=np.float64), (n,)))
        self.E0 = np.ascontiguousarray(np.broadcast_to(np.asarray(E0, dtype=np.complex128), (n,)))
        self.result = None

    def __len__(self): return self.rays.shape[0]

    @classmethod
    def from_params(cls, position, direction, lam=1e-6, w0=1e-3, M2=1.0, P0=1e-3, z0=0.0, support=(1.0, 0.0, 0.0)):
        """Vectorised GaussianBeamlet constructor (Gaussian.jl:215-256) for (N,3) positions."""
        pos = np.ascontiguousarray(position, dtype=np.float64).reshape(-1, 3)
        n = pos.shape[0]
        out = np.zeros((n, 3, 6))
        w0a = np.broadcast_to(np.asarray(w0, dtype=np.float64), (n,))
        lama = np.broadcast_to(np.asarray(lam, dtype=np.float64), (n,))
        P0a = np.broadcast_to(np.asarray(P0, dtype=np.float64), (n,))
        d = la.normalize(la.v3(direction))
        s1 = la.normalize(la.v3(support))
        dv, sv = np.array(la.normalize(d)), np.array(s1)   # Ray(pos, dir, lam) normalises `dir` a second time
        for i in range(n):
            tant = math.tan(M2 * lama[i] / (math.pi * w0a[i]))
            out[i, 0, :3] = pos[i]; out[i, 0, 3:] = dv
            out[i, 1, :3] = pos[i] + sv * w0a[i]; out[i, 1, 3:] = dv
            dd = la.normalize(la.normalize(la.add(d, la.scale(tant, s1))))
            out[i, 2, :3] = pos[i] + sv * (-z0 * tant); out[i, 2, 3:] = dd
        I0 = 2 * P0a / (math.pi * (w0a * w0a))
        E0 = np.sqrt(2 * I0 * Z_VACUUM).astype(np.complex128)
        return cls(out, lama, w0a, E0)


# ---- sources (BeamGroups.jl) -------------------------------------------------------------------
def _basis(dir, b1):
    d = la.normalize(la.v3(dir))
    if b1 is None:   # the reference's normal3d(dir): random Gram-Schmidt
        rng = np.random.default_rng()
        nw = tuple(rng.random(3))
        nn = la.norm(d)
        b1 = la.normalize(la.sub(nw, la.scale(1.0 / (nn * nn), la.scale(la.dot(nw, d), d))))
    return d, la.v3(b1)


def UniformDiscSource(pos, dir, diameter, lam=1e-6, num_rays=1000, e1=None):
    """BeamGroups.jl:222-245 (Fibonacci disc); `e1` fixes the reference's random in-plane basis."""
    d, e1 = _basis(dir, e1)
    e2 = la.normalize(la.cross(d, e1))
    R = diameter / 2
    phi0 = 2 * math.pi / (1 + math.sqrt(5))
    k = np.arange(num_rays, dtype=np.float64)
    rho = np.sqrt((k + 0.5) / num_rays)
    phi = k * phi0
    r = R * rho
    x = (r * np.cos(phi))[:, None] * np.array(e1) + (r * np.sin(phi))[:, None] * np.array(e2)
    p = np.array(la.v3(pos)) + x
    b = RayBundle(p, np.array(d), lam)
    b.diameter = float(diameter)
    return b


def CollimatedSource(pos, dir, diameter, lam=1e-6, num_rings=10, num_rays=None, b1=None):
    """BeamGroups.jl:152-195 (concentric rings)."""
    if num_rays is None:
        num_rays = 100 * num_rings
    if num_rays < num_rings * 20:
        raise ValueError("No. of rays should be atleast 20x no. of rings")
    d, b1 = _basis(dir, b1)
    p0 = la.v3(pos)
    pts = [p0]
    num_rays -= 1
    radii = [(diameter / 2) * i / (num_rings - 1) for i in range(1, num_rings)]
    circm = [r * 2 * math.pi for r in radii]
    ds = sum(circm) / num_rays
    n_rays = [int(round(c / ds)) for c in circm]
    n_rays[-1] += num_rays - sum(n_rays)
    for r, numEl in zip(radii, n_rays):
        if numEl == 0:
            continue
        Rm = la.rotate3d(d, 2 * math.pi / numEl)
        helper = la.scale(r, b1)
        for _ in range(numEl):
            pts.append(la.add(p0, helper))
            helper = la.matvec(Rm, helper)
    b = RayBundle(np.array(pts), np.array(d), lam)
    b.diameter = float(diameter)
    return b


def PointSource(pos, dir, theta, lam=1e-6, num_rings=10, num_rays=None, b1=None):
    """BeamGroups.jl:51-100 (cone of concentric fans)."""
    if num_rays is None:
        num_rays = 100 * num_rings
    if num_rays < num_rings * 20:
        raise ValueError("No. of rays should be atleast 20x no. of rings")
    if theta >= math.pi:
        raise ValueError("Point source opening half-angle must be <= pi")
    d, b1 = _basis(dir, b1)
    b2 = la.normalize(la.cross(d, b1))
    step = theta / (num_rings - 1)
    dirs = [d]
    num_rays -= 1
    ndirs = [la.matvec(la.rotate3d(b2, step * i), d) for i in range(1, num_rings)]
    circm = [la.norm(la.sub(nd, la.scale(la.dot(nd, d), d))) * 2 * math.pi for nd in ndirs]
    ds = sum(circm) / num_rays
    n_rays = [int(round(c / ds)) for c in circm]
    n_rays[-1] += num_rays - sum(n_rays)
    for nd, numEl in zip(ndirs, n_rays):
        if numEl == 0:
            continue
        Rm = la.rotate3d(d, 2 * math.pi / numEl)
        cdir = nd
        for _ in range(numEl):
            dirs.append(cdir)
            cdir = la.matvec(Rm, cdir)
    n = len(dirs)
    b = RayBundle(np.tile(np.array(la.v3(pos)), (n, 1)), np.array(dirs), lam)
    b.NA = math.sin(theta)
    return b
