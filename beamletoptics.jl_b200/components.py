"""Host-side optical components (reference layer L5) and containers: the constructors and the
kinematic API of the reference, nothing that runs per ray.  Citations: /root/reference/src.
"""
import math

import numpy as np

from . import linalg as la
from . import shapes as sh

inch = 25.4e-3   # Constants.jl:8


# ---- refractive index (Utils/RefractiveIndexUtils.jl) ---------------------------------------------
class DiscreteRefractiveIndex:
    """:8-31 -- exact-wavelength lookup, KeyError on a miss."""

    def __init__(self, lambdas, ns):
        if len(lambdas) != len(ns):
            raise ValueError("Number of wavelengths must match number of ref. indices")
        self.data = {float(l): float(n) for l, n in zip(lambdas, ns)}

    def __call__(self, lam):
        return self.data[float(lam)]


class SellmeierEquation:
    """:82-98 -- six-coefficient Sellmeier equation, lambda in metres."""

    def __init__(self, B1, B2, B3, C1, C2, C3):
        self.B = (float(B1), float(B2), float(B3))
        self.C = (float(C1), float(C2), float(C3))

    def __call__(self, lam):
        x = lam * 1e6
        B, C = self.B, self.C
        n2 = 1 + (B[0] * (x * x)) / (x * x - C[0]) + (B[1] * (x * x)) / (x * x - C[1]) + (B[2] * (x * x)) / (x * x - C[2])
        return math.sqrt(n2)


def _as_index(n):
    if callable(n):
        return n
    val = float(n)
    return lambda lam: val


# ---- objects -----------------------------------------------------------------------------------
class AbstractObject:
    """AbstractTypes/AbstractObject.jl + AbstractShapeTrait.jl: kinematics dispatch on the shape trait."""
    multi = False

    # SingleShape (AbstractShapeTrait.jl:34-51)
    def position(self): return self.shape.position()
    def orientation(self): return self.shape.orientation()
    def translate3d_(self, offset): self.shape.translate3d_(offset)
    def translate_to3d_(self, target): self.shape.translate_to3d_(target)
    def rotate3d_(self, axis, theta): self.shape.rotate3d_(axis, theta)
    def xrotate3d_(self, theta): self.rotate3d_((1.0, 0.0, 0.0), theta)
    def yrotate3d_(self, theta): self.rotate3d_((0.0, 1.0, 0.0), theta)
    def zrotate3d_(self, theta): self.rotate3d_((0.0, 0.0, 1.0), theta)
    def align3d_(self, axis): self.shape.align3d_(axis)
    def reset_translation3d_(self): self.shape.reset_translation3d_()
    def reset_rotation3d_(self): self.shape.reset_rotation3d_()
    def set_new_origin3d_(self): self.shape.set_new_origin3d_()
    def thickness(self): return self.shape.thickness()


class MultiShapeObject(AbstractObject):
    """MultiShape trait (AbstractShapeTrait.jl:53-160): `parts` is `shape(object)`."""
    multi = True

    def position(self): return self.parts[0].position()
    def orientation(self): return self.parts[0].orientation()
    def _set_position(self, p): pass
    def _set_orientation(self, d): pass

    def translate3d_(self, offset):     # :88-95
        offset = la.v3(offset)
        self._set_position(la.add(self.position(), offset))
        for p in self.parts:
            p.translate3d_(offset)

    def translate_to3d_(self, target):  # :103-107
        self.translate3d_(la.sub(la.v3(target), self.position()))

    def rotate3d_(self, axis, theta):   # :115-128
        R = la.rotate3d(axis, theta)
        self._set_orientation(la.matmul(R, self.orientation()))
        for p in self.parts:
            p.rotate3d_(axis, theta)
            v = la.sub(p.position(), self.position())
            v = la.sub(la.matvec(R, v), v)
            p.translate3d_(v)

    def align3d_(self, axis): pass      # :130-134 not implemented in the reference

    def reset_translation3d_(self):     # :144-150
        self.translate3d_(la.neg(self.position()))
        self._set_position((0.0, 0.0, 0.0))

    def reset_rotation3d_(self):        # :160-173
        axis, th = la.rotation_axis_angle(self.orientation())
        if axis is None:
            return
        self.rotate3d_(axis, -th)
        self._set_orientation(la.IDENTITY)


class ObjectGroup(MultiShapeObject):
    """ObjectGroups.jl:20-47"""

    def __init__(self, objects):
        self.parts = list(objects)
        self.center = (0.0, 0.0, 0.0)
        self.dir = la.IDENTITY

    @property
    def objects(self): return self.parts
    def position(self): return self.center
    def orientation(self): return self.dir
    def _set_position(self, p): self.center = p
    def _set_orientation(self, d): self.dir = d


class Lens(AbstractObject):
    """Lenses.jl:146-155 (and Prism, Prisms.jl:7-14: same behaviour)."""
    kind = "refractive"

    def __init__(self, shape, n):
        self.shape = shape
        self.n = _as_index(n)

    def refractive_index(self, lam): return float(self.n(lam))


class Prism(Lens):
    pass


class Mirror(AbstractObject):
    """Mirrors.jl:78-80 and the concrete mirror types (all surfaces reflect)."""
    kind = "mirror"

    def __init__(self, shape):
        self.shape = shape


class IntersectableObject(AbstractObject):   # Intersectable.jl:10-15
    kind = "stop"

    def __init__(self, shape): self.shape = shape


class NonInteractableObject(AbstractObject):  # NonInteractable.jl:14-22 (invisible to the tracer)
    kind = "nonint"

    def __init__(self, shape): self.shape = shape


def MeshDummy(path):
    return NonInteractableObject(sh.load_stl(path))


class ThinBeamsplitter(AbstractObject):
    """ThinBeamsplitter.jl:16-51: amplitudes sqrt(R), sqrt(1 - R^2-as-amplitude)."""
    kind = "thin_bs"

    def __init__(self, width, height=None, reflectance=0.5, shape=None):
        if reflectance >= 1 or la.isapprox(reflectance, 0.0):
            raise ValueError("Splitting ratio in (0, 1)!")
        if shape is None:
            shape = sh.RectangularFlatMesh(width, width if height is None else height)
        self.shape = shape
        self.reflectance = math.sqrt(reflectance)
        self.transmittance = math.sqrt(1 - self.reflectance * self.reflectance)


def RoundThinBeamsplitter(diameter, reflectance=0.5):   # :59-67
    return ThinBeamsplitter(None, reflectance=reflectance, shape=sh.CircularFlatMesh(diameter / 2))


class DoubletLens(MultiShapeObject):
    """DoubletLenses.jl:26-43"""
    kind = "doublet"

    def __init__(self, front, back):
        self.front, self.back = front, back
        self.parts = [front, back]

    def thickness(self): return self.front.shape.thickness() + self.back.shape.thickness()


class _PlateBeamsplitter(MultiShapeObject):
    """PlateBeamsplitter.jl:33-46: parts (substrate, coating); kinematic centre = coating position,
    orientation = substrate orientation."""
    kind = "plate_bs"

    def __init__(self, substrate, coating):
        self.substrate, self.coating = substrate, coating
        self.parts = [substrate, coating]

    def position(self): return self.coating.position()
    def orientation(self): return self.substrate.orientation()
    def refractive_index(self, lam): return self.substrate.refractive_index(lam)


def RectangularPlateBeamsplitter(width, height, thickness, n, reflectance=0.5):   # :89-104
    substrate = Prism(sh.BoxSDF(width, thickness, height), n)
    substrate.translate3d_((0.0, thickness / 2, 0.0))
    coating = ThinBeamsplitter(width, height, reflectance=reflectance)
    coating.zrotate3d_(math.pi)
    return _PlateBeamsplitter(substrate, coating)


def RoundPlateBeamsplitter(diameter, thickness, n, reflectance=0.5):              # :146-158
    substrate = Prism(sh.PlanoSurfaceSDF(thickness, diameter), n)
    coating = RoundThinBeamsplitter(diameter, reflectance=reflectance)
    return _PlateBeamsplitter(substrate, coating)


class CubeBeamsplitter(MultiShapeObject):
    """CubeBeamsplitter.jl:22-61: parts (front, back, coating)."""
    kind = "cube_bs"

    def __init__(self, leg_length, n, reflectance=0.5):
        self.front = RightAnglePrism(leg_length, leg_length, n)
        self.back = RightAnglePrism(leg_length, leg_length, n)
        self.coating = ThinBeamsplitter(math.sqrt(2.0) * leg_length, leg_length, reflectance=reflectance)
        self.back.zrotate3d_(la.deg2rad(180))
        self.coating.zrotate3d_(la.deg2rad(180 - 45))
        self.coating.shape.set_new_origin3d_()
        self.parts = [self.front, self.back, self.coating]

    def refractive_index(self, lam): return self.front.refractive_index(lam)


class Photodetector(AbstractObject):
    """Photodetector.jl:31-55.  `field` is filled by solve_system_ (GPU kernel K4)."""
    kind = "pd"

    def __init__(self, width, n):
        self.shape = sh.QuadraticFlatMesh(width)
        sz = float(self.shape.vertices.max())
        self.lo, self.hi, self.n = -sz, sz, int(n)
        self.field = np.zeros((n, n), dtype=np.complex128, order="F")   # column-major like Matrix{ComplexF64}

    @property
    def x(self): return np.array([self._coord(i) for i in range(self.n)])
    y = x

    def _coord(self, i):
        t = 0.0 if self.n == 1 else i / (self.n - 1)
        return (1 - t) * self.lo + t * self.hi

    def empty_(self): self.field[...] = 0

    def intensity(self):   # OpticUtils.jl:108
        return (self.field.real ** 2 + self.field.imag ** 2) / (2 * 376.730313668)

    def optical_power(self):  # Photodetector.jl:116 trapz((x, y), intensity)
        I = self.intensity()
        x = self.x
        dx = x[1:] - x[:-1]
        col = (dx[:, None] * (I[:-1, :] + I[1:, :]) / 2).sum(axis=0)
        return float((dx * (col[:-1] + col[1:]) / 2).sum())


class Spotdetector(AbstractObject):
    """Spotdetector.jl:20-61"""
    kind = "spot"

    def __init__(self, width):
        self.shape = sh.QuadraticFlatMesh(width)
        self.shape.zrotate3d_(math.pi)
        self.data = np.zeros((0, 2))
        self.hw = width / 2

    def empty_(self): self.data = np.zeros((0, 2))


def XYBasis(j11, j12, j21, j22): return ((j11, j12, 0.0), (j21, j22, 0.0), (0.0, 0.0, 1.0))     # PolarizedRays.jl:148
def XZBasis(j11, j12, j21, j22): return ((j11, 0.0, j12), (0.0, 1.0, 0.0), (j21, 0.0, j22))     # :149
def YZBasis(j11, j12, j21, j22): return ((1.0, 0.0, 0.0), (0.0, j22, j21), (0.0, j12, j11))     # :150


class PolarizationFilter(AbstractObject):
    """Polarizers/PolarizationFilter.jl:5-29: zero-thickness ideal polariser, aligned with the y-axis, transmitting
    along its local x-axis and blocking z.  `JMat` is the GlobalJonesBasis (3x3, real) of the unrotated element."""
    kind = "polfilter"

    def __init__(self, edge_length, cutoff_strength=2.220446049250313e-16, JMat=None):
        self.shape = sh.QuadraticFlatMesh(edge_length)
        self.shape.zrotate3d_(math.pi)
        self.shape.set_new_origin3d_()
        self.JMat = tuple(tuple(float(x) for x in row) for row in (JMat if JMat is not None else XZBasis(1, 0, 0, 0)))
        self.cutoff = float(cutoff_strength)


class PSFDetector(AbstractObject):
    """Detectors/PSFDetector.jl:44-68.  The hit records (`PSFData`: hit, dir, opl, proj, k) stay on the device
    (bmo_psf); `data` downloads them, `intensity` runs the coherent sum kernel (PSFDetector.jl:190-237)."""
    kind = "psf"

    def __init__(self, width):
        self.shape = sh.QuadraticFlatMesh(width)
        self.shape.zrotate3d_(math.pi)
        self._psf = None        # bmo_psf handle (ctypes void*), filled by solve_system_
        self._dsys = None       # the DeviceSystem of the last solve (detector pose for intensity / lims)
        self._index = -1

    def __len__(self):
        from . import solver
        return solver.psf_count(self)

    @property
    def data(self):
        """(n, 9) array: hit xyz, dir xyz, optical path length, projection factor, wavenumber."""
        from . import solver
        return solver.psf_data(self)

    def empty_(self):
        from . import solver
        solver.psf_free(self)

    def calc_local_lims(self, crop_factor=1.0, center="centroid"):
        from . import solver
        return solver.psf_lims(self, crop_factor, center)

    def intensity(self, n=100, crop_factor=1.0, center="centroid", x_min=math.inf, x_max=math.inf, z_min=math.inf, z_max=math.inf,
                  x0_shift=0.0, z0_shift=0.0):
        """-> (xs, zs, I) like intensity(psf; ...) (PSFDetector.jl:190-237); I is (n, n), raw / unscaled."""
        from . import solver
        return solver.psf_intensity(self, n, crop_factor, center, x_min, x_max, z_min, z_max, x0_shift, z0_shift)

    def __del__(self):
        try:
            self.empty_()
        except Exception:
            pass


# ---- constructors ------------------------------------------------------------------------------
def _surf_forward(r, d):    # SphericalLensSDF.jl:423-437
    if math.isinf(r):
        return None
    return sh.ConvexSphericalSurfaceSDF(r, d) if r > 0 else sh.ConcaveSphericalSurfaceSDF(abs(r), d)


def _surf_backward(r, d):   # :439-448
    if math.isinf(r):
        return None
    b = sh.ConcaveSphericalSurfaceSDF(r, d) if r > 0 else sh.ConvexSphericalSurfaceSDF(abs(r), d)
    b.zrotate3d_(math.pi)
    return b


def _sign(x): return (x > 0) - (x < 0)


def _meniscus(r1, d1, front, r2, d2, back, ct):   # MeniscusLensSDF.jl:122-189
    if _sign(r1) == _sign(r2) and _sign(r2) > 0: left = True
    elif _sign(r1) == _sign(r2) and _sign(r2) < 0: left = False
    else: raise ValueError("Invalid sign combination for r1 and r2")
    convex_sag, concave_sag = front.sag(), back.sag()
    cylinder_l = ct - convex_sag + concave_sag
    if cylinder_l <= 0:
        raise ValueError("Lens parameters lead to zero lens edge thickness")
    if left: f, b = sh.ConvexSphericalSurfaceSDF(r1, d1), sh.SphereSDF(r2)
    else: f, b = sh.SphereSDF(abs(r1)), sh.ConvexSphericalSurfaceSDF(abs(r2), d2)
    cyl = sh.PlanoSurfaceSDF(cylinder_l, min(d1, d2))
    if left:
        cyl.translate3d_((0.0, f.thickness(), 0.0))
        b.translate3d_((0.0, r2 + ct, 0.0))
        cvx, ccv = f, b
    else:
        b.translate3d_((0.0, -abs(r1), 0.0))
        cyl.translate3d_((0.0, -concave_sag, 0.0))
        f.zrotate3d_(math.pi)
        f.translate3d_((0.0, cyl.thickness() - concave_sag + convex_sag, 0.0))
        cvx, ccv = b, f
    return sh.MeniscusLensSDF(cvx, cyl, ccv, ct)


class SphericalSurface:
    """SphericalLensSDF.jl:380-419: radius of curvature, clear aperture, mechanical diameter (r = Inf: flat)."""
    def __init__(self, radius, diameter, mechanical_diameter=None):
        self.radius, self.diameter = float(radius), float(diameter)
        self.mechanical_diameter = float(diameter if mechanical_diameter is None else mechanical_diameter)

    def sdf_forward(self): return _surf_forward(self.radius, self.diameter)
    def sdf_backward(self): return _surf_backward(self.radius, self.diameter)


def CircularFlatSurface(diameter):    # SphericalLensSDF.jl:456-466
    return SphericalSurface(math.inf, diameter)


class EvenAsphericalSurface(SphericalSurface):
    """AsphericalLensSDF.jl:392-428: base sphere + conic constant + even coefficients (coefficients[i] multiplies r^(2i))."""
    def __init__(self, radius, diameter, conic_constant, coefficients, mechanical_diameter=None):
        super().__init__(radius, diameter, mechanical_diameter)
        self.conic_constant, self.coefficients = float(conic_constant), [float(a) for a in coefficients]

    def sdf_forward(self):      # :452-462
        if math.isinf(self.radius):
            return None
        return sh.AsphericalSurfaceSDF(self.radius > 0, self.coefficients, self.radius, self.conic_constant, self.diameter)

    def sdf_backward(self):     # :464-474 (no rotation: the distance functions read the sign of the curvature)
        if math.isinf(self.radius):
            return None
        return sh.AsphericalSurfaceSDF(not (self.radius > 0), self.coefficients, self.radius, self.conic_constant, self.diameter)


def LensFromSurfaces(front, back_or_ct, ct_or_n, n=None):
    """Lens(front_surface, [back_surface,] center_thickness, n) for rotationally symmetric surfaces (Lenses.jl:176-311)."""
    if n is None:
        front_s, back_s, ct, n = front, CircularFlatSurface(front.diameter), back_or_ct, ct_or_n
    else:
        front_s, back_s, ct = front, back_or_ct, ct_or_n
    return Lens(lens_shape(front_s.radius, front_s.diameter, front_s.mechanical_diameter, back_s.radius, back_s.diameter,
                           back_s.mechanical_diameter, ct, front_s, back_s), n)


def lens_shape(r1, d1, md1, r2, d2, md2, ct, front_surface=None, back_surface=None):
    """Lenses.jl:176-311.  Without surface objects: SphericalSurface / CircularFlatSurface (r = Inf) pairs."""
    d_mid, md_mid = min(d1, d2), max(md1, md2)
    l0 = ct
    front = front_surface.sdf_forward() if front_surface is not None else _surf_forward(r1, d1)
    l0 -= front.thickness() if front is not None else 0.0
    back = back_surface.sdf_backward() if back_surface is not None else _surf_backward(r2, d2)
    l0 -= back.thickness() if back is not None else 0.0
    if front is None and back is None:
        return sh.PlanoSurfaceSDF(ct, d_mid)
    if l0 <= 0:
        if isinstance(front, sh.AsphericalSurfaceSDF) or isinstance(back, sh.AsphericalSurfaceSDF):
            raise ValueError("only spherical meniscus lenses are supported (Lenses.jl:188-190)")
        if _sign(r1) != _sign(r2):
            raise ValueError("Lens parameters lead to cylinder section length of <= 0, use ThinLens instead.")
        shape = _meniscus(r1, d1, front, r2, d2, back, ct)
        if md_mid > d_mid:
            th, p = shape.cylinder.thickness(), shape.cylinder.pos
            ring = sh.RingSDF(d_mid / 2, (md_mid - d_mid) / 2, th)
            ring.translate3d_((0.0, p[1] + th / 2, 0.0))
            shape = shape + ring
        return shape
    mid = sh.PlanoSurfaceSDF(l0, d_mid)
    if front is not None:
        mid.translate3d_((0.0, front.thickness(), 0.0))
        mid = mid + front
    if back is not None:
        back.translate3d_((0.0, mid.thickness() + back.thickness(), 0.0))
        mid = mid + back
    shape = mid
    d_front, d_back, d_min, d_max = d1, d2, min(d1, d2), max(d1, d2)
    if md_mid < d_min:
        return shape
    if d_front != d_back:
        if d_back > d_front:
            lt = l0
            if front is not None and front.sag() < 0:
                lt += abs(front.sag()) + front.thickness()
            ring = sh.RingSDF(d_front / 2, (d_back - d_front) / 2, lt)
            ring.translate3d_((0.0, (front.sag() if front is not None else 0.0) + lt / 2, 0.0))
        else:
            lt = l0
            if back is not None and (back.sag() - back.thickness()) > 0:
                lt += abs(back.sag()) + back.thickness()
            ring = sh.RingSDF(d_back / 2, (d_front - d_back) / 2, lt)
            ring.translate3d_((0.0, (front.thickness() if front is not None else 0.0) + lt / 2, 0.0))
        shape = shape + ring
    if md_mid > d_max:
        ot = mid.thickness()
        oc = mid.pos[1] + ot / 2
        if front is not None:
            ot -= front.sag(); oc += front.sag() / 2
        if back is not None:
            ot += back.sag(); oc += back.sag() / 2
        ring = sh.RingSDF(d_max / 2, (md_mid - d_max) / 2, ot)
        ring.translate3d_((0.0, oc, 0.0))
        shape = shape + ring
    return shape


class CylindricalSurface:
    """CylindricalSDF.jl:149-171: radius of curvature, clear aperture, height of the uncurved direction, mechanical diameter."""
    def __init__(self, radius, diameter, height, mechanical_diameter=None):
        self.radius, self.diameter, self.height = float(radius), float(diameter), float(height)
        self.mechanical_diameter = float(diameter if mechanical_diameter is None else mechanical_diameter)


class AcylindricalSurface(CylindricalSurface):
    """AcylindricalSDF.jl:136-168: a cylindric surface whose profile is an even asphere."""
    def __init__(self, radius, diameter, height, conic_constant, coefficients, mechanical_diameter=None):
        super().__init__(radius, diameter, height, mechanical_diameter)
        self.conic_constant, self.coefficients = float(conic_constant), [float(a) for a in coefficients]


class RectangularFlatSurface:   # CylindricalSDF.jl:218-223
    def __init__(self, size):
        self.size = self.diameter = float(size)


def cylindric_lens_shape(front, back, ct):
    """Lenses.jl:331-379 with the surface -> SDF rules of CylindricalSDF.jl:176-205."""
    def fwd(s):      # CylindricalSDF.jl:184-193, AcylindricalSDF.jl:184-191
        if math.isinf(s.radius):
            return None
        if isinstance(s, AcylindricalSurface):
            return sh.AcylindricalSurfaceSDF(s.radius > 0, s.radius, s.diameter, s.height, s.conic_constant, s.coefficients)
        return sh.ConvexCylinderSDF(s.radius, s.diameter, s.height) if s.radius > 0 else sh.ConcaveCylinderSDF(s.radius, s.diameter, s.height)

    def bwd(s):      # CylindricalSDF.jl:195-204, AcylindricalSDF.jl:192-199
        if isinstance(s, RectangularFlatSurface) or math.isinf(s.radius):
            return None
        if isinstance(s, AcylindricalSurface):
            if s.radius > 0:
                return sh.AcylindricalSurfaceSDF(False, s.radius, s.diameter, s.height, s.conic_constant, s.coefficients)
            return sh.AcylindricalSurfaceSDF(True, -s.radius, s.diameter, s.height, s.conic_constant, s.coefficients)
        return sh.ConcaveCylinderSDF(s.radius, s.diameter, s.height) if s.radius > 0 else sh.ConvexCylinderSDF(-s.radius, s.diameter, s.height)

    def edge_sag(sdf):   # thickness(sdf) for cylinders (CylindricalSDF.jl:173-174), the signed aspheric sag for acylinders (AcylindricalSDF.jl:170-178)
        return sdf.sag() if isinstance(sdf, sh.AcylindricalSurfaceSDF) else sdf.thickness()
    f, b = fwd(front), bwd(back)
    l0 = ct - (f.thickness() if f is not None else 0.0)
    l0 -= b.thickness() if b is not None else 0.0
    if isinstance(back, RectangularFlatSurface):
        d_mid, md_mid, h = front.diameter, front.mechanical_diameter, front.height      # :408-414
    else:
        if front.height != back.height:
            raise ValueError("height of front and back surface have to match for cylindric lenses")
        d_mid, md_mid, h = min(front.diameter, back.diameter), max(front.mechanical_diameter, back.mechanical_diameter), front.height
    if l0 <= 0:
        raise ValueError("Lens parameters lead to a box section length of <= 0")
    mid = sh.BoxSDF(h, l0, d_mid)
    mid.translate3d_((0.0, l0 / 2, 0.0))
    if f is not None:
        mid.translate3d_((0.0, f.thickness(), 0.0))
        mid = mid + f
    if b is not None:
        b.translate3d_((0.0, mid.thickness() + b.thickness(), 0.0))
        mid = mid + b
    shape = mid
    if md_mid > d_mid:
        rt, rc = mid.thickness(), mid.pos[1] + mid.thickness() / 2
        if f is not None:
            rt -= edge_sag(f); rc += edge_sag(f) / 2
        if b is not None:
            rt += edge_sag(b); rc += edge_sag(b) / 2
        ring = sh.RingSDF(d_mid / 2, (md_mid - d_mid) / 2, rt)
        ring.translate3d_((0.0, rc, 0.0))
        shape = shape + ring
    return shape


def CylindricalLens(front, back_or_ct, ct_or_n, n=None):
    """Lens(front::AbstractCylindricalSurface, [back,] center_thickness, n) (Lenses.jl:331-393)."""
    if n is None:
        front_s, back_s, ct, n = front, RectangularFlatSurface(front.diameter), back_or_ct, ct_or_n
    else:
        front_s, back_s, ct = front, back_or_ct, ct_or_n
    return Lens(cylindric_lens_shape(front_s, back_s, ct), n)


def ThinLens(R1, R2, d, n):                       # SphericalLenses.jl:42-46
    return Lens(sh.ThinLensSDF(R1, R2, d), n)


def SphericalLens(r1, r2, l, d=inch, n=1.5):      # SphericalLenses.jl:20-34
    if l == 0:
        return ThinLens(r1, r2, d, n)
    return Lens(lens_shape(r1, d, d, r2, d, d, l), n)


def SphericalDoubletLens(r1, r2, r3, l1, l2, d, n1, n2):   # DoubletLenses.jl:57-64
    front = SphericalLens(r1, r2, l1, d, n1)
    back = SphericalLens(r2, r3, l2, d, n2)
    back.translate3d_((0.0, front.shape.thickness(), 0.0))
    return DoubletLens(front, back)


def RightAnglePrism(leg_length, height, n):       # Prisms.jl:29-32
    return Prism(sh.RightAnglePrismSDF(leg_length, height), n)


def RectangularCompensatorPlate(width, height, thickness, n):   # Compensators.jl:15-24
    m = sh.CuboidMesh(width, thickness, height)
    m.translate3d_((-width / 2, 0.0, -height / 2))
    m.set_new_origin3d_()
    return Prism(m, n)


def RoundPlanoMirror(diameter, thickness):         # Mirrors.jl:154-157
    return Mirror(sh.PlanoSurfaceSDF(thickness, diameter))


def SquarePlanoMirror2D(size):                     # :88-91
    return Mirror(sh.QuadraticFlatMesh(size))


def RectangularPlanoMirror(width, height, thickness):   # :93-102
    m = sh.CuboidMesh(width, thickness, height)
    m.translate3d_((-width / 2, 0.0, -height / 2))
    m.set_new_origin3d_()
    return Mirror(m)


def SquarePlanoMirror(width, thickness):
    return RectangularPlanoMirror(width, width, thickness)


def ConcaveSphericalMirror(radius, thickness, diameter):   # :186-191
    cyl = sh.PlanoSurfaceSDF(thickness, diameter)
    cc = sh.ConcaveSphericalSurfaceSDF(abs(radius), diameter)
    return Mirror(cc + cyl)


def RightAnglePrismMirror(leg_length, height):     # :214-218
    s = sh.RightAnglePrismSDF(leg_length, height)
    s.zrotate3d_(la.deg2rad(45 + 180))
    return Mirror(s)


def Retroreflector(scale):                         # Misc.jl:34-49
    return Mirror(sh.RetroMesh(scale))


# ---- system --------------------------------------------------------------------------------------
class System:
    """System.jl:10-21.  `objects` may contain ObjectGroups; `leaves()` is `Leaves(system.objects)`."""

    def __init__(self, objects, n=1.0):
        self.objects = list(objects) if isinstance(objects, (list, tuple)) else [objects]
        self.n = float(n)
        self._device = None   # managed by solver.py

    def leaves(self):
        out = []

        def rec(o):
            if isinstance(o, ObjectGroup):
                for c in o.parts:
                    rec(c)
            else:
                out.append(o)
        for o in self.objects:
            rec(o)
        return out


StaticSystem = System   # System.jl:38-45: same data, flattens identically
