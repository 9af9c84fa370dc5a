"""Host-side shapes of the reference's geometry layer (L3): SDF primitives, UnionSDF,
MeniscusLensSDF and Mesh, with the kinematic API (`translate3d!` -> `translate3d_`, ...).

Only construction and kinematics live here -- everything that is evaluated per ray
(`sdf`, `normal3d`, `intersect3d`) runs on the GPU from the flattened tables (flatten.py).
Citations are relative to /root/reference/src.
"""
import math

import numpy as np

from . import linalg as la

# bmo_prim_type (include/bmo.h)
PLANO, CYLINDER, SPHERE, CONVEX, CONCAVE, CUTSPHERE, BOX, RING, RAPRISM, MENISCUS, CONVEX_CYL, CONCAVE_CYL, CONVEX_ASPH, CONCAVE_ASPH, CONVEX_ACYL, CONCAVE_ACYL = range(16)


def sag(r, l):
    """Utils/OpticUtils.jl:153"""
    return r - math.sqrt(r * r - 0.25 * (l * l))


def check_sag(r, d):
    """Utils/OpticUtils.jl:155-161"""
    if abs(2 * r) < d:
        raise ValueError(f"Radius of curvature (r = {r}) must be >= than half the diameter (d = {d}) or an illegal shape results!")


class AbstractShape:
    """AbstractTypes/AbstractShape.jl:41-114"""

    def __init__(self):
        self.pos = (0.0, 0.0, 0.0)
        self.dir = la.IDENTITY

    def position(self): return self.pos
    def orientation(self): return self.dir
    def _set_dir(self, d): self.dir = d

    def translate3d_(self, offset):
        self.pos = la.add(self.pos, la.v3(offset))

    def translate_to3d_(self, target):
        self.translate3d_(la.sub(la.v3(target), self.pos))

    def rotate3d_(self, axis, theta):
        self._set_dir(la.matmul(la.rotate3d(axis, theta), self.dir))

    def xrotate3d_(self, theta): self.rotate3d_((1.0, 0.0, 0.0), theta)
    def yrotate3d_(self, theta): self.rotate3d_((0.0, 1.0, 0.0), theta)
    def zrotate3d_(self, theta): self.rotate3d_((0.0, 0.0, 1.0), theta)

    def align3d_(self, target_axis):
        R = la.align3d(la.col(self.dir, 1), target_axis)
        self._set_dir(la.matmul(R, self.dir))

    def reset_translation3d_(self): self.pos = (0.0, 0.0, 0.0)
    def reset_rotation3d_(self): self._set_dir(la.IDENTITY)
    def has_thickness(self): return False
    def thickness(self): return 0.0


class AbstractSDF(AbstractShape):
    """SDFs/AbstractSDF.jl:18-40: keeps `transposed_dir` next to `dir`."""

    def __init__(self):
        super().__init__()
        self.tdir = la.IDENTITY

    def _set_dir(self, d):
        self.dir = d
        self.tdir = la.transpose(d)

    def __add__(self, other):
        """UnionSDF.jl:58-61: `+` builds a flattened union."""
        members = []
        for s in (self, other):
            members.extend(s.sdfs if isinstance(s, UnionSDF) else [s])
        return UnionSDF(members)

    def local_bound(self):
        """(centre, radius) of a sphere enclosing the region where the SDF can be < margin, in the
        frame the shape's own pose maps from (world for top-level shapes)."""
        raise NotImplementedError


class PrimSDF(AbstractSDF):
    def __init__(self, ptype, par):
        super().__init__()
        self.type = ptype
        self.par = tuple(float(x) for x in par) + (0.0,) * (4 - len(par))

    def _set_dir(self, d):
        if self.type == SPHERE:   # SphericalLensSDF.jl:82-84: orientation fixed to I
            return
        super()._set_dir(d)

    def has_thickness(self):
        return self.type in (PLANO, SPHERE, CONVEX, CONCAVE, BOX, CONVEX_CYL, CONCAVE_CYL)

    def thickness(self):
        a, b, c, _ = self.par
        if self.type == CONVEX_CYL:
            return self._thickness            # CylindricalSDF.jl:57  abs(sag(radius, diameter))
        return {PLANO: a, SPHERE: 2 * a, CONVEX: c, CONCAVE: 0.0, BOX: 2 * b}.get(self.type, 0.0)

    def diameter(self):
        if self.type == CONVEX_CYL:
            return self._diameter
        return 2 * self.par[0] if self.type == SPHERE else self.par[1]

    def sag(self):
        return self.par[2]

    def local_box(self):
        """Axis-aligned box, in the primitive's own frame, that contains every point with sdf <= 0."""
        a, b, c, d = self.par
        t = self.type
        if t == PLANO: return (-b / 2, 0.0, -b / 2), (b / 2, a, b / 2)
        if t == CYLINDER: return (-a, -b, -a), (a, b, a)
        if t in (SPHERE, CUTSPHERE): return (-a, -a, -a), (a, a, a)
        if t == CONVEX: return (-b / 2, 0.0, -b / 2), (b / 2, c, b / 2)          # cap between the vertex y = 0 and y = sag
        if t == CONCAVE: return (-b / 2, -c, -b / 2), (b / 2, 0.0, b / 2)        # cylinder slab y in [-sag, 0] minus the sphere
        if t in (BOX, RAPRISM): return (-a, -b, -c), (a, b, c)
        if t == RING: return (-(a + b), -c, -(a + b)), (a + b, c, a + b)
        if t == CONVEX_CYL: return (-d, -c, b), (d, c, a)      # |x| <= height/2, |y| <= w, cut height <= z <= radius
        if t == CONCAVE_CYL:                                    # the box of CylindricalSDF.jl:129-131 (the cylinder is cut out of it)
            y0 = d / 2 if a > 0 else -d / 2
            return (-c / 2, y0 - d / 2, -b / 2), (c / 2, y0 + d / 2, b / 2)
        raise ValueError(t)

    def box_points(self):
        """Corners of local_box() in the frame this shape's pose maps to (world for top-level shapes)."""
        lo, hi = self.local_box()
        return [la.add(self.pos, la.matvec(self.dir, q)) for q in _corners(lo, hi)]

    def local_bound(self):
        a, b, c, d = self.par
        t = self.type
        if t == PLANO: cl, r = (0.0, a / 2, 0.0), math.hypot(b / 2, a / 2)
        elif t == CYLINDER: cl, r = (0.0, 0.0, 0.0), math.hypot(a, b)
        elif t == SPHERE: cl, r = (0.0, 0.0, 0.0), a
        elif t == CONVEX: cl, r = (0.0, c / 2, 0.0), math.hypot(b / 2, c / 2)
        elif t == CONCAVE: cl, r = (0.0, -c / 2, 0.0), math.hypot(b / 2, c / 2)
        elif t == CUTSPHERE: cl, r = (0.0, 0.0, 0.0), a
        elif t in (BOX, RAPRISM): cl, r = (0.0, 0.0, 0.0), math.sqrt(a * a + b * b + c * c)
        elif t == RING: cl, r = (0.0, 0.0, 0.0), math.hypot(a + b, c)
        elif t in (CONVEX_CYL, CONCAVE_CYL):
            lo, hi = self.local_box()
            cl = tuple((lo[k] + hi[k]) / 2 for k in range(3))
            r = math.sqrt(sum(((hi[k] - lo[k]) / 2) ** 2 for k in range(3)))
        else: raise ValueError(t)
        return la.add(self.pos, la.matvec(self.dir, cl)), r


def PlanoSurfaceSDF(thickness, diameter): return PrimSDF(PLANO, (thickness, diameter))           # SphericalLensSDF.jl:49-58
def CylinderSDF(r, h): return PrimSDF(CYLINDER, (r, h))                                        # PrimitiveSDF.jl:61-69
def SphereSDF(r): return PrimSDF(SPHERE, (r,))                                                 # SphericalLensSDF.jl:80
def BoxSDF(x, y, z): return PrimSDF(BOX, (x / 2, y / 2, z / 2))                                # PrimitiveSDF.jl:29-37
def RingSDF(inner_radius, width, thickness): return PrimSDF(RING, (inner_radius + width / 2, width / 2, thickness / 2))  # :146-155
def RightAnglePrismSDF(leg, height): return PrimSDF(RAPRISM, (leg / 2, leg / 2, height / 2))    # :195-202


def ConvexSphericalSurfaceSDF(radius, diameter):   # SphericalLensSDF.jl:203-217
    check_sag(radius, diameter)
    s = sag(radius, diameter)
    return PrimSDF(CONVEX, (radius, diameter, s, radius - s))


def ConcaveSphericalSurfaceSDF(radius, diameter):  # :147-157
    check_sag(radius, diameter)
    return PrimSDF(CONCAVE, (radius, diameter, sag(radius, diameter)))


class AsphericalSurfaceSDF(PrimSDF):
    """Convex / ConcaveAsphericalSurfaceSDF (AsphericalLensSDF.jl:22-31, 88-98).  `ext` is the parameter block the
    device reads (csrc/bmo_asphere.cuh): c, k, d, max_sag[1], sag(d/2), sag'(d/2), n, coefficients."""

    def __init__(self, convex, coefficients, radius, conic_constant, diameter):
        super().__init__(CONVEX_ASPH if convex else CONCAVE_ASPH, (0.0, 0.0, 0.0, 0.0))
        from . import asphere as asp
        self.coefficients = [float(a) for a in coefficients]
        self.radius, self.conic_constant, self._diameter = float(radius), float(conic_constant), float(diameter)
        c = 1 / self.radius
        self.max_sag = asp.max_aspheric_value(c, self.conic_constant, self.coefficients, self._diameter)
        self._edge = asp.aspheric_equation(self._diameter / 2, c, self.conic_constant, self.coefficients)
        gzb = asp.gradient_aspheric_equation(self._diameter / 2, c, self.conic_constant, self.coefficients)
        self.ext = [c, self.conic_constant, self._diameter, self.max_sag[0], self._edge, gzb, float(len(self.coefficients))] + self.coefficients

    def has_thickness(self): return True

    def thickness(self):          # :33-36, :100-103
        sg, ms = self._edge, self.max_sag[0]
        if self.type == CONVEX_ASPH:
            return ms if (ms > 0 and sg < 0) else abs(sg)
        return abs(sg) if (ms > 0 and sg < 0) else 0.0

    def diameter(self): return self._diameter
    def sag(self): return self._edge      # edge_sag(::EvenAsphericalSurface, sdf), :434-443

    def local_box(self):
        # everything with sdf <= 0 lies within the aperture, between the extreme sag values and the closing planes
        h = self._diameter / 2
        zs = (0.0, self._edge, self.max_sag[0])
        return (-h, min(zs), -h), (h, max(zs), h)

    def local_bound(self):
        lo, hi = self.local_box()
        cl = tuple((lo[k] + hi[k]) / 2 for k in range(3))
        r = math.sqrt(sum(((hi[k] - lo[k]) / 2) ** 2 for k in range(3)))
        return la.add(self.pos, la.matvec(self.dir, cl)), r


class AcylindricalSurfaceSDF(AsphericalSurfaceSDF):
    """Aconvex / AconcaveCylinderSDF (AcylindricalSDF.jl:14-120): the aspheric profile extruded along x over `height`."""

    def __init__(self, convex, radius, diameter, height, conic_constant, coefficients):
        super().__init__(convex, coefficients, radius, conic_constant, diameter)
        self.type = CONVEX_ACYL if convex else CONCAVE_ACYL
        self.height = float(height)
        self.par = (0.0, self.height / 2, 0.0, 0.0)

    def thickness(self):          # :52-54, :97-100
        sg, ms = self._edge, self.max_sag[0]
        if self.type == CONVEX_ACYL:
            return abs(sg)
        return abs(sg) if (ms > 0 and sg < 0) else 0.0

    def local_box(self):
        h = self._diameter / 2
        zs = (0.0, self._edge, self.max_sag[0])
        return (-self.height / 2, min(zs), -h), (self.height / 2, max(zs), h)


def ConvexAsphericalSurfaceSDF(coefficients, radius, conic_constant, diameter):
    return AsphericalSurfaceSDF(True, coefficients, radius, conic_constant, diameter)


def ConcaveAsphericalSurfaceSDF(coefficients, radius, conic_constant, diameter):
    return AsphericalSurfaceSDF(False, coefficients, radius, conic_constant, diameter)


def ConvexCylinderSDF(radius, diameter, height):   # CylindricalSDF.jl:30-55: cut cylinder built along x, turned upright, vertex at the origin
    h = math.sqrt(radius * radius - (diameter / 2) * (diameter / 2))
    s = PrimSDF(CONVEX_CYL, (radius, h, math.sqrt(radius * radius - h * h), height / 2))
    s._thickness, s._diameter = abs(sag(radius, diameter)), diameter
    s.xrotate3d_(math.pi / 2)
    s.translate3d_((0.0, radius, 0.0))
    return s


def ConcaveCylinderSDF(radius, diameter, height):  # :92-116 (the sign of the radius selects the side the cylinder is cut from)
    return PrimSDF(CONCAVE_CYL, (radius, diameter, height, sag(abs(radius), diameter)))


def CutSphereSDF(radius, height):                  # PrimitiveSDF.jl:97-110
    if abs(height) >= radius:
        raise ValueError("Cut off height must be smaller than radius")
    return PrimSDF(CUTSPHERE, (radius, height, math.sqrt(radius * radius - height * height)))


def _corners(lo, hi):
    return [(x, y, z) for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])]


def _box_of(points):
    return (tuple(min(q[k] for q in points) for k in range(3)), tuple(max(q[k] for q in points) for k in range(3)))


def _enclose(spheres):
    n = len(spheres)
    c = tuple(sum(s[0][k] for s in spheres) / n for k in range(3))
    r = max(la.norm(la.sub(s[0], c)) + s[1] for s in spheres)
    return c, r


class MeniscusLensSDF(AbstractSDF):
    """SDFs/MeniscusLensSDF.jl:20-46: max(min(convex, cylinder), -concave); children are posed
    relative to the meniscus frame."""

    def __init__(self, convex, cylinder, concave, thickness):
        super().__init__()
        self.convex, self.cylinder, self.concave, self._thickness = convex, cylinder, concave, float(thickness)

    def has_thickness(self): return True
    def thickness(self): return self._thickness
    def diameter(self): return self.cylinder.diameter()

    def local_bound(self):
        c, r = _enclose([self.convex.local_bound(), self.cylinder.local_bound()])
        return la.add(self.pos, la.matvec(self.dir, c)), r

    def box_points(self):   # max(min(convex, cylinder), -concave) <= min(convex, cylinder): inside convex or cylinder
        pts = self.convex.box_points() + self.cylinder.box_points()
        return [la.add(self.pos, la.matvec(self.dir, q)) for q in pts]


class UnionSDF(AbstractSDF):
    """SDFs/UnionSDF.jl"""

    def __init__(self, sdfs):
        super().__init__()
        self.sdfs = list(sdfs)

    def has_thickness(self): return True

    def thickness(self):   # :33-42
        t = 0.0
        for s in self.sdfs:
            if s.has_thickness():
                t += s.thickness()
        return t

    def translate3d_(self, offset):   # :63-67
        offset = la.v3(offset)
        self.pos = la.add(self.pos, offset)
        for s in self.sdfs:
            s.translate3d_(offset)

    def rotate3d_(self, axis, theta):  # :69-82
        R = la.rotate3d(axis, theta)
        self._set_dir(la.matmul(R, self.dir))
        for s in self.sdfs:
            s.rotate3d_(axis, theta)
            v = la.sub(s.pos, self.pos)
            v = la.sub(la.matvec(R, v), v)
            s.translate3d_(v)

    def local_bound(self):
        return _enclose([s.local_bound() for s in self.sdfs])

    def box_points(self):
        return [q for s in self.sdfs for q in s.box_points()]


def ThinLensSDF(r1, r2, d=25.4e-3):   # SphericalLensSDF.jl:245-253
    front = ConvexSphericalSurfaceSDF(r1, d)
    back = ConvexSphericalSurfaceSDF(r2, d)
    back.translate3d_((0.0, front.thickness() + back.thickness(), 0.0))
    back.zrotate3d_(math.pi)
    return front + back


# ---------------------------------------------------------------------------------------------
class Mesh(AbstractShape):
    """Mesh.jl:33-132.  Vertices are stored in WORLD coordinates; kinematics rewrite them.
    `f32=True` reproduces Mesh{Float32} (STL files, Mesh.jl:48-70)."""

    def __init__(self, vertices, faces, scale=1.0, f32=False):
        super().__init__()
        self.f32 = bool(f32)
        self.vertices = np.array(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.array(faces, dtype=np.int32).reshape(-1, 3)   # 0-based
        self.scale = float(scale)
        self._round()

    def _round(self):
        if self.f32:
            self.vertices = self.vertices.astype(np.float32).astype(np.float64)
            self.pos = tuple(float(np.float32(x)) for x in self.pos)
            self.dir = tuple(tuple(float(np.float32(x)) for x in row) for row in self.dir)

    def translate3d_(self, offset):   # :78-82
        offset = la.v3(offset)
        self.pos = la.add(self.pos, offset)
        self.vertices = self.vertices + np.array(offset)
        self._round()

    def _apply(self, R):              # (V .- pos') * R' .+ pos'
        d = self.vertices - np.array(self.pos)
        if self.f32:
            d = d.astype(np.float32).astype(np.float64)
        r = np.empty_like(d)
        for i in range(3):
            r[:, i] = d[:, 0] * R[i][0] + d[:, 1] * R[i][1] + d[:, 2] * R[i][2]
        self.vertices = r + np.array(self.pos)

    def rotate3d_(self, axis, theta):  # :89-96
        R = la.rotate3d(axis, theta)
        self._apply(R)
        self.dir = la.matmul(R, self.dir)
        self._round()

    def align3d_(self, target_axis):   # :113-120 (note: dir * R)
        R = la.align3d(la.col(self.dir, 1), target_axis)
        self._apply(R)
        self.dir = la.matmul(self.dir, R)
        self._round()

    def reset_translation3d_(self):    # :139-142
        self.translate3d_(la.neg(self.pos))

    def reset_rotation3d_(self):       # :149-163
        axis, th = la.rotation_axis_angle(self.dir)
        if axis is None:
            return
        self.rotate3d_(axis, -th)
        self.dir = la.IDENTITY

    def set_new_origin3d_(self):       # :171-175
        self.dir = la.IDENTITY
        self.pos = (0.0, 0.0, 0.0)

    def local_bound(self):
        c = self.vertices.mean(axis=0)
        r = float(np.sqrt(((self.vertices - c) ** 2).sum(axis=1)).max())
        return tuple(float(x) for x in c), r

    def box_points(self):
        lo, hi = self.vertices.min(axis=0), self.vertices.max(axis=0)
        return [tuple(float(x) for x in lo), tuple(float(x) for x in hi)]


def RectangularFlatMesh(width, height):   # Mesh.jl:282-303
    x, z = width / 2, height / 2
    return Mesh([[x, 0, z], [x, 0, -z], [-x, 0, -z], [-x, 0, z]], [[0, 1, 3], [1, 2, 3]])


def QuadraticFlatMesh(width):
    return RectangularFlatMesh(width, width)


def CircularFlatMesh(radius, n=30):       # Mesh.jl:322-348
    stop = 2 * math.pi * (n - 1) / n
    verts = [[0.0, 0.0, 0.0]]
    for i in range(n):
        t = 0.0 if n == 1 else i / (n - 1)
        x = (1 - t) * 0.0 + t * stop       # LinRange lerp
        verts.append([math.cos(x) * radius, 0.0, math.sin(x) * radius])
    faces = [[0, i - 1, i] for i in range(2, n + 2)]
    faces[n - 1][2] = 1                     # faces[end] = 2 (1-based)
    return Mesh(verts, faces)


def CuboidMesh(x, y, z, theta=math.pi / 2):  # Mesh.jl:362-395
    dx = math.cos(theta) * y
    verts = [[0, 0, 0], [x, 0, 0], [x + dx, y, 0], [0 + dx, y, 0], [0 + dx, y, z], [x + dx, y, z], [x, 0, z], [0, 0, z]]
    f = [[1, 3, 2], [1, 4, 3], [3, 4, 5], [3, 5, 6], [2, 3, 6], [2, 6, 7], [1, 8, 5], [1, 5, 4], [6, 5, 8], [6, 8, 7], [1, 7, 8], [1, 2, 7]]
    return Mesh(verts, [[a - 1, b - 1, c - 1] for a, b, c in f])


def CubeMesh(scale):
    return CuboidMesh(float(scale), float(scale), float(scale))


def RetroMesh(scale):                      # OpticalComponents/Misc.jl:8-22
    verts = [[0 * scale, 0 * scale, 0 * scale], [1 * scale, 0 * scale, 0 * scale], [0 * scale, 1 * scale, 0 * scale], [0 * scale, 0 * scale, 1 * scale]]
    return Mesh(verts, [[0, 2, 1], [0, 3, 2], [0, 1, 3]], scale=scale)


def load_stl(path):
    """Binary STL -> Mesh{Float32} scaled by Float32(1e-3) in Float32 (Mesh.jl:48-70, MeshIO)."""
    with open(path, "rb") as f:
        f.read(80)
        n = int(np.frombuffer(f.read(4), dtype="<u4")[0])
        rec = np.frombuffer(f.read(50 * n), dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    v32 = rec["v"].reshape(-1, 3).astype(np.float32) * np.float32(1e-3)
    faces = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    return Mesh(v32.astype(np.float64), faces, scale=float(np.float32(1e-3)), f32=True)
