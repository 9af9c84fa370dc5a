"""Small FP64 vector / matrix helpers on plain Python floats.

Operation order mirrors the reference's StaticArrays expressions (left-to-right sums, no FMA), so
that poses built here are bit-identical to the ones the reference's kinematic API would produce
(`rotate3d`, `align3d`: src/Utils/LinearAlgebraUtils.jl:55-96).
"""
import math

EPS = 2.220446049250313e-16
SQRT_EPS = 1.4901161193847656e-8


def v3(x):
    return (float(x[0]), float(x[1]), float(x[2]))


def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def neg(a): return (-a[0], -a[1], -a[2])
def scale(s, a): return (s * a[0], s * a[1], s * a[2])
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def norm(a): return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def normalize(a):
    i = 1.0 / norm(a)   # StaticArrays: inv(norm(a)) * a
    return (i * a[0], i * a[1], i * a[2])


IDENTITY = ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))


def matvec(A, v):
    return (A[0][0] * v[0] + A[0][1] * v[1] + A[0][2] * v[2],
            A[1][0] * v[0] + A[1][1] * v[1] + A[1][2] * v[2],
            A[2][0] * v[0] + A[2][1] * v[1] + A[2][2] * v[2])


def matmul(A, B):
    return tuple(tuple(A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] for j in range(3)) for i in range(3))


def transpose(A):
    return tuple(tuple(A[j][i] for j in range(3)) for i in range(3))


def col(A, c):
    return (A[0][c], A[1][c], A[2][c])


def isapprox(x, y, atol=0.0, rtol=None):
    if rtol is None:
        rtol = 0.0 if atol > 0 else SQRT_EPS
    if x == y:
        return True
    if not (math.isfinite(x) and math.isfinite(y)):
        return False
    return abs(x - y) <= max(atol, rtol * max(abs(x), abs(y)))


def rotate3d(u, theta):
    """Rodrigues rotation matrix, entries exactly as in LinearAlgebraUtils.jl:55-65."""
    cost, sint = math.cos(theta), math.sin(theta)
    ux, uy, uz = float(u[0]), float(u[1]), float(u[2])
    return ((cost + ux * ux * (1 - cost), ux * uy * (1 - cost) - uz * sint, ux * uz * (1 - cost) + uy * sint),
            (uy * ux * (1 - cost) + uz * sint, cost + uy * uy * (1 - cost), uy * uz * (1 - cost) - ux * sint),
            (uz * ux * (1 - cost) - uy * sint, uz * uy * (1 - cost) + ux * sint, cost + uz * uz * (1 - cost)))


def align3d(start, target):
    """LinearAlgebraUtils.jl:74-96."""
    start = normalize(v3(start))
    target = normalize(v3(target))
    rx, ry, rz = cross(target, start)
    cosA = dot(start, target)
    if isapprox(cosA, 1.0):
        return IDENTITY
    if isapprox(cosA, -1.0):
        return ((-1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (0.0, 0.0, 1.0))
    k = 1 / (1 + cosA)
    return ((rx * rx * k + cosA, rx * ry * k + rz, rx * rz * k - ry),
            (ry * rx * k - rz, ry * ry * k + cosA, ry * rz * k + rx),
            (rz * rx * k + ry, rz * ry * k - rx, rz * rz * k + cosA))


def rotation_axis_angle(R):
    """Axis/angle used by reset_rotation3d! (AbstractShapeTrait.jl / Mesh.jl:149-163)."""
    c = (R[0][0] + R[1][1] + R[2][2] - 1) / 2
    c = 1.0 if c > 1.0 else (-1.0 if c < -1.0 else c)
    th = math.acos(c)
    if th == 0.0:
        return None, 0.0
    f = 1 / (2 * math.sin(th))
    return (f * (R[2][1] - R[1][2]), f * (R[0][2] - R[2][0]), f * (R[1][0] - R[0][1])), th


def deg2rad(d):
    return d * (math.pi / 180.0)
