"""Host-side arithmetic of the even-asphere surfaces (reference: src/SDFs/AsphericalLensSDF.jl:53-157,
src/Utils/MiscUtils.jl:87-110): sag equation, its radial derivative, the extremal sag used to close the
surface.  Only construction-time quantities are computed here; the per-ray evaluation is the CUDA code in
csrc/bmo_asphere.cuh.  `jl_pow` restates Julia's Base.Math.pow_body (see the note there)."""
import ctypes
import ctypes.util
import math

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double] * 3
_fma = _libm.fma


def jl_pow(x, n):
    if n == 0:
        return 1.0
    y, xnlo, ynlo = 1.0, 0.0, 0.0
    if n == 3:
        return x * x * x
    if n < 0:
        rx = 1.0 / x
        if n == -2:
            return rx * rx
        if math.isfinite(x):
            xnlo = -_fma(x, rx, -1.0) * rx
        x, n = rx, -n
    while n > 1:
        if n & 1:
            err = _fma(y, xnlo, x * ynlo)
            xy = x * y
            ynlo = _fma(x, y, -xy)
            y = xy
            ynlo += err
        err = x * 2 * xnlo
        xx = x * x
        xnlo = _fma(x, x, -xx)
        x = xx
        xnlo += err
        n >>= 1
    err = _fma(y, xnlo, x * ynlo)
    return _fma(x, y, err) if (math.isfinite(x) and math.isfinite(err)) else x * y


def aspheric_equation(r, c, k, coeffs):           # AsphericalLensSDF.jl:128-141
    r2 = r * r
    sqrt_arg = 1 - (1 + k) * (c * c) * r2
    if sqrt_arg < 0:
        return math.nan
    s = 0.0
    for i, a in enumerate(coeffs):
        t = a * jl_pow(r2, i + 1)
        s = t if i == 0 else s + t
    return c * r2 / (1 + math.sqrt(sqrt_arg)) + s


def gradient_aspheric_equation(r, c, k, coeffs):  # :147-157, first component
    Ri = 1 / c
    sqrt_arg = 1 - (r * r) * (1 + k) / (Ri * Ri)
    if sqrt_arg < 0:
        return math.nan
    sq = math.sqrt(sqrt_arg)
    gr = 2 * r / (Ri * (sq + 1)) + (r * r * r) * (1 + k) / ((Ri * Ri * Ri) * sq * ((sq + 1) * (sq + 1)))
    s = 0.0
    for i, a in enumerate(coeffs):
        m = i + 1
        t = float(2 * m) * a * jl_pow(r, 2 * (m - 1) + 1)
        s = t if i == 0 else s + t
    return -s - gr


def _sign(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else x)


def find_zero_bisection(f, a, b, tol=1e-10, max_iter=1000):   # MiscUtils.jl:87-110
    fa, fb = f(a), f(b)
    if _sign(fa) == _sign(fb):
        raise ValueError(f"Bisection requires a sign change: f(a)={fa}, f(b)={fb}")
    for _ in range(max_iter):
        mid = (a + b) / 2
        fmid = f(mid)
        if abs(fmid) < tol:
            return mid
        if _sign(fa) == _sign(fmid):
            a, fa = mid, fmid
        else:
            b, fb = mid, fmid
    raise ValueError(f"Bisection did not converge after {max_iter} iterations")


def max_aspheric_value(c, k, coeffs, d):          # AsphericalLensSDF.jl:53-68 -> (f(r_max), r_max)
    f = lambda r: aspheric_equation(r, c, k, coeffs)
    fp = lambda r: gradient_aspheric_equation(r, c, k, coeffs)
    a, b = 1e-8, d / 2
    if _sign(fp(a)) == _sign(fp(b)):
        r_max = a if abs(f(a)) > abs(f(b)) else b
    else:
        r_max = find_zero_bisection(fp, a, b)
    return f(r_max), r_max
