// bmo_asphere.cuh -- even-asphere surface pseudo-distances (SDFs/AsphericalLensSDF.jl), plain double.
// The normals of these surfaces always come from central differences (AsphericalLensSDF.jl:3-5), so there
// is no dual-number path.  Parameter block of one surface in SysView::ext (tables.ext):
//   [0] c = 1/radius  [1] conic constant  [2] diameter  [3] max_sag[1]  [4] aspheric_equation(d/2)
//   [5] gradient_aspheric_equation(d/2)[1]  [6] number of coefficients  [7...] coefficients
// ([3]-[5] depend on the parameters only; the reference recomputes [4], [5] on every call.)
// Citations are relative to /root/reference/src.
#pragma once
#include "bmo_math.cuh"

namespace bmo {

// Base.Math.pow_body(x::Float64, n::Integer) (Julia >= 1.8, base/math.jl): compensated power by squaring,
// muladd taken as fma.  Restated from the published algorithm (no Julia in the build image: last-bit parity
// of this routine with a given Julia version is unpinned; the reference's own asphere tests pin it at their tolerances).
BMO_D double jl_pow(double x, int n) {
    if (n == 0) return 1.0;
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    if (n == 3) return x * x * x;
    if (n < 0) {
        const double rx = 1.0 / x;
        if (n == -2) return rx * rx;
        if (isfinite(x)) xnlo = -fma(x, rx, -1.0) * rx;
        x = rx;
        n = -n;
    }
    while (n > 1) {
        if (n & 1) {
            const double err = fma(y, xnlo, x * ynlo);
            const double xy = x * y;
            ynlo = fma(x, y, -xy);
            y = xy;
            ynlo += err;
        }
        const double err = x * 2 * xnlo;
        const double xx = x * x;
        xnlo = fma(x, x, -xx);
        x = xx;
        xnlo += err;
        n >>= 1;
    }
    const double err = fma(y, xnlo, x * ynlo);
    return (isfinite(x) && isfinite(err)) ? fma(x, y, err) : x * y;
}
// ---- helpers so that the surface functions are written once for double and for dual numbers ---------------------
BMO_D bool operator>(Dual a, Dual b) { return a.v > b.v; }
BMO_D bool operator>(Dual a, double b) { return a.v > b; }
BMO_D Dual operator/(Dual a, Dual b) {      // ForwardDiff: pa * inv(vy) + pb * -(vx / (vy * vy))
    const double ib = 1.0 / b.v, c = -(a.v / (b.v * b.v));
    return mkd(a.v / b.v, a.p0 * ib + b.p0 * c, a.p1 * ib + b.p1 * c, a.p2 * ib + b.p2 * c);
}
BMO_D Dual operator/(double a, Dual b) {
    const double q = a / b.v, c = -(q / b.v);
    return mkd(q, b.p0 * c, b.p1 * c, b.p2 * c);
}
BMO_D Dual jl_pow(Dual x, int n) {          // d/dx x^n = n x^(n-1)
    const double v = jl_pow(x.v, n), dv = n == 0 ? 0.0 : (double)n * jl_pow(x.v, n - 1);
    return mkd(v, x.p0 * dv, x.p1 * dv, x.p2 * dv);
}
BMO_D bool isnan_(double a) { return isnan(a); }
BMO_D bool isnan_(Dual a) { return isnan(a.v); }
BMO_D double clamp_(double x, double lo, double hi) { return jl_clamp(x, lo, hi); }
BMO_D Dual clamp_(Dual x, double lo, double hi) { return x.v > hi ? mkd(hi, 0, 0, 0) : (x.v < lo ? mkd(lo, 0, 0, 0) : x); }
BMO_D double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
BMO_D double jl_sign(Dual x) { return jl_sign(x.v); }
BMO_D double mkT(double, double c) { return c; }
BMO_D Dual mkT(Dual, double c) { return mkd(c, 0, 0, 0); }
template <class T> BMO_D T min3_(T a, T b, T c) { return min_(min_(a, b), c); }
template <class T> BMO_D T min4_(T a, T b, T c, T d) { return min_(min_(min_(a, b), c), d); }

// AsphericalLensSDF.jl:128-141
template <class T> BMO_D T aspheric_equation(T r, const double* e) {
    const double c = e[0], k = e[1];
    const int nc = (int)e[6];
    const T r2 = r * r;
    const T sqrt_arg = 1 - (1 + k) * (c * c) * r2;
    if (sqrt_arg < 0.0) return mkT(r, nan(""));
    T sum_a = r2 * 0.0;
    for (int i = 0; i < nc; i++) {
        const T t = e[7 + i] * jl_pow(r2, i + 1);
        sum_a = i == 0 ? t : sum_a + t;
    }
    return c * r2 / (1 + sqrt_(sqrt_arg)) + sum_a;
}
// :147-157, first component (the second is 1)
template <class T> BMO_D T gradient_aspheric_equation(T r, const double* e) {
    const double c = e[0], k = e[1];
    const int nc = (int)e[6];
    const double Ri = 1 / c;
    const T sqrt_arg = 1 - (r * r) * (1 + k) / (Ri * Ri);
    if (sqrt_arg < 0.0) return mkT(r, nan(""));
    const T sq = sqrt_(sqrt_arg);
    const T gr = 2 * r / (Ri * (sq + 1)) + (r * r * r) * (1 + k) / ((Ri * Ri * Ri) * sq * ((sq + 1) * (sq + 1)));
    T sum_r = r * 0.0;
    for (int i = 0; i < nc; i++) {
        const int m = i + 1;
        const T t = (double)(2 * m) * e[7 + i] * jl_pow(r, 2 * (m - 1) + 1);
        sum_r = i == 0 ? t : sum_r + t;
    }
    return -sum_r - gr;
}
// :165-170
template <class T> BMO_D T sd_line_segment(T px, T py, double ax, double ay, double bx, double by) {
    const T pax = px - ax, pay = py - ay;
    const double bax = bx - ax, bay = by - ay;
    const T h = clamp_((pax * bax + pay * bay) / (bax * bax + bay * bay), 0.0, 1.0);
    const T ex = pax - h * bax, ey = pay - h * bay;
    return sqrt_(ex * ex + ey * ey);
}
// :188-240 (convex) and :242-307 (concave); r = (signed) distance from the optical axis, z = position along it
template <class T> BMO_NI T aspheric_surface_distance(bool convex, T r, T z, const double* e) {
    const double c = e[0], d = e[2], ms = e[3], zb = e[4];
    const T r2 = r * r;
    const double r2_bound = (d / 2) * (d / 2);
    const T zv = aspheric_equation(r, e);
    const T g = gradient_aspheric_equation(r, e);
    const double n_gzb = sqrt(e[5] * e[5] + 1.0 * 1.0);
    const T rr = r - jl_sign(r) * d / 2;
    if (convex) {
        if (isnan_(zv) || isnan_(g) || r2 > mkT(r, r2_bound)) {
            T dist;
            if (z < zb) dist = sqrt_(rr * rr + (z - zb) * (z - zb));
            else if (z > zb && z < 0.0) dist = sqrt_(rr * rr);
            else if (z > 0.0 && (jl_sign(c) == 1 && zb < 0)) dist = sqrt_(rr * rr + z * z);
            else dist = sqrt_(rr * rr + (z - zb) * (z - zb));
            return dist / n_gzb;
        }
        const T da = abs_(z - zv) / sqrt_(g * g + 1.0 * 1.0);
        if (jl_sign(c) == 1 && zb < 0) {
            const T s1 = sd_line_segment(r, z, d / 2, zb, d / 2, ms) / n_gzb;
            const T s2 = sd_line_segment(r, z, d / 2, ms, -d / 2, ms) / n_gzb;
            const T s3 = sd_line_segment(r, z, -d / 2, ms, -d / 2, zb) / n_gzb;
            const T m = min4_(da, s1, s2, s3);
            return (zv < z && z < ms) ? -m : m;
        }
        const T sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        const double sc = jl_sign(c);
        const T m = min_(sdl, da);
        return (sc * zv < sc * z && sc * z < sc * zb) ? -m : m;
    }
    if (isnan_(zv) || isnan_(g)) {
        T dist;
        if (z < 0.0) dist = sqrt_(rr * rr + z * z);
        else if (z > 0.0 && z < zb) dist = sqrt_(rr * rr);
        else dist = sqrt_(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    const T da = abs_(z - zv) / sqrt_(g * g + 1.0 * 1.0);
    if (ms > 0 && zb < 0) {
        const T sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        if (r2 > mkT(r, r2_bound)) return sdl;
        const T m = min_(da, sdl);
        if (z > zb && z < zv) return -m;
        if (zb > 0 && (z > 0.0 && z < zv)) return -m;
        return m;
    }
    const T s1 = sd_line_segment(r, z, d / 2, zb, d / 2, 0.0) / n_gzb;
    const T s2 = sd_line_segment(r, z, d / 2, 0.0, -d / 2, 0.0) / n_gzb;
    const T s3 = sd_line_segment(r, z, -d / 2, 0.0, -d / 2, zb) / n_gzb;
    if (r2 > mkT(r, r2_bound)) return min3_(s1, s2, s3);
    const T m = min4_(da, s1, s2, s3);
    if (zb < 0 && (zv < z && z < 0.0)) return -m;
    if (zb > 0 && (z > 0.0 && z < zv)) return -m;
    return m;
}

}  // namespace bmo
