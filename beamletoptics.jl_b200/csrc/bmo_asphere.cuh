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
BMO_D double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
// AsphericalLensSDF.jl:128-141
BMO_D double aspheric_equation(double r, const double* e) {
    const double c = e[0], k = e[1];
    const int nc = (int)e[6];
    const double r2 = r * r;
    const double sqrt_arg = 1 - (1 + k) * (c * c) * r2;
    if (sqrt_arg < 0) return nan("");
    double sum_a = 0.0;
    for (int i = 0; i < nc; i++) {
        const double t = e[7 + i] * jl_pow(r2, i + 1);
        sum_a = i == 0 ? t : sum_a + t;
    }
    return c * r2 / (1 + sqrt(sqrt_arg)) + sum_a;
}
// :147-157, first component (the second is 1)
BMO_D double gradient_aspheric_equation(double r, const double* e) {
    const double c = e[0], k = e[1];
    const int nc = (int)e[6];
    const double Ri = 1 / c;
    const double sqrt_arg = 1 - (r * r) * (1 + k) / (Ri * Ri);
    if (sqrt_arg < 0) return nan("");
    const double sq = sqrt(sqrt_arg);
    const double gr = 2 * r / (Ri * (sq + 1)) + (r * r * r) * (1 + k) / ((Ri * Ri * Ri) * sq * ((sq + 1) * (sq + 1)));
    double sum_r = 0.0;
    for (int i = 0; i < nc; i++) {
        const int m = i + 1;
        const double t = (double)(2 * m) * e[7 + i] * jl_pow(r, 2 * (m - 1) + 1);
        sum_r = i == 0 ? t : sum_r + t;
    }
    return -sum_r - gr;
}
// :165-170
BMO_D double sd_line_segment(double px, double py, double ax, double ay, double bx, double by) {
    const double pax = px - ax, pay = py - ay, bax = bx - ax, bay = by - ay;
    const double h = jl_clamp((pax * bax + pay * bay) / (bax * bax + bay * bay), 0.0, 1.0);
    const double ex = pax - h * bax, ey = pay - h * bay;
    return sqrt(ex * ex + ey * ey);
}
BMO_D double jl_min3(double a, double b, double c) { return jl_min(jl_min(a, b), c); }
BMO_D double jl_min4(double a, double b, double c, double d) { return jl_min(jl_min(jl_min(a, b), c), d); }
// :188-240 (convex) and :242-307 (concave); r = distance from the optical axis, z = position along it
BMO_NI double aspheric_surface_distance(bool convex, double r, double z, const double* e) {
    const double c = e[0], d = e[2], ms = e[3], zb = e[4];
    const double r2 = r * r, r2_bound = (d / 2) * (d / 2);
    const double zv = aspheric_equation(r, e);
    const double g = gradient_aspheric_equation(r, e);
    const double n_gzb = sqrt(e[5] * e[5] + 1.0 * 1.0);
    const double rr = r - jl_sign(r) * d / 2;
    if (convex) {
        if (isnan(zv) || isnan(g) || r2 > r2_bound) {
            double dist;
            if (z < zb) dist = sqrt(rr * rr + (z - zb) * (z - zb));
            else if (zb < z && z < 0) dist = sqrt(rr * rr);
            else if (z > 0 && (jl_sign(c) == 1 && zb < 0)) dist = sqrt(rr * rr + z * z);
            else dist = sqrt(rr * rr + (z - zb) * (z - zb));
            return dist / n_gzb;
        }
        const double da = fabs(z - zv) / sqrt(g * g + 1.0 * 1.0);
        if (jl_sign(c) == 1 && zb < 0) {
            const double s1 = sd_line_segment(r, z, d / 2, zb, d / 2, ms) / n_gzb;
            const double s2 = sd_line_segment(r, z, d / 2, ms, -d / 2, ms) / n_gzb;
            const double s3 = sd_line_segment(r, z, -d / 2, ms, -d / 2, zb) / n_gzb;
            const double m = jl_min4(da, s1, s2, s3);
            return (zv < z && z < ms) ? -m : m;
        }
        const double sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        const double sc = jl_sign(c);
        const double m = jl_min(sdl, da);
        return (sc * zv < sc * z && sc * z < sc * zb) ? -m : m;
    }
    if (isnan(zv) || isnan(g)) {
        double dist;
        if (z < 0) dist = sqrt(rr * rr + z * z);
        else if (0 < z && z < zb) dist = sqrt(rr * rr);
        else dist = sqrt(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    const double da = fabs(z - zv) / sqrt(g * g + 1.0 * 1.0);
    if (ms > 0 && zb < 0) {
        const double sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        if (r2 > r2_bound) return sdl;
        const double m = jl_min(da, sdl);
        if (zb < z && z < zv) return -m;
        if (zb > 0 && (0.0 < z && z < zv)) return -m;
        return m;
    }
    const double s1 = sd_line_segment(r, z, d / 2, zb, d / 2, 0.0) / n_gzb;
    const double s2 = sd_line_segment(r, z, d / 2, 0.0, -d / 2, 0.0) / n_gzb;
    const double s3 = sd_line_segment(r, z, -d / 2, 0.0, -d / 2, zb) / n_gzb;
    if (r2 > r2_bound) return jl_min3(s1, s2, s3);
    const double m = jl_min4(da, s1, s2, s3);
    if (zb < 0 && (zv < z && z < 0.0)) return -m;
    if (zb > 0 && (0.0 < z && z < zv)) return -m;
    return m;
}

}  // namespace bmo
