// TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// orc_asphere.hpp: CPU restatement of the even-asphere surface SDFs of BeamletOptics.jl
// (src/SDFs/AsphericalLensSDF.jl).  These are pseudo-distances (first-order distance to the sag curve,
// closed by line segments) whose normals always come from central differences (:3-5).
//
// PARITY UNPINNED at the last bit: `r2^i` / `r^(2(m-1)+1)` with a run-time integer exponent go through
// Julia's Base.Math.pow_body (compensated power by squaring, Julia >= 1.8); it is restated here from the
// published algorithm (jl_pow) and cannot be checked against a Julia in this image.  The reference's own
// known-answer tests (test/runtests.jl:1529-1696: surface error <= 1e-10, working distance, focus of the
// three-lens imaging system, ring thicknesses) pin the behaviour at their tolerances (tests/test_asphere.py).
#pragma once
#include <cmath>
#include <vector>
#include "orc_shapes.hpp"

namespace orc {

// Base.Math.pow_body(x::Float64, n::Integer) (julia/base/math.jl): x^n for a run-time integer n.
// muladd is taken as a fused multiply-add (x86-64 / aarch64 builds of Julia lower it to fma).
inline double jl_pow(double x, long n) {
    if (n == 0) return 1.0;
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    if (n == 3) return x * x * x;
    if (n < 0) {
        double rx = 1.0 / x;
        if (n == -2) return rx * rx;
        if (std::isfinite(x)) xnlo = -std::fma(x, rx, -1.0) * rx;
        x = rx;
        n = -n;
    }
    while (n > 1) {
        if (n & 1) {
            double err = std::fma(y, xnlo, x * ynlo);
            double xy = x * y;
            ynlo = std::fma(x, y, -xy);
            y = xy;
            ynlo += err;
        }
        double err = x * 2 * xnlo;
        double xx = x * x;
        xnlo = std::fma(x, x, -xx);
        x = xx;
        xnlo += err;
        n >>= 1;
    }
    double err = std::fma(y, xnlo, x * ynlo);
    return (std::isfinite(x) && std::isfinite(err)) ? std::fma(x, y, err) : x * y;
}

// AsphericalLensSDF.jl:128-141
inline double aspheric_equation(double r, double c, double k, const std::vector<double>& al) {
    double r2 = r * r;
    double sqrt_arg = 1 - (1 + k) * (c * c) * r2;
    if (sqrt_arg < 0) return std::nan("");
    double sum_a = 0.0;
    for (size_t i = 0; i < al.size(); i++) {
        double t = al[i] * jl_pow(r2, (long)i + 1);
        sum_a = (i == 0) ? t : sum_a + t;
    }
    return c * r2 / (1 + std::sqrt(sqrt_arg)) + sum_a;
}
// :147-157 first component of the returned Point2 (the second is 1); NaN when the square root argument is negative
inline double gradient_aspheric_equation(double r, double c, double k, const std::vector<double>& al) {
    double Ri = 1 / c;
    double sqrt_arg = 1 - (r * r) * (1 + k) / (Ri * Ri);
    if (sqrt_arg < 0) return std::nan("");
    double sq = std::sqrt(sqrt_arg);
    double gr = 2 * r / (Ri * (sq + 1)) + (r * r * r) * (1 + k) / ((Ri * Ri * Ri) * sq * ((sq + 1) * (sq + 1)));
    double sum_r = 0.0;
    for (size_t i = 0; i < al.size(); i++) {
        long m = (long)i + 1;
        double t = (double)(2 * m) * al[i] * jl_pow(r, 2 * (m - 1) + 1);
        sum_r = (i == 0) ? t : sum_r + t;
    }
    return -sum_r - gr;
}
// :165-170  distance from p to the segment a-b (2-D)
inline double sd_line_segment(double px, double py, double ax, double ay, double bx, double by) {
    double pax = px - ax, pay = py - ay, bax = bx - ax, bay = by - ay;
    double h = jl_clamp((pax * bax + pay * bay) / (bax * bax + bay * bay), 0.0, 1.0);
    double ex = pax - h * bax, ey = pay - h * bay;
    return std::sqrt(ex * ex + ey * ey);
}
inline double sign_(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }   // Base.sign (keeps the signed zero / NaN)
inline double min3(double a, double b, double c) { return jl_min(jl_min(a, b), c); }
inline double min4(double a, double b, double c, double d) { return jl_min(jl_min(jl_min(a, b), c), d); }

struct AsphParams {
    double c, k, d, max_sag;              // curvature 1/radius, conic constant, diameter, max_sag[1]
    std::vector<double> al;
    double zb, gzb;                       // aspheric_equation(d/2), gradient(d/2)[1] (recomputed by the reference on every call)
};

// :188-240
inline double convex_aspheric_surface_distance(double r, double z, const AsphParams& P) {
    const double c = P.c, d = P.d;
    double r2 = r * r, r2_bound = (d / 2) * (d / 2);
    double zv = aspheric_equation(r, c, P.k, P.al);
    double g = gradient_aspheric_equation(r, c, P.k, P.al);
    double zb = P.zb, n_gzb = std::sqrt(P.gzb * P.gzb + 1.0 * 1.0);
    if (std::isnan(zv) || std::isnan(g) || r2 > r2_bound) {
        double rr = r - sign_(r) * d / 2, dist;
        if (z < zb) dist = std::sqrt(rr * rr + (z - zb) * (z - zb));
        else if (zb < z && z < 0) dist = std::sqrt(rr * rr);
        else if (z > 0 && (sign_(c) == 1 && zb < 0)) dist = std::sqrt(rr * rr + z * z);
        else dist = std::sqrt(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    double da = std::fabs(z - zv) / std::sqrt(g * g + 1.0 * 1.0);
    if (sign_(c) == 1 && zb < 0) {
        double ms = P.max_sag;
        double s1 = sd_line_segment(r, z, d / 2, zb, d / 2, ms) / n_gzb;
        double s2 = sd_line_segment(r, z, d / 2, ms, -d / 2, ms) / n_gzb;
        double s3 = sd_line_segment(r, z, -d / 2, ms, -d / 2, zb) / n_gzb;
        if (zv < z && z < ms) return -min4(da, s1, s2, s3);
        return min4(da, s1, s2, s3);
    }
    double sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
    double sc = sign_(c);
    if (sc * zv < sc * z && sc * z < sc * zb) return -jl_min(sdl, da);
    return jl_min(sdl, da);
}
// :242-307
inline double concave_aspheric_surface_distance(double r, double z, const AsphParams& P) {
    const double c = P.c, d = P.d;
    double r2 = r * r, r2_bound = (d / 2) * (d / 2);
    double zv = aspheric_equation(r, c, P.k, P.al);
    double g = gradient_aspheric_equation(r, c, P.k, P.al);
    double zb = P.zb, n_gzb = std::sqrt(P.gzb * P.gzb + 1.0 * 1.0);
    if (std::isnan(zv) || std::isnan(g)) {
        double rr = r - sign_(r) * d / 2, dist;
        if (z < 0) dist = std::sqrt(rr * rr + z * z);
        else if (0 < z && z < zb) dist = std::sqrt(rr * rr);
        else dist = std::sqrt(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    double da = std::fabs(z - zv) / std::sqrt(g * g + 1.0 * 1.0);
    if (P.max_sag > 0 && zb < 0) {
        double sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        if (r2 > r2_bound) return sdl;
        if (zb < z && z < zv) return -jl_min(da, sdl);
        if (zb > 0 && (0.0 < z && z < zv)) return -jl_min(da, sdl);
        return jl_min(da, sdl);
    }
    double s1 = sd_line_segment(r, z, d / 2, zb, d / 2, 0.0) / n_gzb;
    double s2 = sd_line_segment(r, z, d / 2, 0.0, -d / 2, 0.0) / n_gzb;
    double s3 = sd_line_segment(r, z, -d / 2, 0.0, -d / 2, zb) / n_gzb;
    if (r2 > r2_bound) return min3(s1, s2, s3);
    if (zb < 0 && (zv < z && z < 0.0)) return -min4(da, s1, s2, s3);
    if (zb > 0 && (0.0 < z && z < zv)) return -min4(da, s1, s2, s3);
    return min4(da, s1, s2, s3);
}

// MiscUtils.jl:87-110
template <class F> inline double find_zero_bisection(F f, double a, double b, double tol = 1e-10, int max_iter = 1000) {
    double fa = f(a), fb = f(b);
    if (sign_(fa) == sign_(fb)) throw std::runtime_error("Bisection requires a sign change");
    for (int it = 0; it < max_iter; it++) {
        double mid = (a + b) / 2;
        double fmid = f(mid);
        if (std::fabs(fmid) < tol) return mid;
        if (sign_(fa) == sign_(fmid)) { a = mid; fa = fmid; }
        else { b = mid; fb = fmid; }
    }
    throw std::runtime_error("Bisection did not converge");
}
// AsphericalLensSDF.jl:53-68 -> (f(r_max), r_max)
inline void max_aspheric_value(double c, double k, const std::vector<double>& al, double d, double& fmax, double& rmax) {
    auto f = [&](double r) { return aspheric_equation(r, c, k, al); };
    auto fp = [&](double r) { return gradient_aspheric_equation(r, c, k, al); };
    double a = 1e-8, b = d / 2;
    if (sign_(fp(a)) == sign_(fp(b))) rmax = (std::fabs(f(a)) > std::fabs(f(b))) ? a : b;
    else rmax = find_zero_bisection(fp, a, b);
    fmax = f(rmax);
}

// Convex / ConcaveAsphericalSurfaceSDF (:22-31, :88-98); normal3d = numeric_gradient (:5)
struct AsphSDF : SDF {
    bool convex;
    double radius, diameter;
    AsphParams P;
    double max_sag_r = 0;
    AsphSDF(bool cvx, const std::vector<double>& al, double radius_, double k, double d) : convex(cvx), radius(radius_), diameter(d) {
        P.c = 1 / radius_; P.k = k; P.d = d; P.al = al;
        max_aspheric_value(P.c, k, al, d, P.max_sag, max_sag_r);
        P.zb = aspheric_equation(d / 2, P.c, k, al);
        P.gzb = gradient_aspheric_equation(d / 2, P.c, k, al);
    }
    double edge_sag() const { return aspheric_equation(diameter / 2, 1 / radius, P.k, P.al); }   // :434-443
    double thickness() const override {   // :33-36, :100-103
        double sg = edge_sag();
        if (convex) return (P.max_sag > 0 && sg < 0) ? P.max_sag : std::fabs(sg);
        return (P.max_sag > 0 && sg < 0) ? std::fabs(sg) : 0.0;
    }
    bool has_thickness() const override { return true; }
    double sdf(V3 p) const override {   // :309-349: y is the optical axis, revolve the 2-D distance (AbstractSDF.jl:191-194)
        P3<double> q = w2s(P3<double>{p.x, p.y, p.z});
        double r = std::sqrt(q.x * q.x + q.z * q.z) - 0.0;
        return convex ? convex_aspheric_surface_distance(r, q.y, P) : concave_aspheric_surface_distance(r, q.y, P);
    }
    Dual sdf(P3<Dual>) const override { throw std::runtime_error("aspheric surfaces have no AD path (normal3d = numeric_gradient)"); }
    V3 normal3d(V3 p) const override { return numeric_gradient(p); }
};

}  // namespace orc
