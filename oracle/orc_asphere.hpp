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

// ---- helpers so that the surface functions can be written once for Float64 and for ForwardDiff duals -------------
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline Dual operator/(double a, Dual b) {   // ForwardDiff: divv = x / v; partials * -(divv / v)
    double q = a / b.v, c = -(q / b.v);
    return {q, {b.p[0] * c, b.p[1] * c, b.p[2] * c}};
}
inline Dual jl_pow(Dual x, long n) {        // d/dx x^n = n x^(n-1) (ForwardDiff's rule for Dual ^ Integer)
    double v = jl_pow(x.v, n), dv = n == 0 ? 0.0 : (double)n * jl_pow(x.v, n - 1);
    return {v, {x.p[0] * dv, x.p[1] * dv, x.p[2] * dv}};
}
inline bool isnan_(double a) { return std::isnan(a); }
inline bool isnan_(Dual a) { return std::isnan(a.v); }
inline double clamp_(double x, double lo, double hi) { return jl_clamp(x, lo, hi); }
inline Dual clamp_(Dual x, double lo, double hi) { return x.v > hi ? mkdual(hi) : (x.v < lo ? mkdual(lo) : x); }
inline double sign_(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }   // Base.sign (keeps the signed zero / NaN)
inline double sign_(Dual x) { return sign_(x.v); }
inline double mkT(double, double c) { return c; }
inline Dual mkT(Dual, double c) { return mkdual(c); }
inline double nan_of(double) { return std::nan(""); }
inline Dual nan_of(Dual) { return mkdual(std::nan("")); }
template <class T> inline T min3(T a, T b, T c) { return min_(min_(a, b), c); }
template <class T> inline T min4(T a, T b, T c, T d) { return min_(min_(min_(a, b), c), d); }

// AsphericalLensSDF.jl:128-141
template <class T> inline T aspheric_equation(T r, double c, double k, const std::vector<double>& al) {
    T r2 = r * r;
    T sqrt_arg = 1 - (1 + k) * (c * c) * r2;
    if (sqrt_arg < 0.0) return nan_of(r);
    T sum_a = r2 * 0.0;
    for (size_t i = 0; i < al.size(); i++) {
        T t = al[i] * jl_pow(r2, (long)i + 1);
        sum_a = (i == 0) ? t : sum_a + t;
    }
    return c * r2 / (1 + sqrt_(sqrt_arg)) + sum_a;
}
// :147-157 first component of the returned Point2 (the second is 1); NaN when the square root argument is negative
template <class T> inline T gradient_aspheric_equation(T r, double c, double k, const std::vector<double>& al) {
    double Ri = 1 / c;
    T sqrt_arg = 1 - (r * r) * (1 + k) / (Ri * Ri);
    if (sqrt_arg < 0.0) return nan_of(r);
    T sq = sqrt_(sqrt_arg);
    T gr = 2 * r / (Ri * (sq + 1)) + (r * r * r) * (1 + k) / ((Ri * Ri * Ri) * sq * ((sq + 1) * (sq + 1)));
    T sum_r = r * 0.0;
    for (size_t i = 0; i < al.size(); i++) {
        long m = (long)i + 1;
        T t = (double)(2 * m) * al[i] * jl_pow(r, 2 * (m - 1) + 1);
        sum_r = (i == 0) ? t : sum_r + t;
    }
    return -sum_r - gr;
}
// :165-170  distance from p to the segment a-b (2-D)
template <class T> inline T sd_line_segment(T px, T py, double ax, double ay, double bx, double by) {
    T pax = px - ax, pay = py - ay;
    double bax = bx - ax, bay = by - ay;
    T h = clamp_((pax * bax + pay * bay) / (bax * bax + bay * bay), 0.0, 1.0);
    T ex = pax - h * bax, ey = pay - h * bay;
    return sqrt_(ex * ex + ey * ey);
}

struct AsphParams {
    double c, k, d, max_sag;              // curvature 1/radius, conic constant, diameter, max_sag[1]
    std::vector<double> al;
    double zb, gzb;                       // aspheric_equation(d/2), gradient(d/2)[1] (recomputed by the reference on every call)
};

// :188-240
template <class T> inline T convex_aspheric_surface_distance(T r, T z, const AsphParams& P) {
    const double c = P.c, d = P.d;
    T r2 = r * r;
    double r2_bound = (d / 2) * (d / 2);
    T zv = aspheric_equation(r, c, P.k, P.al);
    T g = gradient_aspheric_equation(r, c, P.k, P.al);
    double zb = P.zb, n_gzb = std::sqrt(P.gzb * P.gzb + 1.0 * 1.0);
    if (isnan_(zv) || isnan_(g) || r2 > mkT(r, r2_bound)) {
        T rr = r - sign_(r) * d / 2, dist;
        if (z < zb) dist = sqrt_(rr * rr + (z - zb) * (z - zb));
        else if (z > mkT(z, zb) && z < 0.0) dist = sqrt_(rr * rr);
        else if (z > mkT(z, 0.0) && (sign_(c) == 1 && zb < 0)) dist = sqrt_(rr * rr + z * z);
        else dist = sqrt_(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    T da = abs_(z - zv) / sqrt_(g * g + 1.0 * 1.0);
    if (sign_(c) == 1 && zb < 0) {
        double ms = P.max_sag;
        T s1 = sd_line_segment(r, z, d / 2, zb, d / 2, ms) / n_gzb;
        T s2 = sd_line_segment(r, z, d / 2, ms, -d / 2, ms) / n_gzb;
        T s3 = sd_line_segment(r, z, -d / 2, ms, -d / 2, zb) / n_gzb;
        if (zv < z && z < ms) return -min4(da, s1, s2, s3);
        return min4(da, s1, s2, s3);
    }
    T sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
    double sc = sign_(c);
    if (sc * zv < sc * z && sc * z < sc * zb) return -min_(sdl, da);
    return min_(sdl, da);
}
// :242-307
template <class T> inline T concave_aspheric_surface_distance(T r, T z, const AsphParams& P) {
    const double c = P.c, d = P.d;
    T r2 = r * r;
    double r2_bound = (d / 2) * (d / 2);
    T zv = aspheric_equation(r, c, P.k, P.al);
    T g = gradient_aspheric_equation(r, c, P.k, P.al);
    double zb = P.zb, n_gzb = std::sqrt(P.gzb * P.gzb + 1.0 * 1.0);
    (void)c;
    if (isnan_(zv) || isnan_(g)) {
        T rr = r - sign_(r) * d / 2, dist;
        if (z < 0.0) dist = sqrt_(rr * rr + z * z);
        else if (z > mkT(z, 0.0) && z < zb) dist = sqrt_(rr * rr);
        else dist = sqrt_(rr * rr + (z - zb) * (z - zb));
        return dist / n_gzb;
    }
    T da = abs_(z - zv) / sqrt_(g * g + 1.0 * 1.0);
    if (P.max_sag > 0 && zb < 0) {
        T sdl = sd_line_segment(r, z, d / 2, zb, -d / 2, zb) / n_gzb;
        if (r2 > mkT(r, r2_bound)) return sdl;
        if (z > mkT(z, zb) && z < zv) return -min_(da, sdl);
        if (zb > 0 && (z > mkT(z, 0.0) && z < zv)) return -min_(da, sdl);
        return min_(da, sdl);
    }
    T s1 = sd_line_segment(r, z, d / 2, zb, d / 2, 0.0) / n_gzb;
    T s2 = sd_line_segment(r, z, d / 2, 0.0, -d / 2, 0.0) / n_gzb;
    T s3 = sd_line_segment(r, z, -d / 2, 0.0, -d / 2, zb) / n_gzb;
    if (r2 > mkT(r, r2_bound)) return min3(s1, s2, s3);
    if (zb < 0 && (zv < z && z < 0.0)) return -min4(da, s1, s2, s3);
    if (zb > 0 && (z > mkT(z, 0.0) && z < zv)) return -min4(da, s1, s2, s3);
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

// Aconvex / AconcaveCylinderSDF (AcylindricalSDF.jl:14-113): the aspheric profile in (z, y), extruded along x
// (AbstractSDF.jl:229-234).  Normals: the generic normal_fd (ForwardDiff, finite-difference fallback).
struct AcylSDF : SDF {
    bool convex;
    double radius, diameter, height;
    AsphParams P;
    double max_sag_r = 0;
    AcylSDF(bool cvx, double radius_, double d, double h, double k, const std::vector<double>& al)
        : convex(cvx), radius(radius_), diameter(d), height(h) {
        P.c = 1 / radius_; P.k = k; P.d = d; P.al = al;
        max_aspheric_value(P.c, k, al, d, P.max_sag, max_sag_r);
        P.zb = aspheric_equation(d / 2, P.c, k, al);
        P.gzb = gradient_aspheric_equation(d / 2, P.c, k, al);
    }
    double edge_sag() const { return aspheric_equation(diameter / 2, 1 / radius, P.k, P.al); }   // AcylindricalSDF.jl:170-178
    double thickness() const override {   // :52-54, :97-100
        double sg = edge_sag();
        if (convex) return std::fabs(sg);
        return (P.max_sag > 0 && sg < 0) ? std::fabs(sg) : 0.0;
    }
    bool has_thickness() const override { return true; }
    template <class T> T eval(P3<T> q) const {
        P3<T> p = w2s(q);
        T d2 = convex ? convex_aspheric_surface_distance(p.z, p.y, P) : concave_aspheric_surface_distance(p.z, p.y, P);
        return cyl_(d2, abs_(p.x) - height / 2);
    }
    double sdf(V3 p) const override { return eval(P3<double>{p.x, p.y, p.z}); }
    Dual sdf(P3<Dual> p) const override { return eval(p); }
};

}  // namespace orc
