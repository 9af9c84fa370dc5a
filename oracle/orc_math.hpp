// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of BeamletOptics.jl's trace hot path.
// Nothing under oracle/ is linked, imported or executed by the product path (libbmo.so / the
// beamletoptics.jl_b200 package); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, as the checker or the timed CPU baseline.
//
// Parity status: the reference is pure Julia and Julia is not installed in the build container,
// so this restatement is pinned against the reference's own known-answer tests
// (test/runtests.jl, re-expressed in tests/test_oracle_kat.py), not against reference outputs.
// Details the reference's dependencies leave unpinned (ForwardDiff max/min tie rule, zero-vector
// norm rule, StaticArrays normalize) follow the conventions stated next to each function.
//
// orc_math.hpp: 3-vectors, 3x3 matrices, complex numbers and forward-mode dual numbers
// (value + 3 partials) with ForwardDiff.jl's selection rules.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

constexpr double kPi = 3.141592653589793;       // Float64(pi)
constexpr double kTwoPi = 6.283185307179586;     // 2pi in Float64
constexpr double kHalfPi = 1.5707963267948966;   // pi/2
constexpr double kZvac = 376.730313668;          // src/Constants.jl:6
constexpr double kInf = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------------------
// Julia Float64 min/max value semantics (NaN-propagating, signed-zero aware)
inline double jl_max(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    if (x > y) return x;
    if (y > x) return y;
    return std::signbit(x) ? y : x;
}
inline double jl_min(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    if (x < y) return x;
    if (y < x) return y;
    return std::signbit(x) ? x : y;
}
inline double jl_clamp(double x, double lo, double hi) {  // Base.clamp: ifelse(x>hi,hi,ifelse(x<lo,lo,x))
    return x > hi ? hi : (x < lo ? lo : x);
}
// Base.isapprox(x, y; atol, rtol) for finite reals: |x-y| <= max(atol, rtol*max(|x|,|y|))
inline bool jl_isapprox(double x, double y, double atol = 0.0, double rtol = -1.0) {
    if (rtol < 0) rtol = atol > 0 ? 0.0 : 1.4901161193847656e-8;  // sqrt(eps) unless atol given
    if (x == y) return true;
    if (!std::isfinite(x) || !std::isfinite(y)) return false;
    double m = std::fabs(x) > std::fabs(y) ? std::fabs(x) : std::fabs(y);
    double tol = atol > rtol * m ? atol : rtol * m;
    return std::fabs(x - y) <= tol;
}

// ---------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // ((a1b1+a2b2)+a3b3)
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// StaticArrays: normalize(a) = inv(norm(a)) * a   (convention; a / norm(a) differs by <= 1 ulp)
inline V3 normalize(V3 a) { double i = 1.0 / norm(a); return i * a; }

struct M3 {
    double m[3][3];  // m[row][col]
    static M3 identity() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
    V3 col(int c) const { return {m[0][c], m[1][c], m[2][c]}; }
};
inline V3 operator*(const M3& A, V3 v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
            A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
inline M3 transpose(const M3& A) {
    M3 T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T.m[i][j] = A.m[j][i];
    return T;
}

// src/Utils/LinearAlgebraUtils.jl:55-65  rotate3d (Rodrigues), entries exactly as written there
inline M3 rotate3d(V3 u, double theta) {
    double cost = std::cos(theta), sint = std::sin(theta);
    double ux = u.x, uy = u.y, uz = u.z;
    M3 R;
    R.m[0][0] = cost + ux * ux * (1 - cost);
    R.m[0][1] = ux * uy * (1 - cost) - uz * sint;
    R.m[0][2] = ux * uz * (1 - cost) + uy * sint;
    R.m[1][0] = uy * ux * (1 - cost) + uz * sint;
    R.m[1][1] = cost + uy * uy * (1 - cost);
    R.m[1][2] = uy * uz * (1 - cost) - ux * sint;
    R.m[2][0] = uz * ux * (1 - cost) - uy * sint;
    R.m[2][1] = uz * uy * (1 - cost) + ux * sint;
    R.m[2][2] = cost + uz * uz * (1 - cost);
    return R;
}
// src/Utils/LinearAlgebraUtils.jl:74-96  align3d
inline M3 align3d(V3 start, V3 target) {
    start = normalize(start);
    target = normalize(target);
    V3 r = cross(target, start);
    double cosA = dot(start, target);
    if (jl_isapprox(cosA, 1.0)) return M3::identity();
    if (jl_isapprox(cosA, -1.0)) return {{{-1, 0, 0}, {0, -1, 0}, {0, 0, 1}}};
    double k = 1 / (1 + cosA);
    M3 R;
    R.m[0][0] = r.x * r.x * k + cosA; R.m[0][1] = r.x * r.y * k + r.z; R.m[0][2] = r.x * r.z * k - r.y;
    R.m[1][0] = r.y * r.x * k - r.z; R.m[1][1] = r.y * r.y * k + cosA; R.m[1][2] = r.y * r.z * k + r.x;
    R.m[2][0] = r.z * r.x * k + r.y; R.m[2][1] = r.z * r.y * k - r.x; R.m[2][2] = r.z * r.z * k + cosA;
    return R;
}
// src/Utils/LinearAlgebraUtils.jl:103-108  angle3d (clamp is NaN-preserving)
inline double angle3d(V3 target, V3 reference) {
    double arg = jl_clamp(dot(target, reference) / (norm(target) * norm(reference)), -1.0, 1.0);
    return std::acos(arg);
}
// src/Utils/LinearAlgebraUtils.jl:6-8  isparallel3d, atol = eps()
inline bool isparallel3d(V3 a, V3 b) {
    double d = std::fabs(dot(normalize(a), normalize(b)));
    return std::fabs(d - 1.0) <= 2.220446049250313e-16;
}
// src/Utils/LinearAlgebraUtils.jl:127-136  line_plane_distance3d; ok=false <=> `nothing`
inline bool line_plane_distance3d(V3 plane_pos, V3 plane_n, V3 line_pos, V3 line_dir, double& t) {
    double denom = dot(plane_n, line_dir);
    if (std::fabs(denom) > 1e-6) {
        double c = dot(plane_pos - line_pos, plane_n);
        t = c / denom;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Complex numbers with Julia's operation order (no C99 Annex-G NaN recovery)
struct Cx {
    double re, im;
};
inline Cx operator+(Cx a, Cx b) { return {a.re + b.re, a.im + b.im}; }
inline Cx operator-(Cx a, Cx b) { return {a.re - b.re, a.im - b.im}; }
inline Cx operator-(Cx a) { return {-a.re, -a.im}; }
inline Cx operator*(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx operator*(Cx a, double s) { return {a.re * s, a.im * s}; }
inline Cx operator*(double s, Cx a) { return {s * a.re, s * a.im}; }
inline Cx operator/(Cx a, double s) { return {a.re / s, a.im / s}; }
inline Cx operator+(double s, Cx a) { return {s + a.re, a.im}; }
inline Cx operator+(Cx a, double s) { return {a.re + s, a.im}; }
inline Cx operator-(double s, Cx a) { return {s - a.re, -a.im}; }
inline double abs2(Cx a) { return a.re * a.re + a.im * a.im; }
// Smith's algorithm (Julia uses the Baudin-Smith robust variant; equal to <= 2 ulp here)
inline Cx operator/(Cx a, Cx b) {
    if (std::fabs(b.re) >= std::fabs(b.im)) {
        double r = b.im / b.re, d = b.re + b.im * r;
        return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    }
    double r = b.re / b.im, d = b.re * r + b.im;
    return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}
inline Cx csqrt(Cx z) {  // principal branch
    if (z.im == 0.0) {
        if (z.re >= 0) return {std::sqrt(z.re), z.im};
        return {0.0, std::copysign(std::sqrt(-z.re), z.im)};
    }
    double r = std::hypot(z.re, z.im);
    double a = std::sqrt(0.5 * (r + std::fabs(z.re)));
    double b = z.im / (2 * a);
    if (z.re >= 0) return {a, b};
    return {std::fabs(b), std::copysign(a, z.im)};
}
inline Cx cis(double phi) { return {std::cos(phi), std::sin(phi)}; }  // exp(im*phi)

// ---------------------------------------------------------------------------------------------
// Forward-mode dual number with 3 partials, following ForwardDiff.jl (NaN-safe mode OFF):
//   * binary ops combine partials as px*dfdx + py*dfdy, so 0*NaN = NaN propagates
//   * abs(d) = signbit(value) ? -d : d
//   * sqrt: partials * inv(2*sqrt(v))  (v == 0 with zero partials -> Inf*0 = NaN)
//   * max/min follow DiffRules >= 1.3: max selects y iff (y > x) | (signbit(y) < signbit(x))
//   * comparisons look at values only
struct Dual {
    double v;
    double p[3];
};
inline Dual mkdual(double v) { return {v, {0, 0, 0}}; }
inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, {a.p[0] + b.p[0], a.p[1] + b.p[1], a.p[2] + b.p[2]}}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, {a.p[0] - b.p[0], a.p[1] - b.p[1], a.p[2] - b.p[2]}}; }
inline Dual operator+(Dual a, double b) { return {a.v + b, {a.p[0], a.p[1], a.p[2]}}; }
inline Dual operator+(double b, Dual a) { return {b + a.v, {a.p[0], a.p[1], a.p[2]}}; }
inline Dual operator-(Dual a, double b) { return {a.v - b, {a.p[0], a.p[1], a.p[2]}}; }
inline Dual operator-(double b, Dual a) { return {b - a.v, {-a.p[0], -a.p[1], -a.p[2]}}; }
inline Dual operator-(Dual a) { return {-a.v, {-a.p[0], -a.p[1], -a.p[2]}}; }
inline Dual operator*(Dual a, Dual b) {
    return {a.v * b.v, {a.p[0] * b.v + b.p[0] * a.v, a.p[1] * b.v + b.p[1] * a.v, a.p[2] * b.v + b.p[2] * a.v}};
}
inline Dual operator*(Dual a, double b) { return {a.v * b, {a.p[0] * b, a.p[1] * b, a.p[2] * b}}; }
inline Dual operator*(double b, Dual a) { return {b * a.v, {a.p[0] * b, a.p[1] * b, a.p[2] * b}}; }
inline Dual operator/(Dual a, double b) { return {a.v / b, {a.p[0] / b, a.p[1] / b, a.p[2] / b}}; }
inline Dual operator/(Dual a, Dual b) {
    double ib = 1.0 / b.v, c = -(a.v / (b.v * b.v));
    return {a.v / b.v, {a.p[0] * ib + b.p[0] * c, a.p[1] * ib + b.p[1] * c, a.p[2] * ib + b.p[2] * c}};
}
inline bool operator<(Dual a, double b) { return a.v < b; }
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, double b) { return a.v > b; }

inline double value(double a) { return a; }
inline double value(Dual a) { return a.v; }

inline double sqrt_(double a) { return std::sqrt(a); }
inline Dual sqrt_(Dual a) {
    double s = std::sqrt(a.v);
    double d = 1.0 / (2 * s);
    return {s, {a.p[0] * d, a.p[1] * d, a.p[2] * d}};
}
inline double abs_(double a) { return std::fabs(a); }
inline Dual abs_(Dual a) { return std::signbit(a.v) ? -a : a; }

inline double max_(double x, double y) { return jl_max(x, y); }
inline double min_(double x, double y) { return jl_min(x, y); }
// d/dx and d/dy weights of max/min per DiffRules
inline void max_w(double x, double y, double& wx, double& wy) {
    if ((y > x) | (std::signbit(y) < std::signbit(x))) {
        wx = std::isnan(x) ? 1.0 : 0.0;
        wy = std::isnan(x) ? 0.0 : 1.0;
    } else {
        wx = std::isnan(y) ? 0.0 : 1.0;
        wy = std::isnan(y) ? 1.0 : 0.0;
    }
}
inline void min_w(double x, double y, double& wx, double& wy) {
    if ((y < x) | (std::signbit(y) > std::signbit(x))) {
        wx = std::isnan(x) ? 1.0 : 0.0;
        wy = std::isnan(x) ? 0.0 : 1.0;
    } else {
        wx = std::isnan(y) ? 0.0 : 1.0;
        wy = std::isnan(y) ? 1.0 : 0.0;
    }
}
inline Dual max_(Dual a, Dual b) {
    double wx, wy; max_w(a.v, b.v, wx, wy);
    return {jl_max(a.v, b.v), {a.p[0] * wx + b.p[0] * wy, a.p[1] * wx + b.p[1] * wy, a.p[2] * wx + b.p[2] * wy}};
}
inline Dual min_(Dual a, Dual b) {
    double wx, wy; min_w(a.v, b.v, wx, wy);
    return {jl_min(a.v, b.v), {a.p[0] * wx + b.p[0] * wy, a.p[1] * wx + b.p[1] * wy, a.p[2] * wx + b.p[2] * wy}};
}
inline Dual max_(Dual a, double b) {
    double wx, wy; max_w(a.v, b, wx, wy);
    return {jl_max(a.v, b), {a.p[0] * wx, a.p[1] * wx, a.p[2] * wx}};
}
inline Dual min_(Dual a, double b) {
    double wx, wy; min_w(a.v, b, wx, wy);
    return {jl_min(a.v, b), {a.p[0] * wx, a.p[1] * wx, a.p[2] * wx}};
}

// norm of a 2-/3-vector: sqrt(sum of squares).  norm_zero_rule: 1 = ZERO (default; StaticArrays'
// norm falls back to its scaled variant when sqrt(sum abs2) is 0 and returns a clean zero for a zero
// vector of duals), 0 = NAN (plain sqrt(sum abs2): zero vector -> NaN partials -> FD normals).
// The reference's own test pins ZERO: runtests.jl:1309-1314 requires the centre-ray normal at the
// cemented surface of a ROTATED doublet to be parallel to the ray; under the NAN rule that normal
// comes from central differences across the kink where the concave cap touches the plano face and
// is garbage (|n.d| = 0.64), under ZERO the AD gradient is exact.  Values are identical under both.
extern int g_norm_zero_rule;
inline double norm2_(double a, double b) { return std::sqrt(a * a + b * b); }
inline double norm3_(double a, double b, double c) { return std::sqrt(a * a + b * b + c * c); }
inline Dual norm2_(Dual a, Dual b) {
    Dual s = a * a + b * b;
    if (g_norm_zero_rule == 1 && s.v == 0.0) return mkdual(0.0);
    return sqrt_(s);
}
inline Dual norm3_(Dual a, Dual b, Dual c) {
    Dual s = a * a + b * b + c * c;
    if (g_norm_zero_rule == 1 && s.v == 0.0) return mkdual(0.0);
    return sqrt_(s);
}

template <class T> struct P3 { T x, y, z; };
template <class T> struct P2 { T a, b; };

}  // namespace orc
