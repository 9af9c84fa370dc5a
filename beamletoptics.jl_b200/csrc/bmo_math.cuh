// bmo_math.cuh -- FP64 vector / complex / dual-number arithmetic for the sm_100a tracer.
//
// Everything here is compiled with -fmad=false: the reference (Julia) never contracts a*b+c, and
// the hit points must agree to 1e-9 relative even where the reference's finite-difference normals
// (eps = 1e-8, src/SDFs/AbstractSDF.jl:81-88) amplify rounding noise by 1e8.  Operation order
// follows the reference expressions cited at each call site.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bmo {

#define BMO_HD __host__ __device__ __forceinline__
#define BMO_D __device__ __forceinline__
#define BMO_NI static __device__ __noinline__

constexpr double kPi = 3.141592653589793;
constexpr double kTwoPi = 6.283185307179586;
constexpr double kHalfPi = 1.5707963267948966;
constexpr double kZvac = 376.730313668;  // src/Constants.jl:6

struct V3 { double x, y, z; };
BMO_HD V3 mk3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
BMO_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
BMO_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
BMO_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
BMO_HD V3 operator*(double s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
BMO_HD V3 operator*(V3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
BMO_HD V3 operator/(V3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
BMO_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
BMO_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
BMO_HD double norm(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
BMO_HD V3 normalize(V3 a) { double i = 1.0 / norm(a); return i * a; }  // inv(norm(a)) * a
BMO_HD double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Julia Float64 max/min: NaN-propagating (CUDA fmax/fmin drop NaNs), signed-zero aware
BMO_HD double jl_max(double x, double y) {
    if (isnan(x) || isnan(y)) return x + y;
    if (x > y) return x;
    if (y > x) return y;
    return signbit(x) ? y : x;
}
BMO_HD double jl_min(double x, double y) {
    if (isnan(x) || isnan(y)) return x + y;
    if (x < y) return x;
    if (y < x) return y;
    return signbit(x) ? x : y;
}
BMO_HD double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
BMO_HD bool jl_isapprox(double x, double y) {  // default rtol = sqrt(eps)
    if (x == y) return true;
    if (!isfinite(x) || !isfinite(y)) return false;
    double m = fabs(x) > fabs(y) ? fabs(x) : fabs(y);
    return fabs(x - y) <= 1.4901161193847656e-8 * m;
}

// ---- complex ---------------------------------------------------------------------------------
struct Cx { double re, im; };
BMO_HD Cx mkc(double re, double im) { Cx c; c.re = re; c.im = im; return c; }
BMO_HD Cx operator+(Cx a, Cx b) { return mkc(a.re + b.re, a.im + b.im); }
BMO_HD Cx operator-(Cx a, Cx b) { return mkc(a.re - b.re, a.im - b.im); }
BMO_HD Cx operator-(Cx a) { return mkc(-a.re, -a.im); }
BMO_HD Cx operator*(Cx a, Cx b) { return mkc(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
BMO_HD Cx operator*(Cx a, double s) { return mkc(a.re * s, a.im * s); }
BMO_HD Cx operator*(double s, Cx a) { return mkc(s * a.re, s * a.im); }
BMO_HD Cx operator/(Cx a, double s) { return mkc(a.re / s, a.im / s); }
BMO_HD Cx operator+(double s, Cx a) { return mkc(s + a.re, a.im); }
BMO_HD Cx operator+(Cx a, double s) { return mkc(a.re + s, a.im); }
BMO_HD Cx operator-(double s, Cx a) { return mkc(s - a.re, -a.im); }
BMO_HD double abs2(Cx a) { return a.re * a.re + a.im * a.im; }
BMO_HD Cx operator/(Cx a, Cx b) {  // Smith
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, d = b.re + b.im * r;
        return mkc((a.re + a.im * r) / d, (a.im - a.re * r) / d);
    }
    double r = b.re / b.im, d = b.re * r + b.im;
    return mkc((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
BMO_HD Cx csqrt_(Cx z) {
    if (z.im == 0.0) {
        if (z.re >= 0) return mkc(sqrt(z.re), z.im);
        return mkc(0.0, copysign(sqrt(-z.re), z.im));
    }
    double r = hypot(z.re, z.im);
    double a = sqrt(0.5 * (r + fabs(z.re)));
    double b = z.im / (2 * a);
    if (z.re >= 0) return mkc(a, b);
    return mkc(fabs(b), copysign(a, z.im));
}
BMO_D Cx cis(double phi) { double s, c; sincos(phi, &s, &c); return mkc(c, s); }

// ---- dual numbers (value + 3 partials) with ForwardDiff.jl's rules ---------------------------
// binary ops combine partials as px*wx + py*wy (0*NaN = NaN propagates: NaN-safe mode is off),
// abs -> sign flip, sqrt(0) -> Inf*0 = NaN partials, max/min per DiffRules (tie keeps x unless the
// signbits differ), comparisons on values only.  These decide whether normal3d takes the AD or the
// finite-difference branch (src/SDFs/AbstractSDF.jl:90-95).
struct Dual { double v, p0, p1, p2; };
BMO_HD Dual mkd(double v, double a, double b, double c) { Dual d; d.v = v; d.p0 = a; d.p1 = b; d.p2 = c; return d; }
BMO_HD Dual operator+(Dual a, Dual b) { return mkd(a.v + b.v, a.p0 + b.p0, a.p1 + b.p1, a.p2 + b.p2); }
BMO_HD Dual operator-(Dual a, Dual b) { return mkd(a.v - b.v, a.p0 - b.p0, a.p1 - b.p1, a.p2 - b.p2); }
BMO_HD Dual operator+(Dual a, double b) { return mkd(a.v + b, a.p0, a.p1, a.p2); }
BMO_HD Dual operator+(double b, Dual a) { return mkd(b + a.v, a.p0, a.p1, a.p2); }
BMO_HD Dual operator-(Dual a, double b) { return mkd(a.v - b, a.p0, a.p1, a.p2); }
BMO_HD Dual operator-(double b, Dual a) { return mkd(b - a.v, -a.p0, -a.p1, -a.p2); }
BMO_HD Dual operator-(Dual a) { return mkd(-a.v, -a.p0, -a.p1, -a.p2); }
BMO_HD Dual operator*(Dual a, Dual b) {
    return mkd(a.v * b.v, a.p0 * b.v + b.p0 * a.v, a.p1 * b.v + b.p1 * a.v, a.p2 * b.v + b.p2 * a.v);
}
BMO_HD Dual operator*(Dual a, double b) { return mkd(a.v * b, a.p0 * b, a.p1 * b, a.p2 * b); }
BMO_HD Dual operator*(double b, Dual a) { return mkd(b * a.v, a.p0 * b, a.p1 * b, a.p2 * b); }
BMO_HD Dual operator/(Dual a, double b) { return mkd(a.v / b, a.p0 / b, a.p1 / b, a.p2 / b); }
BMO_HD bool operator<(Dual a, double b) { return a.v < b; }
BMO_HD bool operator<(Dual a, Dual b) { return a.v < b.v; }

BMO_HD double value(double a) { return a; }
BMO_HD double value(Dual a) { return a.v; }
BMO_HD double sqrt_(double a) { return sqrt(a); }
BMO_HD Dual sqrt_(Dual a) {
    double s = sqrt(a.v);
    double d = 1.0 / (2 * s);
    return mkd(s, a.p0 * d, a.p1 * d, a.p2 * d);
}
BMO_HD double abs_(double a) { return fabs(a); }
BMO_HD Dual abs_(Dual a) { return signbit(a.v) ? -a : a; }
BMO_HD double max_(double x, double y) { return jl_max(x, y); }
BMO_HD double min_(double x, double y) { return jl_min(x, y); }
BMO_HD void max_w(double x, double y, double& wx, double& wy) {
    if ((y > x) | ((int)signbit(y) < (int)signbit(x))) { wx = isnan(x) ? 1.0 : 0.0; wy = isnan(x) ? 0.0 : 1.0; }
    else { wx = isnan(y) ? 0.0 : 1.0; wy = isnan(y) ? 1.0 : 0.0; }
}
BMO_HD void min_w(double x, double y, double& wx, double& wy) {
    if ((y < x) | ((int)signbit(y) > (int)signbit(x))) { wx = isnan(x) ? 1.0 : 0.0; wy = isnan(x) ? 0.0 : 1.0; }
    else { wx = isnan(y) ? 0.0 : 1.0; wy = isnan(y) ? 1.0 : 0.0; }
}
BMO_HD Dual max_(Dual a, Dual b) {
    double wx, wy; max_w(a.v, b.v, wx, wy);
    return mkd(jl_max(a.v, b.v), a.p0 * wx + b.p0 * wy, a.p1 * wx + b.p1 * wy, a.p2 * wx + b.p2 * wy);
}
BMO_HD Dual min_(Dual a, Dual b) {
    double wx, wy; min_w(a.v, b.v, wx, wy);
    return mkd(jl_min(a.v, b.v), a.p0 * wx + b.p0 * wy, a.p1 * wx + b.p1 * wy, a.p2 * wx + b.p2 * wy);
}
BMO_HD Dual max_(Dual a, double b) {
    double wx, wy; max_w(a.v, b, wx, wy);
    return mkd(jl_max(a.v, b), a.p0 * wx, a.p1 * wx, a.p2 * wx);
}
BMO_HD Dual min_(Dual a, double b) {
    double wx, wy; min_w(a.v, b, wx, wy);
    return mkd(jl_min(a.v, b), a.p0 * wx, a.p1 * wx, a.p2 * wx);
}
// norm of small vectors: sqrt(sum abs2); `zr` = norm_zero_rule (1: zero vector -> clean zero dual)
BMO_HD double norm2_(double a, double b, int) { return sqrt(a * a + b * b); }
BMO_HD double norm3_(double a, double b, double c, int) { return sqrt(a * a + b * b + c * c); }
BMO_HD Dual norm2_(Dual a, Dual b, int zr) {
    Dual s = a * a + b * b;
    if (zr == 1 && s.v == 0.0) return mkd(0, 0, 0, 0);
    return sqrt_(s);
}
BMO_HD Dual norm3_(Dual a, Dual b, Dual c, int zr) {
    Dual s = a * a + b * b + c * c;
    if (zr == 1 && s.v == 0.0) return mkd(0, 0, 0, 0);
    return sqrt_(s);
}

// sincos of a large phase (k z is O(1e7) rad; CUDA's sincos leaves its fast path at |x| > 105615):
// Cody-Waite reduction by pi/2 with two FMAs (n < 2^31, residual error n * 1.5e-33), then the
// fdlibm minimax kernels on [-pi/4, pi/4] (error < 1 ulp) and the quadrant swap.
// The coefficients sit in constant memory so that they enter the FMAs as constant-bank operands: as immediates every one of
// them costs two uniform-register moves per evaluation (48 of the 292 instructions of the detector kernel's pair loop).
static __constant__ double kTrigC[16] = {
    0.6366197723675814, -1.5707963267948966, -6.123233995736766e-17,                                    // 2/pi, -pi/2 hi, -pi/2 lo
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,                // sin: S6 .. S1
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,               // cos: C6 .. C1
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, 0.0};
BMO_D void sincos_reduced(double x, double* sn, double* cs) {
    const double n = rint(x * kTrigC[0]);
    double r = fma(n, kTrigC[1], x);
    r = fma(n, kTrigC[2], r);
    const double z = r * r;
    double ps = fma(z, kTrigC[3], kTrigC[4]);
    ps = fma(z, ps, kTrigC[5]);
    ps = fma(z, ps, kTrigC[6]);
    ps = fma(z, ps, kTrigC[7]);
    ps = fma(z, ps, kTrigC[8]);
    const double s0 = fma(z * r, ps, r);
    double pc = fma(z, kTrigC[9], kTrigC[10]);
    pc = fma(z, pc, kTrigC[11]);
    pc = fma(z, pc, kTrigC[12]);
    pc = fma(z, pc, kTrigC[13]);
    pc = fma(z, pc, kTrigC[14]);
    const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int q = (int)(long long)n;
    const double sv = (q & 1) ? c0 : s0, cv = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -sv : sv;
    *cs = ((q + 1) & 2) ? -cv : cv;
}
// exp(x) for x <= 0 (Gaussian envelopes): 2^n * P(r), r = x - n ln2 in [-ln2/2, ln2/2], Taylor polynomial of degree 11 in
// Horner form (truncation 0.35^12 / 12! = 7e-15), coefficients from constant memory; 0 below -708 (the library returns a subnormal)
static __constant__ double kExpC[16] = {
    1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,                       // log2(e), -ln2 hi, -ln2 lo
    2.505210838544172e-08, 2.755731922398589e-07, 2.755731922398589e-06, 2.48015873015873e-05,          // 1/11! .. 1/8!
    1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.166666666666666e-02,         // 1/7! .. 1/4!
    1.666666666666667e-01, 0.5, 0.0, 0.0};
BMO_D double exp_neg(double x) {
    const double n = rint(x * kExpC[0]);
    double r = fma(n, kExpC[1], x);
    r = fma(n, kExpC[2], r);
    double p = fma(r, kExpC[3], kExpC[4]);
    p = fma(r, p, kExpC[5]);
    p = fma(r, p, kExpC[6]);
    p = fma(r, p, kExpC[7]);
    p = fma(r, p, kExpC[8]);
    p = fma(r, p, kExpC[9]);
    p = fma(r, p, kExpC[10]);
    p = fma(r, p, kExpC[11]);
    p = fma(r, p, kExpC[12]);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    // scale by 2^n through the exponent field (n in [-1022, 0] here, p in [0.7, 1.42])
    const int hi = __double2hiint(p) + ((int)n << 20);
    const double v = __hiloint2double(hi, __double2loint(p));
    return x < -708.0 ? 0.0 : (x != x ? x : v);
}

template <class T> struct P3 { T x, y, z; };

}  // namespace bmo
