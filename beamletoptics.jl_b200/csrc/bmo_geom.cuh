// bmo_geom.cuh -- device-side geometry of the tracer: SDF evaluation (value and dual), sphere
// tracing, Moeller-Trumbore behind a BVH, per-object closest hit, trace_all / trace_one.
// Citations are relative to /root/reference/src.
#pragma once
#include "../../include/bmo.h"
#include "bmo_math.cuh"
#include "bmo_asphere.cuh"
#include <type_traits>

namespace bmo {

constexpr double eps_srf = 1e-9;   // SDFs/AbstractSDF.jl:1
constexpr double eps_ray = 1e-10;  // :2
constexpr double eps_ins = 1.0;    // :3
constexpr int kMarchIter = 1000;   // :105,135
constexpr int NBOUND = 10;         // doubles per part bound: sphere centre xyz + radius, box lo xyz + hi xyz

struct BvhNode {
    double lo[3], hi[3];
    int32_t left, right;   // interior: child node indices (relative to the mesh's first node)
    int32_t first, count;  // leaf (count > 0): range in bvh_faces
};
struct MeshView {
    int64_t first_vertex, n_vertices, first_face, n_faces;
    int64_t first_node, n_nodes;  // n_nodes == 0: brute force in face order
    int32_t f32, pad;
};

// Device view of one flattened system (all pointers are device pointers)
struct SysView {
    const bmo_prim* prims;
    const bmo_part* parts;
    const bmo_object* objects;
    const MeshView* meshes;
    const double* vertices;
    const int32_t* faces;
    const BvhNode* nodes;
    const int32_t* bvh_faces;
    const double* n_table;
    const double* bounds;    // [n_poses][n_parts][NBOUND]: sphere (centre, radius), box (lo, hi)
    const double* det_pose;  // [n_poses][n_objects][12] pos(3) dir(9 row-major)
    const double* lambdas;   // [n_lambda]
    const double* jones;     // [n_jones][10] PolarizationFilter: J (3x3 row-major), cutoff
    const double* ext;       // parameter blocks of the aspheric primitives (bmo_asphere.cuh)
    int32_t n_prims, n_parts, n_objects, n_meshes, n_lambda, n_poses, zr;
    int32_t bvh_ok;          // 1: the vertex table is the one the BVH was built from (upload / restore); 0 after a pose update
    int64_t n_vertices;
    double n_system;
};

struct Stats { unsigned sdf, tri; };

struct Hit {
    double t;
    V3 n;
    int32_t part;  // -1: miss
};

// ---- SDF primitives ---------------------------------------------------------------------------
// world -> local:  T * (point - pos)   (SDFs/AbstractSDF.jl:35-40)
template <class T> BMO_D P3<T> w2s(const bmo_prim& pr, P3<T> q) {
    T dx = q.x - pr.pos[0], dy = q.y - pr.pos[1], dz = q.z - pr.pos[2];
    P3<T> p;
    p.x = pr.tdir[0] * dx + pr.tdir[1] * dy + pr.tdir[2] * dz;
    p.y = pr.tdir[3] * dx + pr.tdir[4] * dy + pr.tdir[5] * dz;
    p.z = pr.tdir[6] * dx + pr.tdir[7] * dy + pr.tdir[8] * dz;
    return p;
}
// The same transform of a point that carries ForwardDiff's seed partials (dq/dq = I), as member_normal passes it:
// the generic dual arithmetic multiplies every tdir entry by the seeds 1 / 0 and adds the zeros, which leaves exactly
// the rows of tdir (up to the sign of an exact zero, see the note at the fast double path below) -- written down directly.
BMO_D P3<double> w2s_seeded(const bmo_prim& pr, P3<double> q) { return w2s(pr, q); }
BMO_D P3<Dual> w2s_seeded(const bmo_prim& pr, P3<Dual> q) {
    const double dx = q.x.v - pr.pos[0], dy = q.y.v - pr.pos[1], dz = q.z.v - pr.pos[2];
    P3<Dual> p;
    p.x = mkd(pr.tdir[0] * dx + pr.tdir[1] * dy + pr.tdir[2] * dz, pr.tdir[0], pr.tdir[1], pr.tdir[2]);
    p.y = mkd(pr.tdir[3] * dx + pr.tdir[4] * dy + pr.tdir[5] * dz, pr.tdir[3], pr.tdir[4], pr.tdir[5]);
    p.z = mkd(pr.tdir[6] * dx + pr.tdir[7] * dy + pr.tdir[8] * dz, pr.tdir[6], pr.tdir[7], pr.tdir[8]);
    return p;
}
// min(maximum(d), 0) + norm(max.(d, 0))   (SphericalLensSDF.jl:64)
template <class T> BMO_D T cyl_(T d1, T d2, int zr) {
    return min_(max_(d1, d2), 0.0) + norm2_(max_(d1, 0.0), max_(d2, 0.0), zr);
}
// Aspheric primitives: bmo_system_upload stores the device address of the primitive's parameter block
// (SysView::ext + ext_first) in par[0], bit for bit, so that every evaluation site reaches it through the record.
BMO_D const double* asph_block(const bmo_prim& pr) { return reinterpret_cast<const double*>((unsigned long long)__double_as_longlong(pr.par[0])); }
BMO_D bool is_asph(int type) { return type == BMO_PRIM_CONVEX_ASPH || type == BMO_PRIM_CONCAVE_ASPH; }
BMO_D bool is_acyl(int type) { return type == BMO_PRIM_CONVEX_ACYL || type == BMO_PRIM_CONCAVE_ACYL; }
// Cylindrical and aspheric surfaces, generic (value / dual) evaluation: out of line for the same reason as
// prim_eval_rare below -- a noinline function saves every callee-saved register any of its cases needs.
template <class T> BMO_NI T prim_eval_rare_t(const bmo_prim& pr, P3<T> p, int zr) {
    const double a = pr.par[0], b = pr.par[1], c = pr.par[2], d = pr.par[3];
    switch (pr.type) {
        case BMO_PRIM_CONVEX_ASPH:
        case BMO_PRIM_CONCAVE_ASPH:   // AsphericalLensSDF.jl:309-349: y is the optical axis, the 2-D distance is revolved about it
            if constexpr (std::is_same<T, double>::value)
                return aspheric_surface_distance<double>(pr.type == BMO_PRIM_CONVEX_ASPH, sqrt(p.x * p.x + p.z * p.z) - 0.0, p.y, asph_block(pr));
            else
                return T{};   // no AD path: normal3d = numeric_gradient (:5)
        case BMO_PRIM_CONVEX_ACYL:
        case BMO_PRIM_CONCAVE_ACYL: {  // AcylindricalSDF.jl:55-72, 101-120: profile over (z, y), op_extrude_x (AbstractSDF.jl:229-234)
            T d2 = aspheric_surface_distance<T>(pr.type == BMO_PRIM_CONVEX_ACYL, p.z, p.y, asph_block(pr));
            return cyl_(d2, abs_(p.x) - b, zr);
        }
        case BMO_PRIM_CONVEX_CYL: {  // CylindricalSDF.jl:59-78: sdf_cut_disk in (y, z), op_extrude_x (AbstractSDF.jl:229-234)
            const double r = a, h = b, w = c, hx = d;
            T p1 = abs_(p.y), p2 = p.z;
            T s = max_((h - r) * (p1 * p1) + (w * w) * (h + r - 2 * p2), h * p1 - w * p2);
            T dd;
            if (s < 0.0) dd = norm2_(p1, p2, zr) - r;
            else if (p1 < w) dd = h - p2;
            else dd = norm2_(p1 - w, p2 - h, zr);
            return cyl_(dd, abs_(p.x) - hx, zr);
        }
        case BMO_PRIM_CONCAVE_CYL: {  // CylindricalSDF.jl:120-133
            const double r = a, sg = d;
            T psy = p.y + (-r);
            T d1 = abs_(norm2_(p.z, psy, zr)) - fabs(r);
            T d2 = abs_(p.x) - c / 2;
            T cc = cyl_(d1, d2, zr);
            T ppy = p.y + (-sg / 2 * (r > 0 ? 1.0 : (r < 0 ? -1.0 : 0.0)));
            T qx = abs_(p.x) - c / 2, qy = abs_(ppy) - sg / 2, qz = abs_(p.z) - b / 2;
            T l = norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0), zr) + min_(max_(qx, max_(qy, qz)), 0.0);
            return max_(l, -cc);
        }
        default: break;
    }
    return T{};
}
// RK: the kernel was compiled for systems with cylindrical / aspheric primitives (chosen by the host per system)
// seeded: q carries the identity partials (only meaningful for T = Dual)
template <class T, bool RK> BMO_NI T prim_eval(const bmo_prim& pr, P3<T> q, int zr, Stats& st, bool seeded = false) {
    st.sdf++;
    P3<T> p = seeded ? w2s_seeded(pr, q) : w2s(pr, q);
    if (RK) { if (pr.type >= BMO_PRIM_CONVEX_CYL) return prim_eval_rare_t<T>(pr, p, zr); }
    const double a = pr.par[0], b = pr.par[1], c = pr.par[2], d = pr.par[3];
    switch (pr.type) {
        case BMO_PRIM_PLANO: {  // SphericalLensSDF.jl:60-65
            T d1 = abs_(norm2_(p.x, p.z, zr)) - b / 2;
            T d2 = abs_(p.y - a / 2) - a / 2;
            return cyl_(d1, d2, zr);
        }
        case BMO_PRIM_CYLINDER: {  // PrimitiveSDF.jl:71-76
            T d1 = abs_(norm2_(p.x, p.z, zr)) - a;
            T d2 = abs_(p.y) - b;
            return cyl_(d1, d2, zr);
        }
        case BMO_PRIM_SPHERE:  // SphericalLensSDF.jl:86-89
            return norm3_(p.x, p.y, p.z, zr) - a;
        case BMO_PRIM_CONVEX: {  // SphericalLensSDF.jl:219-232
            T q1 = norm2_(p.x, p.z, zr);
            T q2 = -p.y + a;
            const double h = d, R = a, hd = b / 2;
            T s = max_((h - R) * (q1 * q1) + (hd * hd) * (h + R - 2 * q2), h * q1 - hd * q2);
            if (s < 0.0) return norm2_(q1, q2, zr) - R;
            if (q1 < hd) return h - q2;
            return norm2_(q1 - hd, q2 - h, zr);
        }
        case BMO_PRIM_CONCAVE: {  // SphericalLensSDF.jl:159-170
            T psy = p.y + c / 2;
            T d1 = abs_(norm2_(p.x, p.z, zr)) - b / 2;
            T d2 = abs_(psy) - c / 2;
            T sdf1 = cyl_(d1, d2, zr);
            T sdf2 = norm3_(p.x, p.y + a, p.z, zr) - a;
            return max_(sdf1, -sdf2);
        }
        case BMO_PRIM_CUTSPHERE: {  // PrimitiveSDF.jl:112-124
            T q1 = norm2_(p.x, p.z, zr);
            T q2 = p.y;
            const double h = b, R = a, w = c;
            T s = max_((h - R) * (q1 * q1) + (w * w) * (h + R - 2 * q2), h * q1 - w * q2);
            if (s < 0.0) return norm2_(q1, q2, zr) - R;
            if (q1 < w) return h - q2;
            return norm2_(q1 - w, q2 - h, zr);
        }
        case BMO_PRIM_BOX: {  // PrimitiveSDF.jl:41-46
            T qx = abs_(p.x) - a, qy = abs_(p.y) - b, qz = abs_(p.z) - c;
            return norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0), zr) + min_(max_(qx, max_(qy, qz)), 0.0);
        }
        case BMO_PRIM_RING: {  // PrimitiveSDF.jl:157-166
            T px = norm2_(p.x, p.z, zr) - a;
            T d1 = abs_(px) - b, d2 = abs_(p.y) - c;
            return norm2_(max_(d1, 0.0), max_(d2, 0.0), zr) + min_(max_(d1, d2), 0.0);
        }
        case BMO_PRIM_RAPRISM: {  // PrimitiveSDF.jl:204-210
            T qx = abs_(p.x) - a, qy = abs_(p.y) - b, qz = abs_(p.z) - c;
            T box = norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0), zr) + min_(max_(qx, max_(qy, qz)), 0.0);
            T pln = (p.x + p.y) / 1.4142135623730951;  // sqrt(2)
            return max_(box, pln);
        }
        default: break;
    }
    return T{};
}
// one member of a union: a primitive, or a meniscus frame + 3 children (MeniscusLensSDF.jl:42-46)
template <class T, bool RK> BMO_D T member_eval(const bmo_prim* prims, int i, P3<T> q, int zr, Stats& st, bool seeded = false) {
    const bmo_prim& pr = prims[i];
    if (pr.type != BMO_PRIM_MENISCUS) return prim_eval<T, RK>(pr, q, zr, st, seeded);
    P3<T> p = seeded ? w2s_seeded(pr, q) : w2s(pr, q);
    T cv = prim_eval<T, RK>(prims[i + 1], p, zr, st);
    T cy = prim_eval<T, RK>(prims[i + 2], p, zr, st);
    T cc = prim_eval<T, RK>(prims[i + 3], p, zr, st);
    return max_(min_(cv, cy), -cc);
}
BMO_D int member_advance(const bmo_prim* prims, int i) { return prims[i].type == BMO_PRIM_MENISCUS ? 4 : 1; }

// ---- fast double path ----------------------------------------------------------------------------
// The marching loop evaluates the SDF ~30x per interaction, so the plain-double evaluation gets a
// lean restatement.  For finite arguments it returns bit-identical values to prim_eval<double>
// above (the NaN-propagating branches of Julia's min/max cannot trigger, sqrt(x*x) == |x| is exact in
// binary64 when x*x neither under- nor overflows); only the sign of an exactly-zero intermediate
// may differ, which no comparison or non-zero value downstream can observe.  The dual-number
// (ForwardDiff) and finite-difference evaluations of the normals keep the generic templates.
BMO_D double fmax_jl(double x, double y) { const double d = x - y; return signbit(d) ? y : x; }  // Base.max, finite args
BMO_D double fmin_jl(double x, double y) { const double d = x - y; return signbit(d) ? x : y; }  // Base.min, finite args
BMO_D double pos_part(double x) { return signbit(x) ? 0.0 : x; }                                   // max(x, 0.0)
BMO_D double neg_part(double x) { return signbit(x) ? x : 0.0; }                                   // min(x, 0.0)
BMO_D double hyp2(double a, double b) { return sqrt(a * a + b * b); }
BMO_D bool sq_exact(double a) {   // 2^-463 <= a < 2^465: a*a is a normal number, so sqrt(fl(a*a)) == a
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    return e - 560u < 928u;
}
// norm(max.(d, 0)) of a 2-vector / 3-vector
BMO_D double pnorm2(double d1, double d2) {
    const double a = pos_part(d1), b = pos_part(d2);
    if (b == 0.0) { if (a == 0.0 || sq_exact(a)) return a; }
    else if (a == 0.0 && sq_exact(b)) return b;
    return sqrt(a * a + b * b);
}
BMO_D double pnorm3(double d1, double d2, double d3) {
    const double a = pos_part(d1), b = pos_part(d2), c = pos_part(d3);
    if (a == 0.0 && b == 0.0 && (c == 0.0 || sq_exact(c))) return c;
    if (a == 0.0 && c == 0.0 && sq_exact(b)) return b;
    if (b == 0.0 && c == 0.0 && sq_exact(a)) return a;
    return sqrt(a * a + b * b + c * c);
}
BMO_D V3 w2s_f(const bmo_prim& pr, V3 q) {
    const double dx = q.x - pr.pos[0], dy = q.y - pr.pos[1], dz = q.z - pr.pos[2];
    if (pr.reserved & 1) return mk3(dx, dy, dz);   // transposed_dir == I (set at upload)
    return mk3(pr.tdir[0] * dx + pr.tdir[1] * dy + pr.tdir[2] * dz,
               pr.tdir[3] * dx + pr.tdir[4] * dy + pr.tdir[5] * dz,
               pr.tdir[6] * dx + pr.tdir[7] * dy + pr.tdir[8] * dz);
}
BMO_D double cyl_f(double d1, double d2) { return neg_part(fmax_jl(d1, d2)) + pnorm2(d1, d2); }
// Less common primitives (cylindrical and aspheric lens surfaces) are kept out of line so that the marching
// loop of the common spherical-lens path keeps its register budget.
BMO_NI double prim_eval_rare(const bmo_prim& pr, V3 p) {
    const double a = pr.par[0], b = pr.par[1], c = pr.par[2], d = pr.par[3];
    switch (pr.type) {
        case BMO_PRIM_CONVEX_ASPH:
        case BMO_PRIM_CONCAVE_ASPH:
            return aspheric_surface_distance<double>(pr.type == BMO_PRIM_CONVEX_ASPH, sqrt(p.x * p.x + p.z * p.z) - 0.0, p.y, asph_block(pr));
        case BMO_PRIM_CONVEX_ACYL:
        case BMO_PRIM_CONCAVE_ACYL:
            return cyl_f(aspheric_surface_distance<double>(pr.type == BMO_PRIM_CONVEX_ACYL, p.z, p.y, asph_block(pr)), fabs(p.x) - b);
        case BMO_PRIM_CONVEX_CYL: {
            const double r = a, h = b, w = c, hx = d;
            const double p1 = fabs(p.y), p2 = p.z;
            const double s = fmax_jl((h - r) * (p1 * p1) + (w * w) * (h + r - 2 * p2), h * p1 - w * p2);
            double dd;
            if (s < 0.0) dd = hyp2(p1, p2) - r;
            else if (p1 < w) dd = h - p2;
            else dd = hyp2(p1 - w, p2 - h);
            return cyl_f(dd, fabs(p.x) - hx);
        }
        case BMO_PRIM_CONCAVE_CYL: {
            const double r = a, sg = d;
            const double psy = p.y + (-r);
            const double cc = cyl_f(hyp2(p.z, psy) - fabs(r), fabs(p.x) - c / 2);
            const double ppy = p.y + (-sg / 2 * (r > 0 ? 1.0 : (r < 0 ? -1.0 : 0.0)));
            const double qx = fabs(p.x) - c / 2, qy = fabs(ppy) - sg / 2, qz = fabs(p.z) - b / 2;
            const double l = pnorm3(qx, qy, qz) + neg_part(fmax_jl(qx, fmax_jl(qy, qz)));
            return fmax_jl(l, -cc);
        }
        default: break;
    }
    return 0.0;
}
template <bool RARE> BMO_D double prim_eval_f(const bmo_prim& pr, V3 q) {
    const V3 p = w2s_f(pr, q);
    if (RARE) { if (pr.type >= BMO_PRIM_CONVEX_CYL) return prim_eval_rare(pr, p); }
    const double a = pr.par[0], b = pr.par[1], c = pr.par[2], d = pr.par[3];
    switch (pr.type) {
        case BMO_PRIM_PLANO:
            return cyl_f(hyp2(p.x, p.z) - b / 2, fabs(p.y - a / 2) - a / 2);
        case BMO_PRIM_CYLINDER:
            return cyl_f(hyp2(p.x, p.z) - a, fabs(p.y) - b);
        case BMO_PRIM_SPHERE:
            return sqrt(p.x * p.x + p.y * p.y + p.z * p.z) - a;
        case BMO_PRIM_CONVEX: {
            const double q1 = hyp2(p.x, p.z), q2 = -p.y + a;
            const double h = d, R = a, hd = b / 2;
            const double s = fmax_jl((h - R) * (q1 * q1) + (hd * hd) * (h + R - 2 * q2), h * q1 - hd * q2);
            if (s < 0.0) return hyp2(q1, q2) - R;
            if (q1 < hd) return h - q2;
            return hyp2(q1 - hd, q2 - h);
        }
        case BMO_PRIM_CONCAVE: {
            const double sdf1 = cyl_f(hyp2(p.x, p.z) - b / 2, fabs(p.y + c / 2) - c / 2);
            const double ya = p.y + a;
            const double sdf2 = sqrt(p.x * p.x + ya * ya + p.z * p.z) - a;
            return fmax_jl(sdf1, -sdf2);
        }
        case BMO_PRIM_CUTSPHERE: {
            const double q1 = hyp2(p.x, p.z), q2 = p.y;
            const double h = b, R = a, w = c;
            const double s = fmax_jl((h - R) * (q1 * q1) + (w * w) * (h + R - 2 * q2), h * q1 - w * q2);
            if (s < 0.0) return hyp2(q1, q2) - R;
            if (q1 < w) return h - q2;
            return hyp2(q1 - w, q2 - h);
        }
        case BMO_PRIM_BOX: {
            const double qx = fabs(p.x) - a, qy = fabs(p.y) - b, qz = fabs(p.z) - c;
            return pnorm3(qx, qy, qz) + neg_part(fmax_jl(qx, fmax_jl(qy, qz)));
        }
        case BMO_PRIM_RING: {
            const double px = hyp2(p.x, p.z) - a;
            const double d1 = fabs(px) - b, d2 = fabs(p.y) - c;
            return pnorm2(d1, d2) + neg_part(fmax_jl(d1, d2));
        }
        case BMO_PRIM_RAPRISM: {
            const double qx = fabs(p.x) - a, qy = fabs(p.y) - b, qz = fabs(p.z) - c;
            const double box = pnorm3(qx, qy, qz) + neg_part(fmax_jl(qx, fmax_jl(qy, qz)));
            const double pln = (p.x + p.y) / 1.4142135623730951;
            return fmax_jl(box, pln);
        }
        default: break;
    }
    return 0.0;
}
// One SDF shape (a primitive or a UnionSDF) as the marching loop sees it
struct SdfShape {
    const bmo_prim* prims;
    int first, count, zr;
    double cx, cy, cz, R2;   // conservative bounding sphere (inflated by the flattener)
};
// Lower bounds of the first four member values of a union, carried along one march (result-identical
// work skipping): every member is 1-Lipschitz (exact SDFs, min / max of them), so after the march
// point has moved by s a member that was worth v is worth at least v - s.  A member whose bound
// exceeds a value already found at the new point cannot be (or tie with) the minimum, so its
// evaluation changes neither the union's value nor its arg-min.  Bounds are kept in binary32 with
// directed rounding (down for bounds, up for steps and for the running minimum) plus a 1e-7 m margin,
// orders of magnitude above the rounding of the SDF arithmetic itself.
struct MemberBounds {
    float b0, b1, b2, b3;
    int pred;   // prim index of the previous arg-min if it is a plain member (evaluated first), else -1
    BMO_D void reset() { b0 = b1 = b2 = b3 = -INFINITY; pred = -1; }
    BMO_D float get(int k) const { return k == 0 ? b0 : (k == 1 ? b1 : (k == 2 ? b2 : b3)); }
    BMO_D void set(int k, float v) { if (k == 0) b0 = v; else if (k == 1) b1 = v; else if (k == 2) b2 = v; else b3 = v; }
    BMO_D void moved(double step) {   // the march point moved by at most `step` (>= 0)
        const float s = __double2float_ru(step);
        b0 = __fsub_rd(b0, s); b1 = __fsub_rd(b1, s); b2 = __fsub_rd(b2, s); b3 = __fsub_rd(b3, s);
    }
    BMO_D void moved_f(double step, float len) {   // the same with the path length per unit of step kept in binary32 (rounded up)
        const float s = __fmul_ru(__double2float_ru(step), len);
        b0 = __fsub_rd(b0, s); b1 = __fsub_rd(b1, s); b2 = __fsub_rd(b2, s); b3 = __fsub_rd(b3, s);
    }
};
// UnionSDF.jl:53-56: minimum over the members (left fold); idx = first member attaining it, which
// is the member normal3d(::UnionSDF) dispatches to (UnionSDF.jl:86-91) -- the reference evaluates the
// members a second time for the argmin, the values are the same.  A member is a primitive or a
// meniscus frame followed by its 3 children (convex, cylinder, concave) evaluated at the frame-local
// point: max(min(convex, cylinder), -concave) (MeniscusLensSDF.jl:42-46).  Written as one flat loop
// over the prim records so that prim_eval_f is instantiated once (instruction-cache footprint).
// The previous arg-min member is evaluated first (its value bounds the minimum from above), then the
// others in order unless their lower bound rules them out; ties still go to the lowest member index.
template <bool RARE> BMO_D double shape_sdf_f(const SdfShape& sh, V3 p, unsigned& nsdf, int& idx, MemberBounds& lb) {
    const int first = sh.first, end = sh.first + sh.count;
    double m = 0.0, acc = 0.0;
    V3 q = p;
    int men = 0, start = first;
    bool have = false;
    float m_ub = INFINITY;
    const int ip = lb.pred;
    for (int j = ip >= 0 ? first - 1 : first; j < end; j++) {
        const int i = j < first ? ip : j;
        if (j >= first && i == ip) continue;                 // evaluated first
        const bmo_prim& pr = sh.prims[i];
        if (pr.type == BMO_PRIM_MENISCUS) { q = w2s_f(pr, p); men = 3; start = i; continue; }
        const int k = i - first;
        const bool tracked = men == 0 && k < 4 && !(RARE && (is_asph(pr.type) || is_acyl(pr.type)));   // aspheric pseudo-distances are not 1-Lipschitz
        if (tracked && have && lb.get(k) > m_ub) continue;   // cannot be the minimum at this point
        double v = prim_eval_f<RARE>(pr, q);
        nsdf++;
        if (men) {
            if (men == 3) acc = v;                       // convex
            else if (men == 2) acc = fmin_jl(acc, v);    // cylinder
            else { v = fmax_jl(acc, -v); q = p; }        // concave closes the member
            if (--men) continue;
        } else start = i;
        if (tracked) lb.set(k, __double2float_rd(v));
        if (!have) { m = v; idx = start; have = true; }
        else { if (v < m || (v == m && start < idx)) idx = start; m = fmin_jl(m, v); }
        m_ub = __fadd_ru(__double2float_ru(m), 1e-7f);
    }
    lb.pred = (idx - first < 4 && sh.prims[idx].type != BMO_PRIM_MENISCUS) ? idx : -1;
    return m;
}
// ---- lean union evaluation ------------------------------------------------------------------------
// The same result-identical member skipping for the common shape: a union of at most 4 plain primitives (no meniscus
// frame, no aspheric pseudo-distance) -- flagged at upload (bit 2 of the first prim record).  The bookkeeping of
// shape_sdf_f costs more instructions than the primitives it skips (ncu source page, profiles/r02e_*), so here
//   * the bounds live in shared memory ([4][128] floats per block, one column per thread) and are indexed by the member
//     number: one LDS / STS instead of select chains over four registers;
//   * a bound is stored as B_k = v_k (+) S with S = the path length marched so far, both rounded so that
//     B_k (-) S_now <= v_k - (distance moved since member k was evaluated): moving the march point is one add on S
//     instead of four subtractions;
//   * the evaluation order (previous arg-min first, then ascending) is a nibble-packed word.
// Value and arg-min (lowest index on ties) are those of the plain left fold, as in shape_sdf_f.
constexpr int LB_STRIDE = 128;    // threads per block of every kernel that marches (IBLOCK, Cfg::BLOCK)
BMO_D float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
BMO_D void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v)); }
struct LeanBounds {
    unsigned addr;   // shared-memory address of this thread's column
    float S;         // upper bound of the path length marched since reset()
    int pred;        // member number of the previous arg-min, -1: none
    BMO_D void reset() {
        S = 0.0f; pred = -1;
#pragma unroll
        for (int k = 0; k < 4; k++) sts_f32(addr + 4u * LB_STRIDE * k, -INFINITY);
    }
    BMO_D void moved_f(double step, float len) { S = __fadd_ru(S, __fmul_ru(__double2float_ru(step), len)); }
};
template <bool RARE> BMO_D double shape_eval(const SdfShape& sh, V3 p, unsigned& nsdf, int& idx_out, LeanBounds& lb) {
    const int cnt = sh.count;
    const bmo_prim* base = sh.prims + sh.first;
    // evaluation order, low nibble first: the previous arg-min, then the others ascending
    unsigned ord = lb.pred <= 0 ? 0x3210u : (lb.pred == 1 ? 0x3201u : (lb.pred == 2 ? 0x3102u : 0x2103u));
    double m = 0.0;
    int idx = 0;
    float m_ub = INFINITY;
    for (int j = 0; j < cnt; j++, ord >>= 4) {
        const int k = ord & 3;
        const unsigned a = lb.addr + 4u * LB_STRIDE * k;
        if (j > 0 && __fsub_rd(lds_f32(a), lb.S) > m_ub) continue;   // cannot be (or tie with) the minimum at this point
        const double v = prim_eval_f<RARE>(base[k], p);
        nsdf++;
        sts_f32(a, __fadd_rd(__double2float_rd(v), lb.S));
        if (j == 0) { m = v; idx = k; }
        else { if (v < m || (v == m && k < idx)) idx = k; m = fmin_jl(m, v); }
        m_ub = __fadd_ru(__double2float_ru(m), 1e-7f);
    }
    lb.pred = idx;
    idx_out = sh.first + idx;
    return m;
}
template <bool RARE> BMO_D double shape_eval(const SdfShape& sh, V3 p, unsigned& nsdf, int& idx, MemberBounds& lb) {
    return shape_sdf_f<RARE>(sh, p, nsdf, idx, lb);
}

// ---- gradients of the lens primitives with identity orientation -----------------------------------------
// prim_eval<Dual> with the seeds of an unrotated primitive (transposed_dir == I: p.x = (x; 1,0,0), p.y = (y; 0,1,0),
// p.z = (z; 0,0,1)), written out: every product of a partial with an exact 0 or 1 and every sum with an exact 0 is dropped,
// every other operation is kept in ForwardDiff's order, so the gradient has the same bits as the generic evaluation (up to the
// sign of an exact zero, which no later comparison looks at).  All operands are finite here (the caller checks), so the
// NaN-propagating branches of max / min cannot trigger.  Rule 1 for the norm of a zero vector (bmo_tables.norm_zero_rule).
// Returns false where the generic evaluation has to decide (unknown type, zero vector under the outer square root).
// tests/test_gpu_normals.py compares both evaluations bit for bit on random, on-axis, face and edge points.
BMO_D void radial_dual(double x, double z, double& r, double& gx, double& gz) {   // norm2_(p.x, p.z)
    const double sv = x * x + z * z;
    if (sv == 0.0) { r = 0.0; gx = 0.0; gz = 0.0; return; }
    r = sqrt(sv);
    const double d = 1.0 / (2 * r);
    gx = (x + x) * d; gz = (z + z) * d;
}
// cyl_(d1, d2) for d1 = (v1; gx, 0, gz), d2 = (v2; 0, sg, 0): value and gradient
BMO_D double cyl_dual(double v1, double gx, double gz, double v2, double sg, V3& g) {
    const bool pick2 = (v2 > v1) | ((int)signbit(v2) < (int)signbit(v1));   // max_w: d2 is the maximum
    const double Mv = jl_max(v1, v2);
    const bool keep = !(Mv > 0.0);                                          // min_(M, 0.0) keeps M's partials
    V3 N = mk3(0.0, 0.0, 0.0);
    if (keep) N = pick2 ? mk3(0.0, sg, 0.0) : mk3(gx, 0.0, gz);
    const double Nv = jl_min(Mv, 0.0);
    const bool z1 = signbit(v1), z2 = signbit(v2);                          // max_(d, 0.0) is the constant 0
    const double A1 = z1 ? 0.0 : v1, A2 = z2 ? 0.0 : v2;
    const double a0 = z1 ? 0.0 : gx, a2 = z1 ? 0.0 : gz, b1 = z2 ? 0.0 : sg;
    const double sv = A1 * A1 + A2 * A2;
    if (sv == 0.0) { g = N; return Nv + 0.0; }
    const double rr = sqrt(sv);
    const double dd = 1.0 / (2 * rr);
    const double s0 = a0 * A1 + a0 * A1, s1 = b1 * A2 + b1 * A2, s2 = a2 * A1 + a2 * A1;
    g = mk3(N.x + s0 * dd, N.y + s1 * dd, N.z + s2 * dd);
    return Nv + rr;
}
BMO_D bool identity_gradient(const bmo_prim& pr, V3 q, V3& g) {
    const double x = q.x - pr.pos[0], y = q.y - pr.pos[1], z = q.z - pr.pos[2];
    if (!(fabs(x) < 1e150 && fabs(y) < 1e150 && fabs(z) < 1e150)) return false;   // non-finite or absurd input: generic path
    const double a = pr.par[0], b = pr.par[1], c = pr.par[2], d = pr.par[3];
    double r, gx, gz;
    switch (pr.type) {
        case BMO_PRIM_PLANO: {
            radial_dual(x, z, r, gx, gz);
            const double t = y - a / 2;
            cyl_dual(r - b / 2, gx, gz, fabs(t) - a / 2, signbit(t) ? -1.0 : 1.0, g);
            return true;
        }
        case BMO_PRIM_CYLINDER: {
            radial_dual(x, z, r, gx, gz);
            cyl_dual(r - a, gx, gz, fabs(y) - b, signbit(y) ? -1.0 : 1.0, g);
            return true;
        }
        case BMO_PRIM_SPHERE: {
            const double sv = x * x + y * y + z * z;
            if (sv == 0.0) return false;
            const double dd = 1.0 / (2 * sqrt(sv));
            g = mk3((x + x) * dd, (y + y) * dd, (z + z) * dd);
            return true;
        }
        case BMO_PRIM_CONVEX: {
            radial_dual(x, z, r, gx, gz);
            const double q2 = -y + a;
            const double h = d, R = a, hd = b / 2;
            const double s = jl_max((h - R) * (r * r) + (hd * hd) * (h + R - 2 * q2), h * r - hd * q2);
            double A, B;
            if (s < 0.0) { A = r; B = q2; }
            else if (r < hd) { g = mk3(0.0, 1.0, 0.0); return true; }
            else { A = r - hd; B = q2 - h; }
            const double sv = A * A + B * B;
            if (sv == 0.0) return false;
            const double dd = 1.0 / (2 * sqrt(sv));
            const double s0 = gx * A + gx * A, s1 = -B - B, s2 = gz * A + gz * A;
            g = mk3(s0 * dd, s1 * dd, s2 * dd);
            return true;
        }
        case BMO_PRIM_CONCAVE: {
            radial_dual(x, z, r, gx, gz);
            const double t = y + c / 2;
            V3 g1;
            const double v1 = cyl_dual(r - b / 2, gx, gz, fabs(t) - c / 2, signbit(t) ? -1.0 : 1.0, g1);
            const double Y = y + a;
            const double sv = x * x + Y * Y + z * z;
            if (sv == 0.0) return false;
            const double r3 = sqrt(sv);
            const double d3 = 1.0 / (2 * r3);
            const double nv = -(r3 - a);
            const bool pick2 = (nv > v1) | ((int)signbit(nv) < (int)signbit(v1));   // max_(sdf1, -sdf2)
            g = pick2 ? mk3(-((x + x) * d3), -((Y + Y) * d3), -((z + z) * d3)) : g1;
            return true;
        }
        default: break;
    }
    return false;
}

// AbstractSDF.jl:79-95: ForwardDiff gradient of member idx; central differences (eps = 1e-8) if any
// component of the normalised gradient is NaN.
// The generic evaluation, out of line: ForwardDiff's dual numbers, then central differences if a component is NaN.
// skip_dual: the written-out gradient (same bits as the dual evaluation) has already produced a NaN.
template <bool RK> BMO_NI V3 member_normal_generic(const bmo_prim* prims, int idx, V3 p, int zr, Stats& st, bool skip_dual) {
    if (!skip_dual && !(RK && is_asph(prims[idx].type))) {   // aspheric surfaces: numeric_gradient only (AsphericalLensSDF.jl:3-5)
        P3<Dual> qd;
        qd.x = mkd(p.x, 1, 0, 0); qd.y = mkd(p.y, 0, 1, 0); qd.z = mkd(p.z, 0, 0, 1);
        Dual g = member_eval<Dual, RK>(prims, idx, qd, zr, st, true);
        V3 n = normalize(mk3(g.p0, g.p1, g.p2));
        if (!isnan(n.x) && !isnan(n.y) && !isnan(n.z)) return n;
    }
    const double e = 1e-8;
    P3<double> q; q.x = p.x; q.y = p.y; q.z = p.z;
    P3<double> a, b;
    V3 gr;
    a = q; b = q; a.x = p.x + e; b.x = p.x - e;
    gr.x = member_eval<double, RK>(prims, idx, a, zr, st) - member_eval<double, RK>(prims, idx, b, zr, st);
    a = q; b = q; a.y = p.y + e; b.y = p.y - e;
    gr.y = member_eval<double, RK>(prims, idx, a, zr, st) - member_eval<double, RK>(prims, idx, b, zr, st);
    a = q; b = q; a.z = p.z + e; b.z = p.z - e;
    gr.z = member_eval<double, RK>(prims, idx, a, zr, st) - member_eval<double, RK>(prims, idx, b, zr, st);
    return normalize(gr);
}
// normal3d of one union member.  FAST: unrotated lens primitives take identity_gradient (inline: no call, no callee-saved
// registers to spill) instead of the generic dual evaluation (same bits); everything else goes out of line.
template <bool RK, bool FAST = true> BMO_D V3 member_normal(const bmo_prim* prims, int idx, V3 p, int zr, Stats& st) {
    bool skip_dual = false;
    if (FAST && zr == 1 && (prims[idx].reserved & 1) && !(RK && is_asph(prims[idx].type))) {
        V3 gf;
        if (identity_gradient(prims[idx], p, gf)) {
            st.sdf++;
            const V3 n = normalize(gf);
            if (!isnan(n.x) && !isnan(n.y) && !isnan(n.z)) return n;
            skip_dual = true;
        }
    }
    return member_normal_generic<RK>(prims, idx, p, zr, st, skip_dual);
}

// intersect3d(::AbstractSDF, ray) (AbstractSDF.jl:166-181) with _raymarch_outside (:102-125) and
// _raymarch_inside (:132-159) folded into one loop, so that the SDF and the normal are each evaluated
// from a single call site (the compiler then keeps the whole march in registers):
//   INIT  d0 = sdf(pos): outside (> eps_srf) -> OUT; else normal(pos), leaving (dir.n > 0) -> miss, else IN
//   IN    fixed 1 m steps until sdf > 0, then OUT along -dir; t = t_in - t_out
//   OUT   p += dist*d; dist = sdf(p); t0 += dist; dist < eps_ray -> hit with normal(p)
// The reference re-evaluates sdf(p) when it enters _raymarch_outside; the value is the same.
// The bounding-sphere tests are result-identical: once the march point is outside the (inflated)
// sphere and moving away, every later point p + s*d (s >= 0) stays outside it, so sdf >= margin >
// eps_ray for the rest of the reference's 1000 iterations => miss.
// `nsdf` counts primitive evaluations in a register; the out-of-line normal evaluation reports its
// own count through a stack temporary that only lives around the call.
// RARE = true: the shape may contain cylindrical / aspheric members (flagged at upload); such shapes take an
// out-of-line copy of this routine (sdf_intersect_rare) so that the common path keeps its register budget.
template <bool RARE, class LB> BMO_D bool sdf_intersect_t(const SdfShape& sh, V3 pos, V3 dir, unsigned& nsdf, double& t, V3& n, LB lb) {
    {   // guaranteed miss: origin outside the bounding sphere and the line never enters it
        const V3 v = mk3(pos.x - sh.cx, pos.y - sh.cy, pos.z - sh.cz);
        const double cc = dot(v, v) - sh.R2;
        if (cc > 0.0) {
            const double b = dot(v, dir);
            if (b >= 0.0) return false;
            if (b * b - dot(dir, dir) * cc < 0.0) return false;
        }
    }
    enum { INIT = 0, IN = 1, OUT = 2 };
    // Register diet (the loop state is what the 80-register budget has to hold): the backward march direction is
    // -dir, so p + dist * (-dir) is formed as p + (-dist) * dir (negation is exact, the bits are the same) and the
    // escape test's dot(v, -dir) > 0 as dot(v, dir) < 0; the inside steps are counted (tin = n_in * 1.0 exactly).
    int mode = INIT, it = 0, n_in = 0;
    bool back = false;
    V3 p = pos;
    double t0 = 0.0;
    lb.reset();
    // |d| <= max(1, |d|^2): how far the march point moves per unit of step (directions are unit up to rounding)
    const float dlen = __double2float_ru(fmax(1.0, dot(dir, dir)) * (1.0 + 1e-9));
    for (;;) {
        int idx;
        const double dist = shape_eval<RARE>(sh, p, nsdf, idx, lb);
        if (mode == OUT) {
            t0 += dist;
            if (!(dist < eps_ray)) {
                const V3 v = mk3(p.x - sh.cx, p.y - sh.cy, p.z - sh.cz);
                const double vd = dot(v, dir);
                if (dot(v, v) > sh.R2 && (back ? vd < 0.0 : vd > 0.0)) return false;
                if (++it >= kMarchIter || !(dist == dist)) return false;   // NaN can never satisfy dist < eps_ray again
                p = p + (back ? -dist : dist) * dir; lb.moved_f(dist, dlen);
                continue;
            }
        } else if (mode == INIT) {
            if (dist > eps_srf) { mode = OUT; t0 = dist; p = p + dist * dir; lb.moved_f(dist, dlen); continue; }
        } else {  // IN
            if (dist > 0) { mode = OUT; back = true; t0 = dist; it = 0; p = p + (-dist) * dir; lb.moved_f(dist, dlen); continue; }
            if (++it >= kMarchIter) return false;
            p = p + eps_ins * dir; n_in++; lb.moved_f(eps_ins, dlen);
            continue;
        }
        {
            Stats tmp; tmp.sdf = 0; tmp.tri = 0;
            n = member_normal<RARE>(sh.prims, idx, p, sh.zr, tmp);
            nsdf += tmp.sdf;
        }
        if (mode == OUT) { t = back ? (double)n_in - t0 : t0; return true; }
        if (!(dot(dir, n) <= 0)) return false;   // on the surface, heading out
        mode = IN; it = 0;
        p = p + eps_ins * dir; n_in = 1; lb.moved_f(eps_ins, dlen);
    }
}

BMO_NI bool sdf_intersect_rare(const SdfShape& sh, V3 pos, V3 dir, unsigned& nsdf, double& t, V3& n) {
    return sdf_intersect_t<true, MemberBounds>(sh, pos, dir, nsdf, t, n, MemberBounds());
}

// ---- meshes -----------------------------------------------------------------------------------
BMO_D V3 load_vertex(const double* verts, int64_t i) { return mk3(__ldg(verts + 3 * i), __ldg(verts + 3 * i + 1), __ldg(verts + 3 * i + 2)); }
BMO_D double r32(double x) { return (double)(float)x; }
// Mesh.jl:203-237  (k_eps = l_eps = 1e-9); returns +Inf on a miss.  For Float32 meshes (STL,
// Mesh.jl:48-70) the edge vectors are formed in Float32 like the reference's Point3{Float32} math.
BMO_D double moeller_trumbore(V3 V1, V3 V2, V3 V3_, V3 rpos, V3 rdir, int f32) {
    const double ke = 1e-9, le = 1e-9;
    V3 E1 = V2 - V1, E2 = V3_ - V1;
    if (f32) { E1 = mk3(r32(E1.x), r32(E1.y), r32(E1.z)); E2 = mk3(r32(E2.x), r32(E2.y), r32(E2.z)); }
    V3 Pv = cross(rdir, E2);
    double Det = dot(E1, Pv);
    if (fabs(Det) < ke) return INFINITY;
    V3 Tv = rpos - V1;
    double invDet = 1 / Det;
    double u = dot(Tv, Pv) * invDet;
    if ((u < 0 - ke) || (u > 1 + ke)) return INFINITY;
    V3 Qv = cross(Tv, E1);
    double v = dot(rdir, Qv) * invDet;
    if ((v < 0 - ke) || (u + v > 1 + ke)) return INFINITY;
    double t = dot(E2, Qv) * invDet;
    if (t < le) return INFINITY;
    return t;
}
// Mesh.jl:183-192 then normalize(T.(normal)) of Mesh.jl:265
BMO_D V3 face_normal(V3 a, V3 b, V3 c, int f32) {
    if (f32) {
        float ax = (float)a.x, ay = (float)a.y, az = (float)a.z;
        float ux = __fsub_rn((float)b.x, ax), uy = __fsub_rn((float)b.y, ay), uz = __fsub_rn((float)b.z, az);
        float vx = __fsub_rn((float)c.x, ax), vy = __fsub_rn((float)c.y, ay), vz = __fsub_rn((float)c.z, az);
        float nx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
        float ny = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
        float nz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
        float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
        float il = __fdiv_rn(1.0f, l);
        return normalize(mk3((double)__fmul_rn(il, nx), (double)__fmul_rn(il, ny), (double)__fmul_rn(il, nz)));
    }
    return normalize(normalize(cross(b - a, c - a)));
}
BMO_D bool box_hit(const BvhNode& nd, V3 o, V3 d, double tbest) {
    double tmin = 0.0, tmax = tbest;
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (dd[k] == 0.0) {
            if (oo[k] < nd.lo[k] || oo[k] > nd.hi[k]) return false;
        } else {
            double inv = 1.0 / dd[k];
            double t1 = (nd.lo[k] - oo[k]) * inv, t2 = (nd.hi[k] - oo[k]) * inv;
            if (t1 > t2) { double tt = t1; t1 = t2; t2 = tt; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
        }
    }
    return tmin <= tmax;
}
// Mesh.jl:244-267: closest triangle, strict-min in face order => lowest face index wins ties.
// The mesh tables are passed by value: taking the address of the kernel parameter block would make
// the compiler copy it to local memory.
struct MeshTabs {
    const MeshView* meshes; const double* vertices; const int32_t* faces; const BvhNode* nodes; const int32_t* bvh_faces;
    int64_t n_vertices; int32_t n_poses, bvh_ok;
};
BMO_NI bool mesh_intersect(const MeshTabs S, int mesh_id, int pose, V3 pos, V3 dir, Stats& st, double& t, V3& n) {
    const MeshView mv = S.meshes[mesh_id];
    const double* verts = S.vertices + 3 * ((int64_t)pose * S.n_vertices + mv.first_vertex);
    const int32_t* faces = S.faces + 3 * mv.first_face;
    double t0 = INFINITY;
    int64_t fid = -1;
    if (mv.n_nodes == 0 || !S.bvh_ok) {  // small meshes, and vertex tables that are not the ones the BVH was built from (pose updates): reference order
        for (int64_t i = 0; i < mv.n_faces; i++) {
            st.tri++;
            double tt = moeller_trumbore(load_vertex(verts, __ldg(faces + 3 * i)), load_vertex(verts, __ldg(faces + 3 * i + 1)),
                                         load_vertex(verts, __ldg(faces + 3 * i + 2)), pos, dir, mv.f32);
            if (tt < t0) { t0 = tt; fid = i; }
        }
    } else {
        const BvhNode* nodes = S.nodes + mv.first_node;
        const int32_t* order = S.bvh_faces + mv.first_face;
        int32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const BvhNode nd = nodes[stack[--sp]];
            if (!box_hit(nd, pos, dir, t0)) continue;
            if (nd.count > 0) {
                for (int k = 0; k < nd.count; k++) {
                    int64_t f = order[nd.first + k];
                    st.tri++;
                    double tt = moeller_trumbore(load_vertex(verts, __ldg(faces + 3 * f)), load_vertex(verts, __ldg(faces + 3 * f + 1)),
                                                 load_vertex(verts, __ldg(faces + 3 * f + 2)), pos, dir, mv.f32);
                    if (tt < t0 || (tt == t0 && tt < INFINITY && f < fid)) { t0 = tt; fid = f; }
                }
            } else if (sp < 62) {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
    }
    if (fid < 0) return false;
    t = t0;
    n = face_normal(load_vertex(verts, __ldg(faces + 3 * fid)), load_vertex(verts, __ldg(faces + 3 * fid + 1)), load_vertex(verts, __ldg(faces + 3 * fid + 2)), mv.f32);
    return true;
}

// ---- shapes, objects, system --------------------------------------------------------------------
struct TraceCtx {
    MeshTabs M;
    const bmo_object* objects;
    int n_parts, zr;
    const bmo_prim* prims;   // prim table of this ray's pose (shared-memory copy when staged)
    const bmo_part* parts;   // part table (shared-memory copy when staged)
    const double* bounds;    // [n_parts][NBOUND] bounds of this ray's pose
    int pose;
    unsigned lb_addr;        // LEAN kernels: shared-memory address of this thread's column of member bounds (LeanBounds)
};
// intersect3d(shape, ray)
template <bool RK> BMO_D bool part_intersect(const TraceCtx& C, int part, V3 pos, V3 dir, Stats& st, double& t, V3& n) {
    const bmo_part& pt = C.parts[part];
    if (pt.shape_kind == BMO_SHAPE_SDF) {
        const double* bnd = C.bounds + NBOUND * part;
        SdfShape sh;
        sh.prims = C.prims; sh.first = pt.first; sh.count = pt.count; sh.zr = C.zr;
        sh.cx = bnd[0]; sh.cy = bnd[1]; sh.cz = bnd[2]; sh.R2 = bnd[3] * bnd[3];
        if (RK && (C.prims[pt.first].reserved & 2)) {       // bit 1: the union has cylindrical / aspheric members
            unsigned ns = 0;
            double tm = 0.0; V3 nm = mk3(0, 0, 0);
            const bool hit = sdf_intersect_rare(sh, pos, dir, ns, tm, nm);
            st.sdf += ns;
            t = tm; n = nm;
            return hit;
        }
        return sdf_intersect_t<false, MemberBounds>(sh, pos, dir, st.sdf, t, n, MemberBounds());
    }
    // results through temporaries: the out-of-line callee takes references, and t / n of the caller (shared with the
    // SDF path) must not have their address taken or they live in local memory for every part
    Stats tmp; tmp.sdf = 0; tmp.tri = 0;
    double tm = 0.0; V3 nm = mk3(0, 0, 0);
    const bool hit = mesh_intersect(C.M, pt.first, C.pose, pos, dir, tmp, tm, nm);
    st.tri += tmp.tri;
    t = tm; n = nm;
    return hit;
}
// slab test of the ray (t >= 0) against an axis-aligned box, entry distance compared with t_best
// inv = 1 / d per component, formed once per tracing_step (unused where d is 0)
BMO_D bool box_may_hit(const double* bx, V3 o, V3 d, V3 iv, double t_best) {
    double tmin = 0.0, tmax = INFINITY;
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, ii[3] = {iv.x, iv.y, iv.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (dd[k] == 0.0) {
            if (oo[k] < bx[k] || oo[k] > bx[3 + k]) return false;
        } else {
            const double inv = ii[k];
            double t1 = (bx[k] - oo[k]) * inv, t2 = (bx[3 + k] - oo[k]) * inv;
            if (t1 > t2) { const double tt = t1; t1 = t2; t2 = tt; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
        }
    }
    // 1e-9 relative slack for the rounding of the slab arithmetic (the box itself is inflated by 1e-6)
    return tmin <= tmax * (1 + 1e-9) + 1e-12 && tmin * (1 - 1e-9) - 1e-9 <= t_best;
}

// tracing_step! (System.jl:57-110).  trace_one: the hinted shape is accepted without comparing
// against other objects.  On a miss, trace_all: every object in Leaves order, strict-min t; inside an
// object the parts in shape(object) order, strict-min t (AbstractRay.jl:118-155), except that a plate
// beamsplitter prefers its coating whenever t_coating ~ t_substrate (PlateBeamsplitter.jl:160-187).
// The hinted part is not intersected again by trace_all (same ray, same shape => the same miss).
// One loop over [hint, part lo, part lo+1, ...] so that part_intersect has a single call site.
// [lo, hi) is the range of parts that trace_all looks at: the whole system for tracing_step!, the parts
// of one object (or one hinted shape) for retrace_system!'s `intersect3d(object(_intersection), ray)` /
// `intersect3d(shape(_hint), ray)` (System.jl:209-218); hi < 0 means C.n_parts.
template <bool RK> BMO_D Hit tracing_step(const TraceCtx& C, V3 pos, V3 dir, int hint_part, Stats& st, const int lo = 0, const int hi = -1) {
    const V3 inv_dir = mk3(1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z);
    Hit res; res.part = -1; res.t = INFINITY; res.n = mk3(0, 0, 0);
    Hit ob; ob.part = -1; ob.t = INFINITY; ob.n = mk3(0, 0, 0);   // best of the current object
    int cur_obj = -1;
    const int n_parts = hi < 0 ? C.n_parts : hi;
    for (int it = hint_part >= 0 ? lo - 1 : lo; it <= n_parts; it++) {
        const bool all = it >= lo;                    // false: the trace_one iteration on the hinted part
        const int part = all ? it : hint_part;
        const int obj = (all && it < n_parts) ? C.parts[part].object : -1;
        if (all && obj != cur_obj) {       // object boundary: trace_all's comparison (System.jl:62-67)
            if (ob.part >= 0 && (res.part < 0 || ob.t < res.t)) res = ob;
            ob.part = -1; ob.t = INFINITY;
            cur_obj = obj;
            if (it == n_parts) break;
        }
        double t; V3 n;
        bool hit = false;
        if (!(all && part == hint_part)) {
            // Result-identical cull: a hit point lies within 1e-10 of the part's surface, hence inside its
            // (1e-6-inflated) box.  If the ray misses the box the part cannot be hit; if it enters the box
            // only beyond the closest hit found so far, the part cannot win trace_all's strict `<`.
            // (A plate beamsplitter's substrate / coating are never culled against the best hit: its
            // coating is preferred on approximate equality, PlateBeamsplitter.jl:176-179.)
            const double* bx = C.bounds + NBOUND * part + 4;
            const bool plate = all && C.objects[obj].kind == BMO_OBJ_PLATE_BS;
            double t_best = INFINITY;          // closest hit so far: earlier objects (res) and earlier parts of this object (ob)
            if (all && !plate) {
                if (res.part >= 0) t_best = res.t;
                if (ob.part >= 0 && ob.t < t_best) t_best = ob.t;
            }
            if (box_may_hit(bx, pos, dir, inv_dir, t_best)) hit = part_intersect<RK>(C, part, pos, dir, st, t, n);
        }
        if (!all) {
            if (hit) { res.t = t; res.n = n; res.part = part; return res; }
            continue;
        }
        if (!hit) continue;
        bool take = ob.part < 0 || t < ob.t;
        if (C.parts[part].role == BMO_ROLE_COATING && C.objects[obj].kind == BMO_OBJ_PLATE_BS && ob.part >= 0)
            take = jl_isapprox(t, ob.t) ? true : (t < ob.t);   // parts = (substrate, coating)
        if (take) { ob.t = t; ob.n = n; ob.part = part; }
    }
    return res;
}

}  // namespace bmo
