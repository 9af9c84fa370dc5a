// TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// orc_shapes.hpp: shapes (SDF primitives, UnionSDF, MeniscusLensSDF, Mesh), their kinematics and
// intersect3d.  Each function cites the reference lines it restates (paths relative to
// /root/reference/src).
#pragma once
#include <array>
#include <memory>
#include <vector>
#include "orc_math.hpp"

namespace orc {

struct Shape;
struct Object;

// AbstractTypes/AbstractRay.jl:13-18  Intersection{T}
struct Hit {
    bool valid = false;
    double t = kInf;
    V3 n{0, 0, 0};
    Shape* shape = nullptr;
    Object* object = nullptr;
};

// AbstractTypes/AbstractShape.jl:41-114
struct Shape {
    V3 pos{0, 0, 0};
    M3 dir = M3::identity();
    virtual ~Shape() {}
    virtual bool is_sdf() const { return false; }
    virtual void set_dir(const M3& d) { dir = d; }                       // orientation!
    virtual void translate(V3 off) { pos = pos + off; }                   // AbstractShape.jl:56-59
    virtual void rotate(V3 axis, double th) { set_dir(rotate3d(axis, th) * dir); }  // :77-81
    virtual void align(V3 target) { set_dir(align3d(dir.col(1), target) * dir); }    // :107-111
    virtual void reset_translation() { pos = {0, 0, 0}; }                // :97-100
    virtual void reset_rotation() { set_dir(M3::identity()); }            // :114
    void translate_to(V3 target) { translate(target - pos); }             // :66-70
    virtual Hit intersect(V3 rpos, V3 rdir) = 0;
    virtual double thickness() const { return 0.0; }
    virtual bool has_thickness() const { return false; }
};

// ---------------------------------------------------------------------------------------------
// SDFs/AbstractSDF.jl
constexpr double eps_srf = 1e-9;   // :1
constexpr double eps_ray = 1e-10;  // :2
constexpr double eps_ins = 1.0;    // :3

struct SDF : Shape {
    M3 tdir = M3::identity();  // transposed_dir, AbstractSDF.jl:20-27
    bool is_sdf() const override { return true; }
    void set_dir(const M3& d) override { dir = d; tdir = transpose(d); }
    virtual double sdf(V3 p) const = 0;
    virtual Dual sdf(P3<Dual> p) const = 0;

    // AbstractSDF.jl:35-40  T * (point - pos)
    template <class T> P3<T> w2s(P3<T> q) const {
        T dx = q.x - pos.x, dy = q.y - pos.y, dz = q.z - pos.z;
        return {tdir.m[0][0] * dx + tdir.m[0][1] * dy + tdir.m[0][2] * dz,
                tdir.m[1][0] * dx + tdir.m[1][1] * dy + tdir.m[1][2] * dz,
                tdir.m[2][0] * dx + tdir.m[2][1] * dy + tdir.m[2][2] * dz};
    }
    // AbstractSDF.jl:81-88  numeric_gradient, eps = 1e-8
    V3 numeric_gradient(V3 p) const {
        const double e = 1e-8;
        V3 g{sdf(V3{p.x + e, p.y, p.z}) - sdf(V3{p.x - e, p.y, p.z}),
             sdf(V3{p.x, p.y + e, p.z}) - sdf(V3{p.x, p.y - e, p.z}),
             sdf(V3{p.x, p.y, p.z + e}) - sdf(V3{p.x, p.y, p.z - e})};
        return normalize(g);
    }
    // AbstractSDF.jl:90-95  normal_fd: ForwardDiff gradient, FD fallback if any NaN
    V3 normal_fd(V3 p) const {
        P3<Dual> q{{p.x, {1, 0, 0}}, {p.y, {0, 1, 0}}, {p.z, {0, 0, 1}}};
        Dual d = sdf(q);
        V3 n = normalize(V3{d.p[0], d.p[1], d.p[2]});
        if (!std::isnan(n.x) && !std::isnan(n.y) && !std::isnan(n.z)) return n;
        return numeric_gradient(p);
    }
    virtual V3 normal3d(V3 p) const { return normal_fd(p); }  // :79

    // AbstractSDF.jl:102-125
    Hit raymarch_outside(V3 p, V3 d) {
        double dist = sdf(p);
        double t0 = dist;
        for (int i = 1; i <= 1000; i++) {
            p = p + dist * d;
            dist = sdf(p);
            t0 += dist;
            if (dist < eps_ray) {
                Hit h; h.valid = true; h.t = t0; h.n = normal3d(p); h.shape = this;
                return h;
            }
        }
        return Hit{};
    }
    // AbstractSDF.jl:132-159
    Hit raymarch_inside(V3 p, V3 d) {
        double t0 = 0;
        for (int i = 1; i <= 1000; i++) {
            p = p + eps_ins * d;
            t0 += eps_ins;
            double dist = sdf(p);
            if (dist > 0) {
                Hit h = raymarch_outside(p, -d);
                if (!h.valid) break;
                h.t = t0 - h.t;
                return h;
            }
        }
        return Hit{};
    }
    // AbstractSDF.jl:166-181
    Hit intersect(V3 p, V3 d) override {
        double s = sdf(p);
        if (s > eps_srf) return raymarch_outside(p, d);
        V3 n = normal3d(p);
        if (dot(d, n) <= 0) return raymarch_inside(p, d);
        return Hit{};
    }
};

// cyl(d1,d2) = min(maximum(d), 0) + norm(max.(d, 0))   (SphericalLensSDF.jl:64, PrimitiveSDF.jl:75)
template <class T> inline T cyl_(T d1, T d2) {
    return min_(max_(d1, d2), 0.0) + norm2_(max_(d1, 0.0), max_(d2, 0.0));
}

enum PrimType { P_PLANO, P_CYLINDER, P_SPHERE, P_CONVEX, P_CONCAVE, P_CUTSPHERE, P_BOX, P_RING, P_RAPRISM, P_CONVEX_CYL, P_CONCAVE_CYL };

struct PrimSDF : SDF {
    PrimType type;
    double a = 0, b = 0, c = 0, d = 0;  // type-specific parameters, see ctor helpers below
    // PLANO: a=thickness b=diameter | CYLINDER: a=radius b=half height | SPHERE: a=radius
    // CONVEX: a=radius b=diameter c=sag d=height | CONCAVE: a=radius b=diameter c=sag
    // CUTSPHERE: a=radius b=height c=w | BOX/RAPRISM: a,b,c = half extents
    // RING: a=inner_radius(+hw) b=hwidth c=hthickness
    // CONVEX_CYL: a=radius b=cut height sqrt(r^2-(d/2)^2) c=w=sqrt(r^2-b^2) d=half extrusion height | CONCAVE_CYL: a=radius b=diameter c=height d=sag(|r|, d)
    double thick_cyl = 0.0, diam_cyl = 0.0;   // cylindrical surfaces: thickness(s) and the clear aperture
    explicit PrimSDF(PrimType t) : type(t) {}

    void set_dir(const M3& dd) override {
        if (type == P_SPHERE) return;  // SphericalLensSDF.jl:82-84 orientation fixed to I
        SDF::set_dir(dd);
    }
    double thickness() const override {
        switch (type) {
            case P_PLANO: return a;              // SphericalLensSDF.jl:28
            case P_SPHERE: return 2 * a;         // :78
            case P_CONVEX: return c;             // :196
            case P_CONCAVE: return 0.0;          // :140
            case P_BOX: return 2 * b;            // PrimitiveSDF.jl:39
            case P_CONVEX_CYL: return thick_cyl; // CylindricalSDF.jl:57  abs(sag(radius, diameter))
            case P_CONCAVE_CYL: return 0.0;      // :118
            default: return 0.0;
        }
    }
    bool has_thickness() const override {
        return type == P_PLANO || type == P_SPHERE || type == P_CONVEX || type == P_CONCAVE || type == P_BOX || type == P_CONVEX_CYL || type == P_CONCAVE_CYL;
    }
    double diameter() const { return type == P_SPHERE ? 2 * a : b; }

    template <class T> T eval(P3<T> q) const {
        P3<T> p = w2s(q);
        switch (type) {
            case P_PLANO: {  // SphericalLensSDF.jl:60-65
                T d1 = abs_(norm2_(p.x, p.z)) - b / 2;
                T d2 = abs_(p.y - a / 2) - a / 2;
                return cyl_(d1, d2);
            }
            case P_CYLINDER: {  // PrimitiveSDF.jl:71-76
                T d1 = abs_(norm2_(p.x, p.z)) - a;
                T d2 = abs_(p.y) - b;
                return cyl_(d1, d2);
            }
            case P_SPHERE:  // SphericalLensSDF.jl:86-89
                return norm3_(p.x, p.y, p.z) - a;
            case P_CONVEX: {  // SphericalLensSDF.jl:219-232
                T q1 = norm2_(p.x, p.z);
                T q2 = -p.y + a;
                double h = d, R = a, hd = b / 2;
                T s = max_((h - R) * (q1 * q1) + (hd * hd) * (h + R - 2 * q2), h * q1 - hd * q2);
                if (s < 0.0) return norm2_(q1, q2) - R;
                if (q1 < hd) return h - q2;
                return norm2_(q1 - hd, q2 - h);
            }
            case P_CONCAVE: {  // SphericalLensSDF.jl:159-170
                T psy = p.y + c / 2;
                T d1 = abs_(norm2_(p.x, p.z)) - b / 2;
                T d2 = abs_(psy) - c / 2;
                T sdf1 = cyl_(d1, d2);
                T sdf2 = norm3_(p.x, p.y + a, p.z) - a;
                return max_(sdf1, -sdf2);
            }
            case P_CUTSPHERE: {  // PrimitiveSDF.jl:112-124
                T q1 = norm2_(p.x, p.z);
                T q2 = p.y;
                double h = b, R = a, w = c;
                T s = max_((h - R) * (q1 * q1) + (w * w) * (h + R - 2 * q2), h * q1 - w * q2);
                if (s < 0.0) return norm2_(q1, q2) - R;
                if (q1 < w) return h - q2;
                return norm2_(q1 - w, q2 - h);
            }
            case P_BOX: {  // PrimitiveSDF.jl:41-46
                T qx = abs_(p.x) - a, qy = abs_(p.y) - b, qz = abs_(p.z) - c;
                return norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0)) + min_(max_(qx, max_(qy, qz)), 0.0);
            }
            case P_RING: {  // PrimitiveSDF.jl:157-166
                T px = norm2_(p.x, p.z) - a;
                T d1 = abs_(px) - b, d2 = abs_(p.y) - c;
                return norm2_(max_(d1, 0.0), max_(d2, 0.0)) + min_(max_(d1, d2), 0.0);
            }
            case P_CONVEX_CYL: {  // CylindricalSDF.jl:59-78: sdf_cut_disk in (y, z), extruded along x (AbstractSDF.jl:229-234)
                const double r = a, h = b, w = c, hx = d;
                T p1 = abs_(p.y), p2 = p.z;
                T s = max_((h - r) * (p1 * p1) + (w * w) * (h + r - 2 * p2), h * p1 - w * p2);
                T dd = (s < 0.0) ? norm2_(p1, p2) - r : ((p1 < w) ? h - p2 : norm2_(p1 - w, p2 - h));
                return cyl_(dd, abs_(p.x) - hx);
            }
            case P_CONCAVE_CYL: {  // CylindricalSDF.jl:120-133
                const double r = a, sg = d;
                T psy = p.y + (-r);
                T d1 = abs_(norm2_(p.z, psy)) - std::fabs(r);
                T d2 = abs_(p.x) - c / 2;
                T cc = cyl_(d1, d2);
                T ppy = p.y + (-sg / 2 * (r > 0 ? 1.0 : (r < 0 ? -1.0 : 0.0)));
                T qx = abs_(p.x) - c / 2, qy = abs_(ppy) - sg / 2, qz = abs_(p.z) - b / 2;
                T l = norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0)) + min_(max_(qx, max_(qy, qz)), 0.0);
                return max_(l, -cc);
            }
            case P_RAPRISM: {  // PrimitiveSDF.jl:204-210
                T qx = abs_(p.x) - a, qy = abs_(p.y) - b, qz = abs_(p.z) - c;
                T box = norm3_(max_(qx, 0.0), max_(qy, 0.0), max_(qz, 0.0)) + min_(max_(qx, max_(qy, qz)), 0.0);
                T pln = (p.x + p.y) / std::sqrt(2.0);
                return max_(box, pln);
            }
        }
        return T{};
    }
    double sdf(V3 p) const override { return eval(P3<double>{p.x, p.y, p.z}); }
    Dual sdf(P3<Dual> p) const override { return eval(p); }
};

// Utils/OpticUtils.jl:153  sag(r, l) = r - sqrt(r^2 - 0.25*l^2)
inline double sag(double r, double l) { return r - std::sqrt(r * r - 0.25 * (l * l)); }

inline PrimSDF* mk_plano(double thickness, double diameter) {  // SphericalLensSDF.jl:49-58
    auto* s = new PrimSDF(P_PLANO); s->a = thickness; s->b = diameter; return s;
}
inline PrimSDF* mk_cylinder(double r, double h) { auto* s = new PrimSDF(P_CYLINDER); s->a = r; s->b = h; return s; }
inline PrimSDF* mk_sphere(double r) { auto* s = new PrimSDF(P_SPHERE); s->a = r; return s; }
inline PrimSDF* mk_convex(double r, double d) {  // SphericalLensSDF.jl:203-217
    auto* s = new PrimSDF(P_CONVEX); s->a = r; s->b = d; s->c = sag(r, d); s->d = r - s->c; return s;
}
inline PrimSDF* mk_concave(double r, double d) {  // :147-157
    auto* s = new PrimSDF(P_CONCAVE); s->a = r; s->b = d; s->c = sag(r, d); return s;
}
inline PrimSDF* mk_cutsphere(double r, double h) {  // PrimitiveSDF.jl:97-110
    auto* s = new PrimSDF(P_CUTSPHERE); s->a = r; s->b = h; s->c = std::sqrt(r * r - h * h); return s;
}
inline PrimSDF* mk_box(double x, double y, double z) {  // :29-37
    auto* s = new PrimSDF(P_BOX); s->a = x / 2; s->b = y / 2; s->c = z / 2; return s;
}
inline PrimSDF* mk_ring(double inner_radius, double width, double thick) {  // :146-155
    auto* s = new PrimSDF(P_RING); s->a = inner_radius + width / 2; s->b = width / 2; s->c = thick / 2; return s;
}
// CylindricalSDF.jl:41-55: the cut cylinder is built along x, then xrotate3d!(s, pi/2) and translate3d!(s, [0, radius, 0])
inline PrimSDF* mk_convex_cyl(double r, double d, double h) {
    auto* s = new PrimSDF(P_CONVEX_CYL);
    s->a = r; s->b = std::sqrt(r * r - (d / 2) * (d / 2)); s->c = std::sqrt(r * r - s->b * s->b); s->d = h / 2;
    s->thick_cyl = std::fabs(sag(r, d)); s->diam_cyl = d;
    s->rotate(V3{1, 0, 0}, kPi / 2);
    s->translate(V3{0, r, 0});
    return s;
}
inline PrimSDF* mk_concave_cyl(double r, double d, double h) {  // :103-116
    auto* s = new PrimSDF(P_CONCAVE_CYL);
    s->a = r; s->b = d; s->c = h; s->d = sag(std::fabs(r), d); s->diam_cyl = d;
    return s;
}
inline PrimSDF* mk_raprism(double leg, double height) {  // :195-202
    auto* s = new PrimSDF(P_RAPRISM); s->a = leg / 2; s->b = leg / 2; s->c = height / 2; return s;
}

// SDFs/MeniscusLensSDF.jl:20-46
struct MeniscusSDF : SDF {
    SDF *convex, *cylinder, *concave;
    double thick;
    MeniscusSDF(SDF* cv, SDF* cy, SDF* cc, double l) : convex(cv), cylinder(cy), concave(cc), thick(l) {}
    double thickness() const override { return thick; }
    bool has_thickness() const override { return true; }
    template <class T> T eval(P3<T> q) const {
        P3<T> p = w2s(q);
        return max_(min_(convex->sdf(p), cylinder->sdf(p)), -concave->sdf(p));
    }
    double sdf(V3 p) const override {
        P3<double> q = w2s(P3<double>{p.x, p.y, p.z});
        V3 v{q.x, q.y, q.z};
        return jl_max(jl_min(convex->sdf(v), cylinder->sdf(v)), -concave->sdf(v));
    }
    Dual sdf(P3<Dual> p) const override { return eval(p); }
};

// SDFs/UnionSDF.jl
struct UnionSDF : SDF {
    std::vector<SDF*> sdfs;
    // UnionSDF.jl:33-42  thickness = sum over members that define thickness
    double thickness() const override {
        double t = 0;
        for (auto* s : sdfs) if (s->has_thickness()) t += s->thickness();
        return t;
    }
    bool has_thickness() const override { return true; }
    double sdf(V3 p) const override {  // :53-56 minimum over members (left fold)
        double m = sdfs[0]->sdf(p);
        for (size_t i = 1; i < sdfs.size(); i++) m = jl_min(m, sdfs[i]->sdf(p));
        return m;
    }
    Dual sdf(P3<Dual> p) const override {
        Dual m = sdfs[0]->sdf(p);
        for (size_t i = 1; i < sdfs.size(); i++) m = min_(m, sdfs[i]->sdf(p));
        return m;
    }
    void translate(V3 off) override {  // :63-67
        pos = pos + off;
        for (auto* s : sdfs) s->translate(off);
    }
    void rotate(V3 axis, double th) override {  // :69-82
        M3 R = rotate3d(axis, th);
        set_dir(R * dir);
        for (auto* s : sdfs) {
            s->rotate(axis, th);
            V3 v = s->pos - pos;
            v = (R * v) - v;
            s->translate(v);
        }
    }
    V3 normal3d(V3 p) const override {  // :86-91 normal of arg-min member (first minimum)
        size_t idx = 0;
        double m = sdfs[0]->sdf(p);
        for (size_t i = 1; i < sdfs.size(); i++) {
            double v = sdfs[i]->sdf(p);
            if (v < m) { m = v; idx = i; }
        }
        return sdfs[idx]->normal3d(p);
    }
};
// UnionSDF.jl:58-61  `+`: unions are flattened
inline UnionSDF* sdf_union(SDF* a, SDF* b) {
    auto* u = new UnionSDF();
    auto add = [&](SDF* s) {
        if (auto* us = dynamic_cast<UnionSDF*>(s)) for (auto* m : us->sdfs) u->sdfs.push_back(m);
        else u->sdfs.push_back(s);
    };
    add(a); add(b);
    return u;
}

// ---------------------------------------------------------------------------------------------
// Mesh.jl
struct Mesh : Shape {
    std::vector<V3> vertices;            // world coordinates (Mesh.jl:33-39)
    std::vector<std::array<int, 3>> faces;  // 0-based here
    double scale = 1.0;
    bool f32 = false;  // eltype Float32 (STL meshes, Mesh.jl:48-70): kinematics round to f32

    static double r32(double x) { return (double)(float)x; }
    // Mesh{Float32}: vertices, pos and dir are stored as Float32 after every kinematic call
    void round_vertices() {
        if (!f32) return;
        for (auto& v : vertices) v = {r32(v.x), r32(v.y), r32(v.z)};
        pos = {r32(pos.x), r32(pos.y), r32(pos.z)};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dir.m[i][j] = r32(dir.m[i][j]);
    }
    void translate(V3 off) override {  // Mesh.jl:78-82
        pos = pos + off;
        for (auto& v : vertices) v = v + off;
        round_vertices();
    }
    void apply_rotation(const M3& R) {  // (V .- pos') * R' .+ pos'   Mesh.jl:93,117
        for (auto& v : vertices) {
            V3 d = v - pos;
            if (f32) d = {r32(d.x), r32(d.y), r32(d.z)};  // Float32 .- Float32 stays Float32
            V3 r{d.x * R.m[0][0] + d.y * R.m[0][1] + d.z * R.m[0][2],
                 d.x * R.m[1][0] + d.y * R.m[1][1] + d.z * R.m[1][2],
                 d.x * R.m[2][0] + d.y * R.m[2][1] + d.z * R.m[2][2]};
            v = r + pos;
        }
        round_vertices();
    }
    void rotate(V3 axis, double th) override {  // Mesh.jl:89-96
        M3 R = rotate3d(axis, th);
        V3 p0 = pos;
        apply_rotation(R);
        pos = p0;
        dir = R * dir;
        round_vertices();
    }
    void align(V3 target) override {  // Mesh.jl:113-120 (note dir * R)
        M3 R = align3d(dir.col(1), target);
        apply_rotation(R);
        dir = dir * R;
        round_vertices();
    }
    void reset_translation() override { translate(-pos); }  // Mesh.jl:139-142
    void reset_rotation() override {                        // Mesh.jl:149-163
        const M3& R = dir;
        double th = std::acos(jl_clamp((R.m[0][0] + R.m[1][1] + R.m[2][2] - 1) / 2, -1.0, 1.0));
        if (th == 0.0) return;
        double f = 1 / (2 * std::sin(th));
        V3 axis{f * (R.m[2][1] - R.m[1][2]), f * (R.m[0][2] - R.m[2][0]), f * (R.m[1][0] - R.m[0][1])};
        rotate(axis, -th);
        dir = M3::identity();
    }
    void set_new_origin() { dir = M3::identity(); pos = {0, 0, 0}; }  // Mesh.jl:171-175

    // Mesh.jl:203-237  (k_eps = l_eps = 1e-9); returns Inf on miss
    // For Mesh{Float32} the edges are Float32 differences (Point3{Float32} arithmetic) before promotion.
    static double moeller_trumbore(V3 V1, V3 V2, V3 V3_, V3 rpos, V3 rdir, bool f32 = false) {
        const double ke = 1e-9, le = 1e-9;
        V3 E1 = V2 - V1, E2 = V3_ - V1;
        if (f32) { E1 = {r32(E1.x), r32(E1.y), r32(E1.z)}; E2 = {r32(E2.x), r32(E2.y), r32(E2.z)}; }
        V3 Pv = cross(rdir, E2);
        double Det = dot(E1, Pv);
        if (std::fabs(Det) < ke) return kInf;
        V3 Tv = rpos - V1;
        double invDet = 1 / Det;
        double u = dot(Tv, Pv) * invDet;
        if ((u < 0 - ke) || (u > 1 + ke)) return kInf;
        V3 Qv = cross(Tv, E1);
        double v = dot(rdir, Qv) * invDet;
        if ((v < 0 - ke) || (u + v > 1 + ke)) return kInf;
        double t = dot(E2, Qv) * invDet;
        if (t < le) return kInf;
        return t;
    }
    // Mesh.jl:183-192
    V3 face_normal(int f) const {
        V3 a = vertices[faces[f][0]], b = vertices[faces[f][1]], c = vertices[faces[f][2]];
        if (f32) {  // cross / normalize in Float32 (Point3{Float32})
            float ax = (float)a.x, ay = (float)a.y, az = (float)a.z;
            float ux = (float)b.x - ax, uy = (float)b.y - ay, uz = (float)b.z - az;
            float vx = (float)c.x - ax, vy = (float)c.y - ay, vz = (float)c.z - az;
            float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
            float il = 1.0f / std::sqrt(nx * nx + ny * ny + nz * nz);
            return V3{(double)(il * nx), (double)(il * ny), (double)(il * nz)};
        }
        return normalize(cross(b - a, c - a));
    }
    // Mesh.jl:244-267  brute force, strict-min => lowest face index wins ties
    Hit intersect(V3 rpos, V3 rdir) override {
        int fid = -1;
        double t0 = kInf;
        for (size_t i = 0; i < faces.size(); i++) {
            double t = moeller_trumbore(vertices[faces[i][0]], vertices[faces[i][1]], vertices[faces[i][2]], rpos, rdir, f32);
            if (t < t0) { t0 = t; fid = (int)i; }
        }
        if (std::isinf(t0)) return Hit{};
        Hit h; h.valid = true; h.t = t0; h.n = normalize(face_normal(fid)); h.shape = this;
        return h;
    }
};

inline Mesh* mk_rect_flat_mesh(double width, double height) {  // Mesh.jl:282-303
    auto* m = new Mesh();
    double x = width / 2, z = height / 2;
    m->vertices = {{x, 0, z}, {x, 0, -z}, {-x, 0, -z}, {-x, 0, z}};
    m->faces = {{0, 1, 3}, {1, 2, 3}};
    return m;
}
inline Mesh* mk_circ_flat_mesh(double radius, int n = 30) {  // Mesh.jl:322-348
    auto* m = new Mesh();
    m->vertices.push_back({0, 0, 0});
    // LinRange(0, 2pi*(n-1)/n, n)[i] = lerp: start + (i-1)/(n-1)*(stop-start)
    double stop = kTwoPi * (n - 1) / n;
    for (int i = 0; i < n; i++) {
        double tt = (n == 1) ? 0.0 : (double)i / (double)(n - 1);
        double x = (1 - tt) * 0.0 + tt * stop;  // Base.lerpi: (1-t)*a + t*b
        m->vertices.push_back({std::cos(x) * radius, 0, std::sin(x) * radius});
    }
    for (int i = 2; i <= n + 1; i++) m->faces.push_back({0, i - 1, i});  // [1, i, i+1] 1-based
    // faces[end] = 2 (column-major last element = faces[n,3])
    m->faces[n - 1][2] = 1;
    return m;
}
inline Mesh* mk_cuboid_mesh(double x, double y, double z, double theta = kHalfPi) {  // Mesh.jl:362-395
    auto* m = new Mesh();
    double dx = std::cos(theta) * y;
    m->vertices = {{0, 0, 0}, {x, 0, 0}, {x + dx, y, 0}, {0 + dx, y, 0}, {0 + dx, y, z}, {x + dx, y, z}, {x, 0, z}, {0, 0, z}};
    int f[12][3] = {{1, 3, 2}, {1, 4, 3}, {3, 4, 5}, {3, 5, 6}, {2, 3, 6}, {2, 6, 7},
                    {1, 8, 5}, {1, 5, 4}, {6, 5, 8}, {6, 8, 7}, {1, 7, 8}, {1, 2, 7}};
    for (auto& r : f) m->faces.push_back({r[0] - 1, r[1] - 1, r[2] - 1});
    return m;
}
inline Mesh* mk_retro_mesh(double scale) {  // OpticalComponents/Misc.jl:8-22
    auto* m = new Mesh();
    m->vertices = {{0 * scale, 0 * scale, 0 * scale}, {1 * scale, 0 * scale, 0 * scale}, {0 * scale, 1 * scale, 0 * scale}, {0 * scale, 0 * scale, 1 * scale}};
    m->faces = {{0, 2, 1}, {0, 3, 2}, {0, 1, 3}};
    m->scale = scale;
    return m;
}

}  // namespace orc
