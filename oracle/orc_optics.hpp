// TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// orc_optics.hpp: rays, beams, Gaussian beamlets, optical components (interact3d), detectors and
// the solver (trace_all / trace_one / trace_system! / solve_system!).  Citations are relative to
// /root/reference/src.
#pragma once
#include <deque>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include "orc_shapes.hpp"

namespace orc {

// Utils/RefractiveIndexUtils.jl: DiscreteRefractiveIndex (:8-31), SellmeierEquation (:82-98), lambda->n
struct RefIndex {
    int kind = 0;  // 0 const, 1 discrete, 2 sellmeier
    double c = 1.0;
    std::vector<double> lam, n;
    double B[3] = {0, 0, 0}, C[3] = {0, 0, 0};
    double operator()(double l) const {
        if (kind == 0) return c;
        if (kind == 1) {
            for (size_t i = 0; i < lam.size(); i++) if (lam[i] == l) return n[i];
            throw std::runtime_error("KeyError: wavelength not tabulated in DiscreteRefractiveIndex");
        }
        double x = l * 1e6;
        double n2 = 1 + (B[0] * (x * x)) / (x * x - C[0]) + (B[1] * (x * x)) / (x * x - C[1]) + (B[2] * (x * x)) / (x * x - C[2]);
        return std::sqrt(n2);
    }
};

// Rays.jl:14-42, PolarizedRays.jl:37-95
struct Ray {
    V3 pos, dir;
    Hit hit;  // intersection (valid=false <=> nothing)
    double lambda = 1000e-9, n = 1.0;
    bool polarized = false;
    Cx E0[3] = {{0, 0}, {0, 0}, {0, 0}};
    double length() const { return hit.valid ? hit.t : kInf; }          // AbstractRay.jl:173-176
    double opl() const { return hit.valid ? hit.t * n : kInf; }          // :183-186
};
inline Ray make_ray(V3 pos, V3 dir, double lambda) {  // Ray(pos, dir, lambda): dir normalised, n = 1
    Ray r; r.pos = pos; r.dir = normalize(dir); r.lambda = lambda; r.n = 1.0; return r;
}
// PolarizedRays.jl:54-56  isorthogonal3d(dir, E0; atol=1e-14): |dot(dir,E0)| <= 1e-14
inline void check_orthogonal(V3 dir, const Cx* E0) {
    Cx d{dir.x * E0[0].re + dir.y * E0[1].re + dir.z * E0[2].re, dir.x * E0[0].im + dir.y * E0[1].im + dir.z * E0[2].im};
    if (std::sqrt(abs2(d)) > 1e-14) throw std::runtime_error("Ray dir. and E0 must be orthogonal.");
}
inline Ray make_pol_ray_raw(V3 pos, V3 dir, double lambda, double n, const Cx* E0) {
    check_orthogonal(dir, E0);
    Ray r; r.pos = pos; r.dir = dir; r.lambda = lambda; r.n = n; r.polarized = true;
    for (int i = 0; i < 3; i++) r.E0[i] = E0[i];
    return r;
}
inline Ray make_pol_ray(V3 pos, V3 dir, double lambda, const Cx* E0) {  // PolarizedRays.jl:80-95
    return make_pol_ray_raw(pos, normalize(dir), lambda, 1.0, E0);
}

struct Hint {
    Object* object = nullptr;
    Shape* shape = nullptr;
    bool valid() const { return object != nullptr; }
};

// Beam.jl:13-17
struct Beam {
    std::vector<Ray> rays;
    Beam* parent = nullptr;
    std::vector<Beam*> children;
    ~Beam() { for (auto* c : children) delete c; }
    double length_rays() const {  // Beam.jl:160-169
        double l = 0;
        for (auto& r : rays) { if (!r.hit.valid) break; l += r.length(); }
        return l;
    }
    double length() const { return length_rays() + (parent ? parent->length() : 0.0); }  // :125-130 (l + l0)
    double opl() const {  // :137-149
        double l0 = parent ? parent->opl() : 0.0;
        for (auto& r : rays) { if (!r.hit.valid) break; l0 += r.opl(); }
        return l0;
    }
    // Beam.jl:177-205
    V3 point_on_beam(double t, int& index) const {
        double temp = parent ? parent->length() : 0.0;
        int numEl = (int)rays.size();
        for (int i = 0; i < numEl; i++) {
            if (i == numEl - 1) break;
            temp += rays[i].length();
            if (t < temp) {
                double b = temp - t;
                index = i;
                return rays[i].pos + (rays[i].length() - b) * rays[i].dir;
            }
        }
        const Ray& r = rays.back();
        double b = t - temp;
        index = numEl - 1;
        return r.pos + b * r.dir;
    }
    void drop_children() { for (auto* c : children) delete c; children.clear(); }
};

struct BeamInteraction {
    bool valid = false;  // false <=> nothing
    Hint hint;
    Ray ray;
};

// Gaussian.jl:33-42
struct Gauss {
    Beam chief, waist, divergence;
    double lambda, w0;
    Cx E0;
    Gauss* parent = nullptr;
    std::vector<Gauss*> children;
    ~Gauss() { for (auto* c : children) delete c; }
    double length() const { return chief.length(); }
    double opl() const { return chief.opl(); }
    void drop_children() { for (auto* c : children) delete c; children.clear(); }
};
struct GaussInteraction {
    bool valid = false;
    BeamInteraction c, w, d;
};

// Gaussian.jl:215-256 (support must be given: the reference's default is random)
inline Gauss* make_gauss(V3 position, V3 direction, double lambda, double w0, double M2, double P0, double z0, V3 support) {
    V3 dir = normalize(direction);
    V3 s1 = normalize(support);
    double tant = std::tan(M2 * lambda / (kPi * w0));  // OpticUtils.jl:63 divergence_angle
    auto* g = new Gauss();
    Ray wst = make_ray(position + s1 * w0, dir, lambda);
    V3 div_dir = normalize(dir + s1 * tant);
    double dz = -z0 * tant;
    Ray dv = make_ray(position + s1 * dz, div_dir, lambda);
    Ray chf = make_ray(position, dir, lambda);
    double I0 = 2 * P0 / (kPi * (w0 * w0));
    g->chief.rays.push_back(chf); g->waist.rays.push_back(wst); g->divergence.rays.push_back(dv);
    g->lambda = lambda; g->w0 = w0;
    g->E0 = Cx{std::sqrt(2 * I0 * kZvac), 0.0} * cis(0.0);  // OpticUtils.jl:105
    return g;
}

// Gaussian.jl:298-353
inline void gauss_parameters(const Gauss& g, V3 p0, int index, double& w, double& R, double& psi, double& w0) {
    const Ray& chief = g.chief.rays[index];
    const Ray& div = g.divergence.rays[index];
    const Ray& waist = g.waist.rays[index];
    double il = std::nan("");
    line_plane_distance3d(p0, chief.dir, div.pos, div.dir, il);
    V3 y0 = div.pos + il * div.dir - p0;
    double y_d = norm(y0);
    y0 = y0 / y_d;
    double m_d = std::tan(kHalfPi - angle3d(y0, div.dir));
    il = std::nan("");
    line_plane_distance3d(p0, chief.dir, waist.pos, waist.dir, il);
    y0 = waist.pos + il * waist.dir - p0;
    double y_w = norm(y0);
    y0 = y0 / y_w;
    double m_w = std::tan(kHalfPi - angle3d(y0, waist.dir));
    double n = chief.n;
    double H = std::fabs(n * (y_w * m_d - y_d * m_w));
    double lam = g.lambda;
    if (!(std::fabs(H - lam / kPi) <= 1e-6)) H = lam / kPi;  // isapprox(H, lam/pi, atol=1e-6); NaN -> reset
    double E_kt = y_d * m_d + y_w * m_w;
    double F_kt = std::sqrt(m_d * m_d + m_w * m_w);
    w = std::sqrt(y_d * y_d + y_w * y_w);
    R = E_kt / (w * w);
    double z = E_kt / (F_kt * F_kt);
    psi = -std::atan2(1.0, std::sqrt(1 / (R * z) - 1));  // sqrt(<0) -> NaN here (Julia would throw)
    w0 = H / (n * F_kt);
    if (std::isnan(R)) R = 0.0;
    if (std::isnan(psi)) psi = 0.0;
    if (std::isnan(w0)) w0 = w;
    if (R < 0) psi = -psi;
}
inline void gauss_parameters(const Gauss& g, double z, double& w, double& R, double& psi, double& w0) {
    int idx; V3 p0 = g.chief.point_on_beam(z, idx);
    gauss_parameters(g, p0, idx, w, R, psi, w0);
}
// Gaussian.jl:381-392 + OpticUtils.jl:87-89
inline Cx electric_field(const Gauss& g, double r, double z) {
    int index; V3 point = g.chief.point_on_beam(z, index);
    double w, R, psi, w0;
    gauss_parameters(g, point, index, w, R, psi, w0);
    double k = kTwoPi / g.lambda;
    Cx E0 = g.E0 * (g.w0 / w0);
    double dl = g.opl() - g.length();
    double ref_phi = dl / g.lambda * kTwoPi;
    Cx e = ((E0 * w0) / w) * std::exp(-(r * r) / (w * w));
    e = e * cis(k * z + psi + (k * (r * r) * R) / 2);
    return e * cis(ref_phi);
}

// OpticUtils.jl:7-9
inline V3 reflection3d(V3 dir, V3 normal) { return dir - (2 * dot(dir, normal)) * normal; }
// OpticUtils.jl:31-45
inline V3 refraction3d(V3 dir, V3 normal, double n1, double n2, bool& tir) {
    if (!jl_isapprox(norm(dir), 1.0)) throw std::invalid_argument("dir must have  unit length");
    if (!jl_isapprox(norm(normal), 1.0)) throw std::invalid_argument("norm must have  unit length");
    double n = n1 / n2;
    double cosi = -dot(normal, dir);
    double sint2 = (n * n) * (1 - cosi * cosi);
    if (sint2 > 1.0) { tir = true; return reflection3d(dir, normal); }
    tir = false;
    double cost = std::sqrt(1 - sint2);
    double f = n * cosi - cost;
    return V3{n * dir.x + f * normal.x, n * dir.y + f * normal.y, n * dir.z + f * normal.z};
}
// AbstractRay.jl:234-253
inline bool isentering(const Ray& r) { return r.hit.valid && dot(r.dir, r.hit.n) < 0; }
inline V3 refraction3d(const Ray& r, double n2, bool& tir) {
    V3 nml = r.hit.n;
    if (!isentering(r)) nml = nml * -1.0;
    return refraction3d(r.dir, nml, r.n, n2, tir);
}
// OpticUtils.jl:121-131
inline void fresnel_coefficients(double theta, double n, Cx& rs, Cx& rp, Cx& ts, Cx& tp) {
    double cost = std::cos(theta);
    double sn = std::sin(theta);
    Cx n2s2 = csqrt(Cx{n * n - sn * sn, 0.0});
    rs = (cost - n2s2) / (cost + n2s2);
    rp = ((-(n * n)) * cost + n2s2) / ((n * n) * cost + n2s2);
    ts = rs + 1.0;
    tp = Cx{2 * n * cost, 0.0} / ((n * n) * cost + n2s2);
}
inline bool is_internally_reflected(Cx rp, Cx rs) {  // :144-146
    return std::fabs(abs2(rs) - 1) <= 1e-6 && std::fabs(abs2(rp) - 1) <= 1e-6;
}

// PolarizedRays.jl:165-207.  J = diag(j11, j22, 1) in the s-p-k basis.  The random vector of the
// exactly-normal-incidence branch (LinearAlgebraUtils.jl:35-41) is pinned to Gram-Schmidt of a
// fixed seed vector (0.3, 0.5, 0.8) -- the result is independent of it whenever |j11| == |j22|.
inline V3 pinned_normal3d(V3 input) {
    V3 nw{0.3, 0.5, 0.8};
    double nn = norm(input);
    nw = nw - ((dot(nw, input) * input) / (nn * nn));
    return normalize(nw);
}
inline void calculate_global_E0(V3 in_dir, V3 out_dir, V3 normal, Cx j11, Cx j22, const Cx* Ein, Cx* Eout) {
    V3 v = !isparallel3d(in_dir, out_dir) ? out_dir : normal;
    if (isparallel3d(in_dir, normal)) v = pinned_normal3d(in_dir);
    V3 s = normalize(cross(in_dir, v));
    V3 p1 = cross(in_dir, s);
    V3 oc0 = s, oc1, oc2;
    V3 negout = -out_dir;
    bool approx_neg = norm(in_dir - negout) <= 1.4901161193847656e-8 * std::max(norm(in_dir), norm(negout));
    if (isparallel3d(in_dir, out_dir) && !approx_neg) { oc1 = p1; oc2 = in_dir; }
    else { oc1 = cross(out_dir, s); oc2 = out_dir; }
    // P = O_out * J * O_in ; O_in rows = (s, p1, in_dir); O_out cols = (oc0, oc1, oc2); E' = P * E
    // (O_out*J) first, then *(O_in), then P*E0 -- same association as Julia's left-to-right `*`.
    Cx OJ[3][3];
    V3 oc[3] = {oc0, oc1, oc2};
    Cx J[3] = {j11, j22, Cx{1, 0}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) OJ[i][j] = oc[j][i] * J[j];
    V3 rows[3] = {s, p1, in_dir};
    Cx P[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            P[i][j] = (OJ[i][0] * rows[0][j] + OJ[i][1] * rows[1][j]) + OJ[i][2] * rows[2][j];
    for (int i = 0; i < 3; i++) Eout[i] = (P[i][0] * Ein[0] + P[i][1] * Ein[1]) + P[i][2] * Ein[2];
}

// ---------------------------------------------------------------------------------------------
struct System;

enum ObjKind { O_REFRACTIVE, O_MIRROR, O_THIN_BS, O_PLATE_BS, O_CUBE_BS, O_DOUBLET, O_PD, O_SPOT, O_STOP, O_NONINT, O_GROUP, O_PSF, O_POLFILTER };

// PSFDetector.jl:1-7
struct PSFData { V3 hit, dir; double opl, proj, k; };

struct Object {
    ObjKind kind;
    Shape* shape = nullptr;          // SingleShape objects
    std::vector<Object*> parts;      // MultiShape objects: shape(object) tuple, in order
    RefIndex n;                      // refractive optics
    double reflectance = 0, transmittance = 0;  // thin BS amplitudes (sqrt(R), sqrt(1-R^2))
    // ObjectGroup state
    V3 gcenter{0, 0, 0};
    M3 gdir = M3::identity();
    // detectors
    int pd_n = 0; double pd_lo = 0, pd_hi = 0; std::vector<Cx> field;  // Photodetector (column-major [i + n*j])
    std::vector<std::array<double, 2>> spots;                          // Spotdetector
    std::vector<PSFData> psf;                                          // PSFDetector.data
    double jones[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 0}};            // PolarizationFilter.JMat (GlobalJonesBasis, real)
    double cutoff = 2.220446049250313e-16;                             // PolarizationFilter.cutoff (eps())
    double sd_hw = 0;
    explicit Object(ObjKind k) : kind(k) {}
    virtual ~Object() {}

    bool multi() const { return kind == O_PLATE_BS || kind == O_CUBE_BS || kind == O_DOUBLET || kind == O_GROUP; }
    // position/orientation of the kinematic centre (AbstractShapeTrait.jl:77-83, PlateBeamsplitter.jl:38-40,
    // ObjectGroups.jl:31-35, DoubletLenses.jl:35-36)
    V3 position() const {
        if (kind == O_GROUP) return gcenter;
        if (kind == O_PLATE_BS) return parts[1]->position();  // coating
        if (multi()) return parts[0]->position();
        return shape->pos;
    }
    M3 orientation() const {
        if (kind == O_GROUP) return gdir;
        if (multi()) return parts[0]->orientation();
        return shape->dir;
    }
    void translate(V3 off) {  // AbstractShapeTrait.jl:41,88-95 ; ObjectGroup via MultiShape
        if (!multi()) { shape->translate(off); return; }
        if (kind == O_GROUP) gcenter = gcenter + off;
        for (auto* p : parts) p->translate(off);
    }
    void translate_to(V3 target) { translate(target - position()); }
    void rotate(V3 axis, double th) {  // AbstractShapeTrait.jl:45,115-128
        if (!multi()) { shape->rotate(axis, th); return; }
        M3 R = rotate3d(axis, th);
        if (kind == O_GROUP) gdir = R * gdir;
        for (auto* p : parts) {
            p->rotate(axis, th);
            V3 v = p->position() - position();
            v = (R * v) - v;
            p->translate(v);
        }
    }
    void align(V3 target) { if (!multi()) shape->align(target); }
    void reset_translation() {
        if (!multi()) { shape->reset_translation(); return; }
        translate(-position());
        if (kind == O_GROUP) gcenter = {0, 0, 0};
    }
    void reset_rotation() {
        if (!multi()) { shape->reset_rotation(); return; }
        M3 R = orientation();
        double th = std::acos(jl_clamp((R.m[0][0] + R.m[1][1] + R.m[2][2] - 1) / 2, -1.0, 1.0));
        if (th == 0.0) return;
        double f = 1 / (2 * std::sin(th));
        V3 axis{f * (R.m[2][1] - R.m[1][2]), f * (R.m[0][2] - R.m[2][0]), f * (R.m[1][0] - R.m[0][1])};
        rotate(axis, -th);
        if (kind == O_GROUP) gdir = M3::identity();
    }
    double thickness() const {
        if (kind == O_DOUBLET) return parts[0]->shape->thickness() + parts[1]->shape->thickness();
        return shape ? shape->thickness() : 0.0;
    }

    // AbstractRay.jl:118-155 ; PlateBeamsplitter.jl:160-187 ; NonInteractable.jl:19
    Hit intersect(const Ray& ray) {
        if (kind == O_NONINT) return Hit{};
        if (kind == O_PLATE_BS) {
            Hit ic = parts[1]->intersect(ray), is = parts[0]->intersect(ray);
            if (!ic.valid && !is.valid) return Hit{};
            Hit r;
            if (!is.valid) r = ic;
            else if (!ic.valid) r = is;
            else if (jl_isapprox(ic.t, is.t)) r = ic;
            else if (ic.t < is.t) r = ic;
            else r = is;
            r.object = this;
            return r;
        }
        if (!multi()) {
            Hit h = shape->intersect(ray.pos, ray.dir);
            if (h.valid) h.object = this;
            return h;
        }
        Hit best;
        for (auto* p : parts) {
            Hit t = p->intersect(ray);
            if (!t.valid) continue;
            if (!best.valid) { best = t; continue; }
            if (t.t < best.t) best = t;
        }
        if (best.valid) best.object = this;
        return best;
    }
};

struct System {
    std::vector<Object*> objects;  // top level (may contain groups)
    std::vector<Object*> leaves;   // Leaves(objects), System.jl:21
    void flatten() {
        leaves.clear();
        std::function<void(Object*)> rec = [&](Object* o) {
            if (o->kind == O_GROUP) for (auto* c : o->parts) rec(c);
            else leaves.push_back(o);
        };
        for (auto* o : objects) rec(o);
    }
    double n(double) const { return 1.0; }  // AbstractSystem.jl:21
};

// ---------------------------------------------------------------------------------------------
// interact3d for Beam{Ray} / Beam{PolarizedRay}.  Returns an invalid interaction for `nothing`.
inline BeamInteraction interact3d(System& sys, Object* obj, Beam& beam, Ray& ray);

// Lenses.jl:46-126
inline BeamInteraction interact_refractive(System& sys, Object* optic, Beam&, Ray& ray) {
    V3 normal = ray.hit.n;
    double lambda = ray.lambda;
    double n1, n2;
    Hint hint;
    V3 npos = ray.pos + ray.length() * ray.dir;
    if (isentering(ray)) {
        n1 = ray.n; n2 = optic->n(lambda);
        hint = Hint{optic, optic->shape};
    } else {
        n1 = optic->n(lambda); n2 = sys.n(lambda);
        normal = -normal;
    }
    BeamInteraction out; out.valid = true;
    if (!ray.polarized) {
        bool tir;
        V3 ndir = refraction3d(ray.dir, normal, n1, n2, tir);
        if (tir) { hint = Hint{optic, optic->shape}; n2 = optic->n(lambda); }
        out.hint = hint;
        Ray r; r.pos = npos; r.dir = ndir; r.lambda = lambda; r.n = n2;
        out.ray = r;
        return out;
    }
    double thi = angle3d(ray.dir, -normal);
    Cx rs, rp, ts, tp;
    fresnel_coefficients(thi, n2 / n1, rs, rp, ts, tp);
    V3 ndir; Cx j11, j22;
    if (is_internally_reflected(rp, rs)) {
        hint = Hint{optic, optic->shape};
        n2 = optic->n(lambda);
        ndir = reflection3d(ray.dir, normal);
        j11 = -rs; j22 = rp;
    } else {
        bool tir;
        ndir = refraction3d(ray.dir, normal, n1, n2, tir);
        j11 = ts; j22 = tp;
    }
    Cx E0[3];
    calculate_global_E0(ray.dir, ndir, ray.hit.n, j11, j22, ray.E0, E0);
    out.hint = hint;
    out.ray = make_pol_ray_raw(npos, ndir, lambda, n2, E0);
    return out;
}
// Mirrors.jl:39-69
inline BeamInteraction interact_mirror(System&, Object*, Beam&, Ray& ray) {
    V3 normal = ray.hit.n;
    V3 npos = ray.pos + ray.length() * ray.dir;
    V3 ndir = reflection3d(ray.dir, normal);
    BeamInteraction out; out.valid = true;
    if (!ray.polarized) {
        Ray r; r.pos = npos; r.dir = ndir; r.lambda = ray.lambda; r.n = ray.n;
        out.ray = r;
        return out;
    }
    Cx E0[3];
    calculate_global_E0(ray.dir, ndir, normal, Cx{-1, 0}, Cx{1, 0}, ray.E0, E0);
    out.ray = make_pol_ray_raw(npos, ndir, ray.lambda, ray.n, E0);
    return out;
}
// ThinBeamsplitter.jl:73-106
inline Beam* bs_transmitted_beam(Object* bs, Ray& ray) {
    V3 pos = ray.pos + ray.length() * ray.dir;
    V3 dir = ray.dir;
    auto* b = new Beam();
    if (!ray.polarized) { b->rays.push_back(make_ray(pos, dir, ray.lambda)); return b; }
    Cx E0[3];
    calculate_global_E0(ray.dir, dir, ray.hit.n, Cx{bs->transmittance, 0}, Cx{bs->transmittance, 0}, ray.E0, E0);
    b->rays.push_back(make_pol_ray(pos, dir, ray.lambda, E0));
    return b;
}
inline Beam* bs_reflected_beam(Object* bs, Ray& ray) {
    V3 normal = ray.hit.n;
    V3 pos = ray.pos + ray.length() * ray.dir;
    V3 dir = reflection3d(ray.dir, normal);
    auto* b = new Beam();
    if (!ray.polarized) { b->rays.push_back(make_ray(pos, dir, ray.lambda)); return b; }
    Cx E0[3];
    calculate_global_E0(ray.dir, dir, normal, Cx{-bs->reflectance, 0}, Cx{bs->reflectance, 0}, ray.E0, E0);
    b->rays.push_back(make_pol_ray(pos, dir, ray.lambda, E0));
    return b;
}
// Beam.jl:97-112 _modify_beam_head!(old, new): the first ray of a stored child takes position, direction,
// wavelength, refractive index (and polarisation) of the freshly computed child; its stored path stays.
inline void modify_beam_head(Beam& old_, const Beam& new_) {
    Ray& o = old_.rays.front(); const Ray& n = new_.rays.front();
    o.pos = n.pos; o.dir = normalize(n.dir) /* direction! normalises, AbstractRay.jl:83-86 */; o.lambda = n.lambda; o.n = n.n;
    if (o.polarized) for (int k = 0; k < 3; k++) o.E0[k] = n.E0[k];
}
// AbstractBeam.jl:59-76 children!(beam, [t, r]): a childless beam adopts the children; a beam that
// already has as many children (retracing) only gets their heads modified; anything else is an error.
inline void set_children(Beam& beam, Beam* t, Beam* r) {
    if (beam.children.empty()) {
        t->parent = &beam; r->parent = &beam;
        beam.children = {t, r};
        return;
    }
    if (beam.children.size() == 2) {
        modify_beam_head(*beam.children[0], *t);
        modify_beam_head(*beam.children[1], *r);
        delete t; delete r;
        return;
    }
    throw std::runtime_error("Adding children to beam failed");
}
// ThinBeamsplitter.jl:108-115
inline BeamInteraction interact_thin_bs(System&, Object* bs, Beam& beam, Ray& ray) {
    set_children(beam, bs_transmitted_beam(bs, ray), bs_reflected_beam(bs, ray));
    return BeamInteraction{};
}
// direction!(ray, dir) normalises (AbstractRay.jl:83-86)
inline void set_direction(Ray& r, V3 d) { r.dir = normalize(d); }

inline BeamInteraction interact3d(System& sys, Object* obj, Beam& beam, Ray& ray) {
    switch (obj->kind) {
        case O_REFRACTIVE: return interact_refractive(sys, obj, beam, ray);
        case O_MIRROR: return interact_mirror(sys, obj, beam, ray);
        case O_THIN_BS: return interact_thin_bs(sys, obj, beam, ray);
        case O_DOUBLET: {  // DoubletLenses.jl:66-76 (only defined for Ray; PolarizedRay -> warn + nothing)
            if (ray.polarized) return BeamInteraction{};
            Object *front = obj->parts[0], *back = obj->parts[1];
            BeamInteraction i;
            if (ray.hit.shape == front->shape) { i = interact_refractive(sys, front, beam, ray); i.hint = Hint{obj, back->shape}; }
            else if (ray.hit.shape == back->shape) { i = interact_refractive(sys, back, beam, ray); i.hint = Hint{obj, front->shape}; }
            return i;
        }
        case O_CUBE_BS: {  // CubeBeamsplitter.jl:63-92
            Object *front = obj->parts[0], *back = obj->parts[1], *coat = obj->parts[2];
            if (ray.hit.shape == front->shape) { auto i = interact_refractive(sys, front, beam, ray); i.hint = Hint{obj, coat->shape}; return i; }
            if (ray.hit.shape == coat->shape) {
                interact_thin_bs(sys, coat, beam, ray);
                double n_ = front->n(ray.lambda);
                beam.children[0]->rays[0].n = n_;
                beam.children[1]->rays[0].n = n_;
                return BeamInteraction{};
            }
            if (ray.hit.shape == back->shape) { auto i = interact_refractive(sys, back, beam, ray); i.hint = Hint{obj, coat->shape}; return i; }
            return BeamInteraction{};
        }
        case O_PLATE_BS: {  // PlateBeamsplitter.jl:189-228
            Object *sub = obj->parts[0], *coat = obj->parts[1];
            if (ray.hit.shape == sub->shape) { auto i = interact_refractive(sys, sub, beam, ray); i.hint = Hint{obj, coat->shape}; return i; }
            if (ray.hit.shape == coat->shape) {
                interact_thin_bs(sys, coat, beam, ray);
                double n_optics = sub->n(ray.lambda), n_system = sys.n(ray.lambda);
                double nt, nr; V3 nd; bool tir;
                if (isentering(ray)) { nt = n_optics; nr = n_system; nd = refraction3d(ray, n_optics, tir); }
                else { nt = n_system; nr = n_optics; nd = refraction3d(ray, n_system, tir); }
                beam.children[0]->rays[0].n = nt;
                beam.children[1]->rays[0].n = nr;
                set_direction(beam.children[0]->rays[0], nd);
                return BeamInteraction{};
            }
            return BeamInteraction{};
        }
        case O_SPOT: {  // Spotdetector.jl:50-61
            V3 hit_pos = ray.pos + ray.length() * ray.dir;
            V3 loc = hit_pos - obj->shape->pos;
            double x = dot(loc, obj->shape->dir.col(0));
            double z = dot(loc, obj->shape->dir.col(2));
            obj->spots.push_back({x, z});
            return BeamInteraction{};
        }
        case O_POLFILTER: {  // Polarizers/PolarizationFilter.jl:31-47 + JonesCalculus.jl:29-46 (PolarizedRay only)
            if (!ray.polarized) return BeamInteraction{};
            V3 npos = ray.pos + ray.length() * ray.dir;
            V3 d = ray.dir;
            const M3& R = obj->shape->dir;
            double RJ[3][3], P[3][3], Q[3][3], QP[3][3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)          // R * J
                RJ[i][j] = (R.m[i][0] * obj->jones[0][j] + R.m[i][1] * obj->jones[1][j]) + R.m[i][2] * obj->jones[2][j];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)          // (R * J) * transpose(R)
                P[i][j] = (RJ[i][0] * R.m[j][0] + RJ[i][1] * R.m[j][1]) + RJ[i][2] * R.m[j][2];
            const double dv[3] = {d.x, d.y, d.z};
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)          // Q = I - in_dir * transpose(in_dir)
                Q[i][j] = (i == j ? 1.0 : 0.0) - dv[i] * dv[j];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)          // Q * P
                QP[i][j] = (Q[i][0] * P[0][j] + Q[i][1] * P[1][j]) + Q[i][2] * P[2][j];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)          // (Q * P) * Q
                P[i][j] = (QP[i][0] * Q[0][j] + QP[i][1] * Q[1][j]) + QP[i][2] * Q[2][j];
            Cx E0[3];
            for (int i = 0; i < 3; i++) E0[i] = (P[i][0] * ray.E0[0] + P[i][1] * ray.E0[1]) + P[i][2] * ray.E0[2];
            double nrm = std::sqrt((abs2(E0[0]) + abs2(E0[1])) + abs2(E0[2]));
            if (jl_isapprox(nrm, obj->cutoff)) return BeamInteraction{};     // :41-44 "terminate blocked rays"
            BeamInteraction out; out.valid = true;
            out.ray = make_pol_ray_raw(npos, d, ray.lambda, ray.n, E0);
            return out;
        }
        case O_PSF: {  // PSFDetector.jl:77-89 (defined for Beam{T, Ray{T}} only; PolarizedRay -> generic fallback, AbstractSystem.jl:30-33)
            if (ray.polarized) return BeamInteraction{};
            PSFData h;
            h.hit = ray.pos + ray.length() * ray.dir;
            h.dir = ray.dir;
            h.opl = beam.opl();
            h.proj = std::fabs(dot(ray.dir, ray.hit.n));
            h.k = kTwoPi / ray.lambda;
            obj->psf.push_back(h);
            return BeamInteraction{};
        }
        default: return BeamInteraction{};  // PD (Photodetector.jl:57-60), STOP, NONINT
    }
}

// ---------------------------------------------------------------------------------------------
// System.jl:57-110
inline Hit trace_all(System& sys, const Ray& ray) {
    Hit result;
    for (auto* obj : sys.leaves) {
        Hit temp = obj->intersect(ray);
        if (!temp.valid) continue;
        if (!result.valid || temp.t < result.t) result = temp;
    }
    return result;
}
inline Hit trace_one(System& sys, const Ray& ray, const Hint& hint) {
    Hit h = hint.shape->intersect(ray.pos, ray.dir);
    if (!h.valid) return trace_all(sys, ray);
    h.object = hint.object;
    return h;
}
inline void tracing_step(System& sys, Ray& ray, const Hint& hint) {
    ray.hit = hint.valid() ? trace_one(sys, ray, hint) : trace_all(sys, ray);
}

// System.jl:130-154
inline void trace_system(System& sys, Beam& beam, int r_max) {
    BeamInteraction interaction;
    while ((int)beam.rays.size() < r_max) {
        Ray& ray = beam.rays.back();
        tracing_step(sys, ray, interaction.valid ? interaction.hint : Hint{});
        if (!ray.hit.valid) break;
        interaction = interact3d(sys, ray.hit.object, beam, ray);
        if (!interaction.valid) break;
        beam.rays.push_back(interaction.ray);
    }
}

// ---- Gaussian interactions -------------------------------------------------------------------
inline GaussInteraction interact3d(System& sys, Object* obj, Gauss& g, int id);

// Gaussian.jl:124-135
inline GaussInteraction interact_gauss_generic(System& sys, Object* obj, Gauss& g, int id) {
    GaussInteraction gi;
    gi.c = interact3d(sys, obj, g.chief, g.chief.rays[id]);
    gi.w = interact3d(sys, obj, g.waist, g.waist.rays[id]);
    gi.d = interact3d(sys, obj, g.divergence, g.divergence.rays[id]);
    gi.valid = gi.c.valid && gi.w.valid && gi.d.valid;
    return gi;
}
// ThinBeamsplitter.jl:117-168
inline Gauss* bs_child_gauss(Object* bs, Gauss& g, int id, bool reflected) {
    auto* ch = new Gauss();
    Beam *c, *w, *d;
    if (!reflected) { c = bs_transmitted_beam(bs, g.chief.rays[id]); w = bs_transmitted_beam(bs, g.waist.rays[id]); d = bs_transmitted_beam(bs, g.divergence.rays[id]); }
    else { c = bs_reflected_beam(bs, g.chief.rays[id]); w = bs_reflected_beam(bs, g.waist.rays[id]); d = bs_reflected_beam(bs, g.divergence.rays[id]); }
    ch->chief.rays = c->rays; ch->waist.rays = w->rays; ch->divergence.rays = d->rays;
    delete c; delete w; delete d;
    double ww, R, psi, w0;
    gauss_parameters(g, g.length(), ww, R, psi, w0);
    ch->lambda = g.lambda; ch->w0 = w0;
    double amp = reflected ? bs->reflectance : bs->transmittance;
    ch->E0 = (amp * g.E0) * (g.w0 / w0);
    return ch;
}
inline GaussInteraction interact_gauss_thin_bs(System&, Object* bs, Gauss& g, int id) {
    Ray& ray = g.chief.rays[id];
    double df = dot(ray.dir, ray.hit.n);
    double phi = df < 0 ? kPi : 0.0;
    Gauss* t = bs_child_gauss(bs, g, id, false);
    Gauss* r = bs_child_gauss(bs, g, id, true);
    r->E0 = r->E0 * cis(phi);
    if (g.children.empty()) {   // AbstractBeam.jl:59-76 children!
        for (Gauss* c : {t, r}) { c->parent = &g; c->chief.parent = &g.chief; }  // Gaussian.jl:107-111
        g.children = {t, r};
    } else if (g.children.size() == 2) {   // retracing: Gaussian.jl:154-162 _modify_beam_head! (w0 is NOT updated)
        Gauss* nw[2] = {t, r};
        for (int k = 0; k < 2; k++) {
            Gauss* o = g.children[k];
            modify_beam_head(o->chief, nw[k]->chief);
            modify_beam_head(o->waist, nw[k]->waist);
            modify_beam_head(o->divergence, nw[k]->divergence);
            o->lambda = nw[k]->lambda;
            o->E0 = nw[k]->E0;
            delete nw[k];
        }
    } else throw std::runtime_error("Adding children to beam failed");
    return GaussInteraction{};
}
inline void gauss_set_n(Gauss& g, int id, double n) { g.chief.rays[id].n = n; g.waist.rays[id].n = n; g.divergence.rays[id].n = n; }

// Photodetector.jl:69-107
inline void pd_accumulate(Object* pd, Gauss& g, int ray_id) {
    Ray& ray = g.chief.rays[ray_id];
    double l0 = g.length() - ray.length();
    V3 p0 = ray.pos, d0 = ray.dir;
    M3 T = transpose(pd->shape->dir);
    V3 p = pd->shape->pos;
    if (!ray.hit.valid) return;
    double proj = std::fabs(dot(d0, ray.hit.n));
    int n = pd->pd_n;
    double sq = std::sqrt(proj);
#pragma omp parallel for schedule(static)  // Threads.@threads over pixel rows, Photodetector.jl:87
    for (int j = 0; j < n; j++) {
        double tj = (n == 1) ? 0.0 : (double)j / (double)(n - 1);
        double y = (1 - tj) * pd->pd_lo + tj * pd->pd_hi;  // LinRange lerp (Base.lerpi)
        for (int i = 0; i < n; i++) {
            double ti = (n == 1) ? 0.0 : (double)i / (double)(n - 1);
            double x = (1 - ti) * pd->pd_lo + ti * pd->pd_hi;
            V3 p1{T.m[0][0] * x + T.m[0][2] * y + p.x, T.m[1][0] * x + T.m[1][2] * y + p.y, T.m[2][0] * x + T.m[2][2] * y + p.z};
            double l1 = dot(p1 - p0, d0);
            V3 p2 = p0 + l1 * d0;
            double r = norm(p1 - p2);
            double z = l0 + l1;
            Cx e = electric_field(g, r, z) * sq;
            Cx& f = pd->field[(size_t)i + (size_t)n * j];
            f = f + e;
        }
    }
}

inline GaussInteraction interact3d(System& sys, Object* obj, Gauss& g, int id) {
    switch (obj->kind) {
        case O_THIN_BS: return interact_gauss_thin_bs(sys, obj, g, id);
        case O_PD: pd_accumulate(obj, g, id); return GaussInteraction{};
        case O_CUBE_BS: {  // CubeBeamsplitter.jl:94-121
            Object *front = obj->parts[0], *back = obj->parts[1], *coat = obj->parts[2];
            Shape* sh = g.chief.rays[id].hit.shape;
            if (sh == front->shape) { auto i = interact_gauss_generic(sys, front, g, id); if (i.valid) i.c.hint = Hint{obj, coat->shape}; return i; }
            if (sh == coat->shape) {
                interact_gauss_thin_bs(sys, coat, g, id);
                double n_ = front->n(g.lambda);
                gauss_set_n(*g.children[0], 0, n_);
                gauss_set_n(*g.children[1], 0, n_);
                return GaussInteraction{};
            }
            if (sh == back->shape) { auto i = interact_gauss_generic(sys, back, g, id); if (i.valid) i.c.hint = Hint{obj, coat->shape}; return i; }
            return GaussInteraction{};
        }
        case O_PLATE_BS: {  // PlateBeamsplitter.jl:230-275
            Object *sub = obj->parts[0], *coat = obj->parts[1];
            Shape* sh = g.chief.rays[id].hit.shape;
            if (sh == sub->shape) { auto i = interact_gauss_generic(sys, sub, g, id); if (i.valid) i.c.hint = Hint{obj, coat->shape}; return i; }
            if (sh == coat->shape) {
                interact_gauss_thin_bs(sys, coat, g, id);
                double lam = g.chief.rays[id].lambda;
                double n_optics = sub->n(lam), n_system = sys.n(lam);
                double nt, nr, n2; bool tir;
                if (isentering(g.chief.rays[id])) { nt = n_optics; nr = n_system; n2 = n_optics; }
                else { nt = n_system; nr = n_optics; n2 = n_system; }
                V3 nc = refraction3d(g.chief.rays[id], n2, tir);
                V3 nw = refraction3d(g.waist.rays[id], n2, tir);
                V3 nd = refraction3d(g.divergence.rays[id], n2, tir);
                gauss_set_n(*g.children[0], 0, nt);
                gauss_set_n(*g.children[1], 0, nr);
                set_direction(g.children[0]->chief.rays[0], nc);
                set_direction(g.children[0]->waist.rays[0], nw);
                set_direction(g.children[0]->divergence.rays[0], nd);
                return GaussInteraction{};
            }
            return GaussInteraction{};
        }
        default: return interact_gauss_generic(sys, obj, g, id);
    }
}

// Gaussian.jl:171-180
inline bool beams_hit_same_shape(Gauss& g, int id) {
    Hit &c = g.chief.rays[id].hit, &w = g.waist.rays[id].hit, &d = g.divergence.rays[id].hit;
    if (!c.valid || !w.valid || !d.valid) return !c.valid && !w.valid && !d.valid;
    return c.shape == w.shape && w.shape == d.shape;
}
// System.jl:274-318
inline void trace_system(System& sys, Gauss& g, int r_max) {
    GaussInteraction interaction;
    int seg = (int)g.chief.rays.size();
    while (seg < r_max) {
        Hint hint = interaction.valid ? interaction.c.hint : Hint{};
        Ray* ray = &g.chief.rays.back();
        tracing_step(sys, *ray, hint);
        if (!ray->hit.valid) break;
        Object* object = ray->hit.object;
        ray = &g.waist.rays.back();
        tracing_step(sys, *ray, hint);
        if (!ray->hit.valid) break;
        ray = &g.divergence.rays.back();
        tracing_step(sys, *ray, hint);
        if (!ray->hit.valid) break;
        if (!beams_hit_same_shape(g, seg - 1)) {
            g.chief.rays.back().hit = Hit{};
            g.waist.rays.back().hit = Hit{};
            g.divergence.rays.back().hit = Hit{};
            break;
        }
        interaction = interact3d(sys, object, g, seg - 1);
        if (!interaction.valid) break;
        g.chief.rays.push_back(interaction.c.ray);
        g.waist.rays.push_back(interaction.w.ray);
        g.divergence.rays.push_back(interaction.d.ray);
        seg += 1;
    }
}

// System.jl:188-255 retrace_system!(system, beam): re-validate the stored path against the previously
// hit objects (or the hinted shapes) only; the tail is dropped where the path changes.
inline void retrace_system(System& sys, Beam& beam) {
    bool cleanup_children = false, cleanup_tail = false, reset_intersection = false;
    size_t cutoff = 0;   // 1-based like the reference
    Hint hint;
    const size_t n0 = beam.rays.size();
    for (size_t i = 1; i <= n0; i++) {
        Ray& ray = beam.rays[i - 1];
        if (!ray.hit.valid) { cleanup_children = cleanup_tail = reset_intersection = true; cutoff = i; break; }
        if (!hint.valid()) ray.hit = ray.hit.object->intersect(ray);
        else {
            Object* ho = hint.object;
            ray.hit = hint.shape->intersect(ray.pos, ray.dir);
            if (ray.hit.valid) ray.hit.object = ho;
        }
        if (!ray.hit.valid) { cleanup_children = cleanup_tail = reset_intersection = true; cutoff = i; break; }
        BeamInteraction in = interact3d(sys, ray.hit.object, beam, beam.rays[i - 1]);
        hint = in.valid ? in.hint : Hint{};
        if (!in.valid) {
            if (beam.rays.size() > i) { cleanup_tail = true; cutoff = i; }
            break;
        }
        if (i < beam.rays.size()) {   // Beam.jl:82-95 replace!
            Ray& nx = beam.rays[i];
            nx.pos = in.ray.pos; nx.dir = normalize(in.ray.dir) /* direction!, AbstractRay.jl:83-86 */; nx.lambda = in.ray.lambda; nx.n = in.ray.n;
            if (nx.polarized) for (int k = 0; k < 3; k++) nx.E0[k] = in.ray.E0[k];
        } else {
            cleanup_children = true;
            beam.rays.push_back(in.ray);
            break;
        }
    }
    if (cleanup_children) beam.drop_children();
    if (cleanup_tail) beam.rays.resize(cutoff);
    if (reset_intersection) beam.rays.back().hit = Hit{};
}
// System.jl:326-428
inline void retrace_system(System& sys, Gauss& g) {
    bool cleanup_children = false, cleanup_tail = false, reset_intersection = false;
    size_t cutoff = 0;
    Hint hint;
    const size_t n_c = g.chief.rays.size();
    if (!(n_c == g.waist.rays.size() && n_c == g.divergence.rays.size())) throw std::runtime_error("Gaussian beamlet is broken");
    for (size_t i = 1; i <= n_c; i++) {
        Ray &c = g.chief.rays[i - 1], &w = g.waist.rays[i - 1], &d = g.divergence.rays[i - 1];
        if (!c.hit.valid) { cleanup_children = cleanup_tail = reset_intersection = true; cutoff = i; break; }
        Object* object;
        if (!hint.valid()) {
            object = c.hit.object;
            c.hit = object->intersect(c); w.hit = object->intersect(w); d.hit = object->intersect(d);
        } else {
            object = hint.object;
            c.hit = hint.shape->intersect(c.pos, c.dir); w.hit = hint.shape->intersect(w.pos, w.dir); d.hit = hint.shape->intersect(d.pos, d.dir);
        }
        if (!beams_hit_same_shape(g, (int)i - 1)) { cleanup_children = cleanup_tail = reset_intersection = true; cutoff = i; break; }
        if (!c.hit.valid) { cleanup_children = cleanup_tail = reset_intersection = true; cutoff = i; break; }
        c.hit.object = object; w.hit.object = object; d.hit.object = object;
        GaussInteraction in = interact3d(sys, object, g, (int)i - 1);
        hint = in.valid ? in.c.hint : Hint{};
        if (!in.valid) {
            if (n_c > i) { cleanup_tail = true; cutoff = i; }
            break;
        }
        if (i < n_c) {
            Beam* bs[3] = {&g.chief, &g.waist, &g.divergence};
            const BeamInteraction* is[3] = {&in.c, &in.w, &in.d};
            for (int k = 0; k < 3; k++) {
                Ray& nx = bs[k]->rays[i];
                nx.pos = is[k]->ray.pos; nx.dir = normalize(is[k]->ray.dir); nx.lambda = is[k]->ray.lambda; nx.n = is[k]->ray.n;
            }
        } else {
            cleanup_children = true;
            g.chief.rays.push_back(in.c.ray); g.waist.rays.push_back(in.w.ray); g.divergence.rays.push_back(in.d.ray);
            break;
        }
    }
    if (cleanup_children) g.drop_children();
    if (cleanup_tail) { g.chief.rays.resize(cutoff); g.waist.rays.resize(cutoff); g.divergence.rays.resize(cutoff); }
    if (reset_intersection) { g.chief.rays.back().hit = Hit{}; g.waist.rays.back().hit = Hit{}; g.divergence.rays.back().hit = Hit{}; }
}

// System.jl:444-475  BFS over the beam tree: optional retrace of every beam, then trace of the leaves.
// (Retracing a fresh 1-ray beam is a no-op: System.jl:200-206 only resets an intersection that is
// already nothing.)
inline void solve_system(System& sys, Beam& root, int r_max = 100, bool retrace = false) {
    std::deque<Beam*> q{&root};
    while (!q.empty()) {
        Beam* cur = q.front(); q.pop_front();
        if (retrace) retrace_system(sys, *cur);
        if (!cur->rays.back().hit.valid) trace_system(sys, *cur, r_max);
        for (auto* c : cur->children) q.push_back(c);
    }
}
inline void solve_system(System& sys, Gauss& root, int r_max = 100, bool retrace = false) {
    std::deque<Gauss*> q{&root};
    while (!q.empty()) {
        Gauss* cur = q.front(); q.pop_front();
        if (retrace) retrace_system(sys, *cur);
        if (!cur->chief.rays.back().hit.valid) trace_system(sys, *cur, r_max);
        for (auto* c : cur->children) q.push_back(c);
    }
}

}  // namespace orc
