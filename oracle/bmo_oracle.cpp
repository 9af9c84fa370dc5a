// TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// bmo_oracle.cpp: C API of the CPU oracle.  Objects are built from the reference's
// *constructor-level* arguments (SphericalLens(r1, r2, l, d, n), translate3d!, ...), independently
// of the product's flattener, so that parity tests exercise the flattener as well as the kernels.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).
#include <cstring>
#include <functional>
#include <omp.h>
#include "orc_optics.hpp"
#include "orc_asphere.hpp"

namespace orc { int g_norm_zero_rule = 1; }  // ZERO rule: pinned by test/runtests.jl:1309-1314 (see orc_math.hpp)
using namespace orc;

namespace {

struct Entry {
    Shape* shape = nullptr;
    Object* object = nullptr;
    System* system = nullptr;
    Beam* beam = nullptr;
    Gauss* gauss = nullptr;
    RefIndex* ref = nullptr;
};
std::vector<Entry> g_reg;
thread_local std::string g_err;

int reg_shape(Shape* s) { Entry e; e.shape = s; g_reg.push_back(e); return (int)g_reg.size() - 1; }
int reg_object(Object* o) { Entry e; e.object = o; g_reg.push_back(e); return (int)g_reg.size() - 1; }

Shape* S(int h) { if (h < 0 || h >= (int)g_reg.size() || !g_reg[h].shape) throw std::runtime_error("bad shape handle"); return g_reg[h].shape; }
SDF* SD(int h) { auto* s = dynamic_cast<SDF*>(S(h)); if (!s) throw std::runtime_error("handle is not an SDF"); return s; }
Mesh* M(int h) { auto* s = dynamic_cast<Mesh*>(S(h)); if (!s) throw std::runtime_error("handle is not a Mesh"); return s; }
Object* O(int h) { if (h < 0 || h >= (int)g_reg.size() || !g_reg[h].object) throw std::runtime_error("bad object handle"); return g_reg[h].object; }
RefIndex RI(int h) { if (h < 0 || h >= (int)g_reg.size() || !g_reg[h].ref) throw std::runtime_error("bad refindex handle"); return *g_reg[h].ref; }

Object* mk_obj(ObjKind k, Shape* s) { auto* o = new Object(k); o->shape = s; return o; }
Object* mk_refr(Shape* s, const RefIndex& n) { auto* o = mk_obj(O_REFRACTIVE, s); o->n = n; return o; }

// ---- spherical surfaces -> SDFs (SphericalLensSDF.jl:423-454) ---------------------------------
SDF* surf_forward(double r, double d) {
    if (std::isinf(r)) return nullptr;
    return r > 0 ? (SDF*)mk_convex(r, d) : (SDF*)mk_concave(std::fabs(r), d);
}
SDF* surf_backward(double r, double d) {
    if (std::isinf(r)) return nullptr;
    SDF* b = r > 0 ? (SDF*)mk_concave(r, d) : (SDF*)mk_convex(std::fabs(r), d);
    b->rotate(V3{0, 0, 1}, kPi);
    return b;
}
// edge_sag(surface, sdf): SphericalLensSDF.jl:421 (the stored sag) / AsphericalLensSDF.jl:434-443 (signed sag at the rim)
double edge_sag_of(SDF* s) {
    if (auto* a = dynamic_cast<AsphSDF*>(s)) return a->edge_sag();
    return static_cast<PrimSDF*>(s)->c;
}
// EvenAsphericalSurface -> SDF (AsphericalLensSDF.jl:452-474): no rotation for the backward orientation
struct AsphDesc { double k; std::vector<double> coeffs; };
SDF* asph_forward(const AsphDesc& a, double r, double d) {
    if (std::isinf(r)) return nullptr;
    return new AsphSDF(r > 0, a.coeffs, r, a.k, d);
}
SDF* asph_backward(const AsphDesc& a, double r, double d) {
    if (std::isinf(r)) return nullptr;
    return new AsphSDF(!(r > 0), a.coeffs, r, a.k, d);
}
double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

// SDFs/MeniscusLensSDF.jl:122-189
SDF* meniscus_lens_sdf(double r1, double d1, SDF* front, double r2, double d2, SDF* back, double ct) {
    bool left;
    if (sgn(r1) == sgn(r2) && sgn(r2) > 0) left = true;
    else if (sgn(r1) == sgn(r2) && sgn(r2) < 0) left = false;
    else throw std::invalid_argument("Invalid sign combination for r1 and r2");
    double convex_sag = edge_sag_of(front);
    double concave_sag = edge_sag_of(back);
    double cylinder_l = ct - convex_sag + concave_sag;
    if (cylinder_l <= 0) throw std::runtime_error("Lens parameters lead to zero lens edge thickness");
    SDF *f, *b;
    if (left) { f = mk_convex(r1, d1); b = mk_sphere(r2); }
    else { f = mk_sphere(std::fabs(r1)); b = mk_convex(std::fabs(r2), d2); }
    double d_mid = std::min(d1, d2);
    PrimSDF* cyl = mk_plano(cylinder_l, d_mid);
    SDF *cvx, *ccv;
    if (left) {
        cyl->translate(V3{0, f->thickness(), 0});
        b->translate(V3{0, r2 + ct, 0});
        cvx = f; ccv = b;
    } else {
        b->translate(V3{0, -std::fabs(r1), 0});
        cyl->translate(V3{0, -concave_sag, 0});
        f->rotate(V3{0, 0, 1}, kPi);
        f->translate(V3{0, cyl->thickness() - concave_sag + convex_sag, 0});
        cvx = b; ccv = f;
    }
    return new MeniscusSDF(cvx, cyl, ccv, ct);
}

// OpticalComponents/Lenses.jl:176-291  Lens(front_surface, back_surface, center_thickness, n)
SDF* lens_shape(double r1, double d1, double md1, double r2, double d2, double md2, double ct,
                const AsphDesc* a1 = nullptr, const AsphDesc* a2 = nullptr) {
    double d_mid = std::min(d1, d2), md_mid = std::max(md1, md2);
    double l0 = ct;
    SDF* front = a1 ? asph_forward(*a1, r1, d1) : surf_forward(r1, d1);
    l0 -= front ? front->thickness() : 0.0;
    SDF* back = a2 ? asph_backward(*a2, r2, d2) : surf_backward(r2, d2);
    l0 -= back ? back->thickness() : 0.0;
    if (l0 <= 0 && (a1 || a2)) throw std::invalid_argument("only spherical meniscus lenses are supported (Lenses.jl:188-190)");
    if (!front && !back) return mk_plano(ct, d_mid);  // Lenses.jl:303-311
    SDF* shape;
    if (l0 <= 0) {
        if (sgn(r1) == sgn(r2)) {
            shape = meniscus_lens_sdf(r1, d1, front, r2, d2, back, ct);
            if (md_mid > d_mid) {
                auto* men = static_cast<MeniscusSDF*>(shape);
                double th = men->cylinder->thickness();
                V3 p = men->cylinder->pos;
                SDF* ring = mk_ring(d_mid / 2, (md_mid - d_mid) / 2, th);
                ring->translate(V3{0, p.y + th / 2, 0});
                shape = sdf_union(shape, ring);
            }
        } else throw std::invalid_argument("Lens parameters lead to cylinder section length of <= 0, use ThinLens instead.");
        return shape;
    }
    SDF* mid = mk_plano(l0, d_mid);
    SDF* plano = mid;
    if (front) { mid->translate(V3{0, front->thickness(), 0}); mid = sdf_union(mid, front); }
    if (back) { back->translate(V3{0, mid->thickness() + back->thickness(), 0}); mid = sdf_union(mid, back); }
    shape = mid;
    double d_front = d1, d_back = d2, d_min = std::min(d1, d2), d_max = std::max(d1, d2);
    if (md_mid < d_min) return shape;
    if (d_front != d_back) {
        if (d_back > d_front) {
            double lt = l0;
            if (front) { double sf = edge_sag_of(front); if (sf < 0) lt += std::fabs(sf) + front->thickness(); }
            SDF* ring = mk_ring(d_front / 2, (d_back - d_front) / 2, lt);
            ring->translate(V3{0, (front ? edge_sag_of(front) : 0.0) + lt / 2, 0});
            shape = sdf_union(shape, ring);
        } else {
            double lt = l0;
            if (back) { double sb = edge_sag_of(back); if ((sb - back->thickness()) > 0) lt += std::fabs(sb) + back->thickness(); }
            SDF* ring = mk_ring(d_back / 2, (d_front - d_back) / 2, lt);
            ring->translate(V3{0, (front ? front->thickness() : 0.0) + lt / 2, 0});
            shape = sdf_union(shape, ring);
        }
    }
    if (md_mid > d_max) {
        double ot = mid->thickness();
        double oc = mid->pos.y + ot / 2;   // position(mid) of the union = 0 unless translated
        (void)plano;
        if (front) { double sf = edge_sag_of(front); ot -= sf; oc += sf / 2; }
        if (back) { double sb = edge_sag_of(back); ot += sb; oc += sb / 2; }
        SDF* ring = mk_ring(d_max / 2, (md_mid - d_max) / 2, ot);
        ring->translate(V3{0, oc, 0});
        shape = sdf_union(shape, ring);
    }
    return shape;
}
// OpticalComponents/Lenses.jl:331-379 Lens(front::AbstractCylindricalSurface, back, center_thickness, n) with
// CylindricalSDF.jl:176-205 (surface -> SDF; r = Inf stands for RectangularFlatSurface / a flat side)
// edge_sag of a cylindric surface: thickness(sdf) (CylindricalSDF.jl:173-174) / the signed aspheric sag (AcylindricalSDF.jl:170-178)
static double cyl_edge_sag(SDF* s) {
    if (auto* a = dynamic_cast<AcylSDF*>(s)) return a->edge_sag();
    return s->thickness();
}
SDF* cyl_lens_shape(double r1, double d1, double h1, double md1, double r2, double d2, double h2, double md2, double ct,
                    const AsphDesc* a1 = nullptr, const AsphDesc* a2 = nullptr) {
    const bool flat1 = std::isinf(r1), flat2 = std::isinf(r2);
    SDF* front = flat1 ? nullptr : (a1 ? (SDF*)new AcylSDF(r1 > 0, r1, d1, h1, a1->k, a1->coeffs)     // AcylindricalSDF.jl:184-191
                                       : (r1 > 0 ? (SDF*)mk_convex_cyl(r1, d1, h1) : (SDF*)mk_concave_cyl(r1, d1, h1)));
    SDF* back = flat2 ? nullptr : (a2 ? (r2 > 0 ? (SDF*)new AcylSDF(false, r2, d2, h2, a2->k, a2->coeffs)
                                                : (SDF*)new AcylSDF(true, -r2, d2, h2, a2->k, a2->coeffs))      // :192-199
                                      : (r2 > 0 ? (SDF*)mk_concave_cyl(r2, d2, h2) : (SDF*)mk_convex_cyl(-r2, d2, h2)));
    double l0 = ct;
    l0 -= front ? front->thickness() : 0.0;
    l0 -= back ? back->thickness() : 0.0;
    double d_mid, md_mid, h;
    if (flat1) throw std::runtime_error("cylindric lens: the front surface must be a CylindricalSurface");
    if (flat2) { d_mid = d1; md_mid = md1; h = h1; }
    else {
        if (h1 != h2) throw std::runtime_error("height of front and back surface have to match for cylindric lenses");
        d_mid = std::min(d1, d2); md_mid = std::max(md1, md2); h = h1;
    }
    if (l0 <= 0) throw std::runtime_error("Lens parameters lead to a box section length of <= 0");
    SDF* mid = mk_box(h, l0, d_mid);
    mid->translate(V3{0, l0 / 2, 0});
    if (front) { mid->translate(V3{0, front->thickness(), 0}); mid = sdf_union(mid, front); }
    if (back) { back->translate(V3{0, mid->thickness() + back->thickness(), 0}); mid = sdf_union(mid, back); }
    SDF* shape = mid;
    if (md_mid > d_mid) {
        double ring_thickness = mid->thickness();
        double ring_center = mid->pos.y + ring_thickness / 2;   // position(mid): of the box if there is no union, else 0
        if (front) { double sg = cyl_edge_sag(front); ring_thickness -= sg; ring_center += sg / 2; }
        if (back) { double sg = cyl_edge_sag(back); ring_thickness += sg; ring_center += sg / 2; }
        SDF* ring = mk_ring(d_mid / 2, (md_mid - d_mid) / 2, ring_thickness);
        ring->translate(V3{0, ring_center, 0});
        shape = sdf_union(shape, ring);
    }
    return shape;
}

// SphericalLensSDF.jl:245-253
SDF* thin_lens_sdf(double r1, double r2, double d) {
    SDF* front = mk_convex(r1, d);
    SDF* back = mk_convex(r2, d);
    back->translate(V3{0, front->thickness() + back->thickness(), 0});
    back->rotate(V3{0, 0, 1}, kPi);
    return sdf_union(front, back);
}
// SphericalLenses.jl:20-32
Object* spherical_lens(double r1, double r2, double l, double d, const RefIndex& n) {
    if (l == 0.0) return mk_refr(thin_lens_sdf(r1, r2, d), n);
    return mk_refr(lens_shape(r1, d, d, r2, d, d, l), n);
}
// ThinBeamsplitter.jl:43-51
Object* thin_bs(Shape* shape, double reflectance) {
    if (reflectance >= 1 || jl_isapprox(reflectance, 0.0)) throw std::runtime_error("Splitting ratio in (0, 1)!");
    auto* o = mk_obj(O_THIN_BS, shape);
    o->reflectance = std::sqrt(reflectance);
    o->transmittance = std::sqrt(1 - o->reflectance * o->reflectance);
    return o;
}
double deg2rad(double d) { return d * (kPi / 180.0); }  // Base.deg2rad: z * (pi/180)

struct ObjExport { std::vector<Object*> leaves; };

int leaf_index(System* sys, Object* o) {
    for (size_t i = 0; i < sys->leaves.size(); i++) if (sys->leaves[i] == o) return (int)i;
    return -1;
}
int part_index(Object* o, Shape* s) {
    if (!o) return -1;
    if (!o->multi()) return 0;
    for (size_t i = 0; i < o->parts.size(); i++) if (o->parts[i]->shape == s) return (int)i;
    return -1;
}

// thread-local sink so that bulk (OpenMP) traces do not race on Spotdetector::spots
}  // namespace

#define ORC_TRY try {
#define ORC_CATCH(rv) } catch (const std::exception& e) { g_err = e.what(); return rv; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_set_norm_zero_rule(int r) { g_norm_zero_rule = r; }
void orc_reset() { g_reg.clear(); }  // leaks by design (test process lifetime)

int orc_refindex(int kind, const double* a, int na) {
    ORC_TRY
    auto* r = new RefIndex();
    r->kind = kind;
    if (kind == 0) r->c = a[0];
    else if (kind == 1) { int m = na / 2; for (int i = 0; i < m; i++) { r->lam.push_back(a[i]); r->n.push_back(a[m + i]); } }
    else { for (int i = 0; i < 3; i++) { r->B[i] = a[i]; r->C[i] = a[3 + i]; } }
    Entry e; e.ref = r; g_reg.push_back(e);
    return (int)g_reg.size() - 1;
    ORC_CATCH(-1)
}

// Generic constructor.  `kind` names follow the reference's constructors.
int orc_new(const char* kind, const double* d, int nd, const int* ih, int ni) {
    ORC_TRY
    std::string k(kind);
    (void)nd;
    // ---- shapes
    if (k == "PlanoSurfaceSDF") return reg_shape(mk_plano(d[0], d[1]));
    if (k == "CylinderSDF") return reg_shape(mk_cylinder(d[0], d[1]));
    if (k == "SphereSDF") return reg_shape(mk_sphere(d[0]));
    if (k == "ConvexSphericalSurfaceSDF") return reg_shape(mk_convex(d[0], d[1]));
    if (k == "ConcaveSphericalSurfaceSDF") return reg_shape(mk_concave(d[0], d[1]));
    if (k == "CutSphereSDF") return reg_shape(mk_cutsphere(d[0], d[1]));
    if (k == "BoxSDF") return reg_shape(mk_box(d[0], d[1], d[2]));
    if (k == "RingSDF") return reg_shape(mk_ring(d[0], d[1], d[2]));
    if (k == "RightAnglePrismSDF") return reg_shape(mk_raprism(d[0], d[1]));
    if (k == "ThinLensSDF") return reg_shape(thin_lens_sdf(d[0], d[1], d[2]));
    if (k == "UnionSDF") { SDF* u = SD(ih[0]); for (int i = 1; i < ni; i++) u = sdf_union(u, SD(ih[i])); return reg_shape(u); }
    if (k == "LensSDF") return reg_shape(lens_shape(d[0], d[1], d[2], d[3], d[4], d[5], d[6]));
    if (k == "RectangularFlatMesh") return reg_shape(mk_rect_flat_mesh(d[0], d[1]));
    if (k == "CircularFlatMesh") return reg_shape(mk_circ_flat_mesh(d[0], ni > 0 ? ih[0] : 30));
    if (k == "CuboidMesh") return reg_shape(mk_cuboid_mesh(d[0], d[1], d[2], nd > 3 ? d[3] : kHalfPi));
    if (k == "RetroMesh") return reg_shape(mk_retro_mesh(d[0]));
    if (k == "Mesh") {  // ih = [nv, nf, f32, faces(0-based)...], d = vertices xyz...
        auto* m = new Mesh();
        int nv = ih[0], nf = ih[1];
        m->f32 = ih[2] != 0;
        for (int i = 0; i < nv; i++) m->vertices.push_back({d[3 * i], d[3 * i + 1], d[3 * i + 2]});
        for (int i = 0; i < nf; i++) m->faces.push_back({ih[3 + 3 * i], ih[4 + 3 * i], ih[5 + 3 * i]});
        if (m->f32) m->scale = (double)(float)1e-3;
        return reg_shape(m);
    }
    // ---- objects
    if (k == "Lens" || k == "Prism") return reg_object(mk_refr(S(ih[0]), RI(ih[1])));
    if (k == "Mirror" || k == "Retroreflector_from_mesh") return reg_object(mk_obj(O_MIRROR, S(ih[0])));
    if (k == "IntersectableObject") return reg_object(mk_obj(O_STOP, S(ih[0])));
    if (k == "NonInteractableObject") return reg_object(mk_obj(O_NONINT, S(ih[0])));
    if (k == "SphericalLens") return reg_object(spherical_lens(d[0], d[1], d[2], d[3], RI(ih[0])));
    if (k == "LensFromSurfaces") return reg_object(mk_refr(lens_shape(d[0], d[1], d[2], d[3], d[4], d[5], d[6]), RI(ih[0])));
    if (k == "AsphericLens") {
        // r1 d1 md1 k1 nc1 | r2 d2 md2 k2 nc2 | ct | coeffs1... coeffs2...   (nc < 0: SphericalSurface / CircularFlatSurface)
        int nc1 = (int)d[4], nc2 = (int)d[9];
        AsphDesc a1, a2;
        int o = 11;
        if (nc1 >= 0) { a1.k = d[3]; a1.coeffs.assign(d + o, d + o + nc1); o += nc1; }
        if (nc2 >= 0) { a2.k = d[8]; a2.coeffs.assign(d + o, d + o + nc2); o += nc2; }
        return reg_object(mk_refr(lens_shape(d[0], d[1], d[2], d[5], d[6], d[7], d[10], nc1 >= 0 ? &a1 : nullptr, nc2 >= 0 ? &a2 : nullptr), RI(ih[0])));
    }
    if (k == "AcylindricalLens") {
        // r1 d1 h1 md1 k1 nc1 | r2 d2 h2 md2 k2 nc2 | ct | coeffs1... coeffs2...   (nc < 0: CylindricalSurface; r2 = Inf: flat)
        int nc1 = (int)d[5], nc2 = (int)d[11];
        AsphDesc a1, a2;
        int o = 13;
        if (nc1 >= 0) { a1.k = d[4]; a1.coeffs.assign(d + o, d + o + nc1); o += nc1; }
        if (nc2 >= 0) { a2.k = d[10]; a2.coeffs.assign(d + o, d + o + nc2); o += nc2; }
        return reg_object(mk_refr(cyl_lens_shape(d[0], d[1], d[2], d[3], d[6], d[7], d[8], d[9], d[12], nc1 >= 0 ? &a1 : nullptr, nc2 >= 0 ? &a2 : nullptr), RI(ih[0])));
    }
    if (k == "CylindricalLens")   // r1 d1 h1 md1 r2 d2 h2 md2 ct (r2 = Inf: RectangularFlatSurface)
        return reg_object(mk_refr(cyl_lens_shape(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]), RI(ih[0])));
    if (k == "ThinLens") return reg_object(mk_refr(thin_lens_sdf(d[0], d[1], d[2]), RI(ih[0])));
    if (k == "SphericalDoubletLens") {  // DoubletLenses.jl:57-64: r1 r2 r3 l1 l2 d, n1 n2
        Object* front = spherical_lens(d[0], d[1], d[3], d[5], RI(ih[0]));
        Object* back = spherical_lens(d[1], d[2], d[4], d[5], RI(ih[1]));
        back->translate(V3{0, front->shape->thickness(), 0});
        auto* dl = new Object(O_DOUBLET);
        dl->parts = {front, back};
        return reg_object(dl);
    }
    if (k == "RoundPlanoMirror") return reg_object(mk_obj(O_MIRROR, mk_plano(d[1], d[0])));  // (diameter, thickness) Mirrors.jl:154-157
    if (k == "SquarePlanoMirror2D") return reg_object(mk_obj(O_MIRROR, mk_rect_flat_mesh(d[0], d[0])));
    if (k == "RectangularPlanoMirror") {  // Mirrors.jl:93-102 (width, height, thickness)
        Mesh* m = mk_cuboid_mesh(d[0], d[2], d[1]);
        m->translate(V3{-d[0] / 2, 0, -d[1] / 2});
        m->set_new_origin();
        return reg_object(mk_obj(O_MIRROR, m));
    }
    if (k == "ConcaveSphericalMirror") {  // Mirrors.jl:186-191 (radius, thickness, diameter)
        SDF* cyl = mk_plano(d[1], d[2]);
        SDF* cc = mk_concave(std::fabs(d[0]), d[2]);
        return reg_object(mk_obj(O_MIRROR, sdf_union(cc, cyl)));
    }
    if (k == "RightAnglePrismMirror") {  // Mirrors.jl:214-218
        SDF* s = mk_raprism(d[0], d[1]);
        s->rotate(V3{0, 0, 1}, deg2rad(45 + 180));
        return reg_object(mk_obj(O_MIRROR, s));
    }
    if (k == "Retroreflector") return reg_object(mk_obj(O_MIRROR, mk_retro_mesh(d[0])));
    if (k == "RightAnglePrism") return reg_object(mk_refr(mk_raprism(d[0], d[1]), RI(ih[0])));
    if (k == "RectangularCompensatorPlate") {  // Compensators.jl:15-24
        Mesh* m = mk_cuboid_mesh(d[0], d[2], d[1]);
        m->translate(V3{-d[0] / 2, 0, -d[1] / 2});
        m->set_new_origin();
        return reg_object(mk_refr(m, RI(ih[0])));
    }
    if (k == "ThinBeamsplitter") return reg_object(thin_bs(mk_rect_flat_mesh(d[0], d[1]), d[2]));
    if (k == "RoundThinBeamsplitter") return reg_object(thin_bs(mk_circ_flat_mesh(d[0] / 2), d[1]));
    if (k == "ThinBeamsplitterFromShape") return reg_object(thin_bs(S(ih[0]), d[0]));
    if (k == "RectangularPlateBeamsplitter") {  // PlateBeamsplitter.jl:89-104 (w, h, t, R)
        Object* sub = mk_refr(mk_box(d[0], d[2], d[1]), RI(ih[0]));
        sub->translate(V3{0, d[2] / 2, 0});
        Object* coat = thin_bs(mk_rect_flat_mesh(d[0], d[1]), d[3]);
        coat->rotate(V3{0, 0, 1}, kPi);
        auto* o = new Object(O_PLATE_BS);
        o->parts = {sub, coat};
        return reg_object(o);
    }
    if (k == "RoundPlateBeamsplitter") {  // :146-158 (diameter, thickness, R)
        Object* sub = mk_refr(mk_plano(d[1], d[0]), RI(ih[0]));
        Object* coat = thin_bs(mk_circ_flat_mesh(d[0] / 2), d[2]);
        auto* o = new Object(O_PLATE_BS);
        o->parts = {sub, coat};
        return reg_object(o);
    }
    if (k == "CubeBeamsplitter") {  // CubeBeamsplitter.jl:49-61 (leg, R)
        Object* front = mk_refr(mk_raprism(d[0], d[0]), RI(ih[0]));
        Object* back = mk_refr(mk_raprism(d[0], d[0]), RI(ih[0]));
        Object* bs = thin_bs(mk_rect_flat_mesh(std::sqrt(2.0) * d[0], d[0]), d[1]);
        back->rotate(V3{0, 0, 1}, deg2rad(180));
        bs->rotate(V3{0, 0, 1}, deg2rad(180 - 45));
        static_cast<Mesh*>(bs->shape)->set_new_origin();
        auto* o = new Object(O_CUBE_BS);
        o->parts = {front, back, bs};
        return reg_object(o);
    }
    if (k == "Photodetector") {  // Photodetector.jl:49-55
        auto* o = mk_obj(O_PD, mk_rect_flat_mesh(d[0], d[0]));
        o->pd_n = ih[0]; o->pd_lo = -(d[0] / 2); o->pd_hi = d[0] / 2;
        o->field.assign((size_t)ih[0] * ih[0], Cx{0, 0});
        return reg_object(o);
    }
    if (k == "Spotdetector") {  // Spotdetector.jl:39-45
        Mesh* m = mk_rect_flat_mesh(d[0], d[0]);
        m->rotate(V3{0, 0, 1}, kPi);
        auto* o = mk_obj(O_SPOT, m);
        o->sd_hw = d[0] / 2;
        return reg_object(o);
    }
    if (k == "PolarizationFilter") {  // PolarizationFilter.jl:17-29: d = edge_length [, cutoff, J row-major (9)]
        Mesh* m = mk_rect_flat_mesh(d[0], d[0]);
        m->rotate(V3{0, 0, 1}, kPi);
        m->set_new_origin();
        auto* o = mk_obj(O_POLFILTER, m);
        const double xz[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};   // XZBasis(1, 0, 0, 0) = [j11 0 j12; 0 1 0; j21 0 j22]
        if (nd >= 2) o->cutoff = d[1];
        for (int i = 0; i < 9; i++) o->jones[i / 3][i % 3] = nd >= 11 ? d[2 + i] : xz[i];
        return reg_object(o);
    }
    if (k == "PSFDetector") {  // PSFDetector.jl:62-68
        Mesh* m = mk_rect_flat_mesh(d[0], d[0]);
        m->rotate(V3{0, 0, 1}, kPi);
        return reg_object(mk_obj(O_PSF, m));
    }
    if (k == "ObjectGroup") { auto* o = new Object(O_GROUP); for (int i = 0; i < ni; i++) o->parts.push_back(O(ih[i])); return reg_object(o); }
    if (k == "System") {
        auto* s = new System();
        for (int i = 0; i < ni; i++) s->objects.push_back(O(ih[i]));
        s->flatten();
        Entry e; e.system = s; g_reg.push_back(e);
        return (int)g_reg.size() - 1;
    }
    // ---- beams
    if (k == "Beam") {  // pos dir lambda
        auto* b = new Beam(); b->rays.push_back(make_ray(V3{d[0], d[1], d[2]}, V3{d[3], d[4], d[5]}, d[6]));
        Entry e; e.beam = b; g_reg.push_back(e); return (int)g_reg.size() - 1;
    }
    if (k == "PolarizedBeam") {  // pos dir lambda E0(re,im x3)
        Cx E0[3] = {{d[7], d[8]}, {d[9], d[10]}, {d[11], d[12]}};
        auto* b = new Beam(); b->rays.push_back(make_pol_ray(V3{d[0], d[1], d[2]}, V3{d[3], d[4], d[5]}, d[6], E0));
        Entry e; e.beam = b; g_reg.push_back(e); return (int)g_reg.size() - 1;
    }
    if (k == "GaussianBeamlet") {  // pos dir lambda w0 M2 P0 z0 support
        Gauss* g = make_gauss(V3{d[0], d[1], d[2]}, V3{d[3], d[4], d[5]}, d[6], d[7], d[8], d[9], d[10], V3{d[11], d[12], d[13]});
        Entry e; e.gauss = g; g_reg.push_back(e); return (int)g_reg.size() - 1;
    }
    throw std::runtime_error("orc_new: unknown kind " + k);
    ORC_CATCH(-1)
}

// sub-handles: part i of a multi-shape object / shape of an object (registered on demand)
// Mesh(load(path)) for a binary STL (Mesh.jl:48-70 on top of MeshIO's binary STL reader): 80-byte header, uint32 triangle
// count, per triangle 12 little-endian Float32 (normal, 3 vertices) + uint16; vertex 3(i-1)+j = j-th corner of triangle i,
// faces are the consecutive triples, everything scaled by Float32(1e-3) in Float32 arithmetic.
int orc_load_stl(const char* path) {
    ORC_TRY
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("orc_load_stl: cannot open ") + path);
    unsigned char hdr[84];
    if (fread(hdr, 1, 84, f) != 84) { fclose(f); throw std::runtime_error("orc_load_stl: short file"); }
    uint32_t n = (uint32_t)hdr[80] | ((uint32_t)hdr[81] << 8) | ((uint32_t)hdr[82] << 16) | ((uint32_t)hdr[83] << 24);
    auto* m = new Mesh();
    m->f32 = true;
    const float sc = 1e-3f;
    m->scale = (double)sc;
    std::vector<unsigned char> rec(50);
    for (uint32_t i = 0; i < n; i++) {
        if (fread(rec.data(), 1, 50, f) != 50) { fclose(f); delete m; throw std::runtime_error("orc_load_stl: truncated triangle record"); }
        for (int j = 0; j < 3; j++) {
            float c[3];
            std::memcpy(c, rec.data() + 12 + 12 * j, 12);
            m->vertices.push_back({(double)(c[0] * sc), (double)(c[1] * sc), (double)(c[2] * sc)});
        }
        m->faces.push_back({(int)(3 * i), (int)(3 * i + 1), (int)(3 * i + 2)});
    }
    fclose(f);
    return reg_shape(m);
    ORC_CATCH(-1)
}
int orc_part(int h, int i) { ORC_TRY return reg_object(O(h)->parts.at(i)); ORC_CATCH(-1) }
int orc_shape_of(int h) { ORC_TRY return reg_shape(O(h)->shape); ORC_CATCH(-1) }

// kinematics on a shape or object handle
int orc_kin(int h, const char* op, const double* a) {
    ORC_TRY
    std::string k(op);
    Entry& e = g_reg.at(h);
    V3 v{a ? a[0] : 0, a ? a[1] : 0, a ? a[2] : 0};
    if (e.shape) {
        Shape* s = e.shape;
        if (k == "translate3d") s->translate(v);
        else if (k == "translate_to3d") s->translate_to(v);
        else if (k == "rotate3d") s->rotate(v, a[3]);
        else if (k == "xrotate3d") s->rotate(V3{1, 0, 0}, a[0]);
        else if (k == "yrotate3d") s->rotate(V3{0, 1, 0}, a[0]);
        else if (k == "zrotate3d") s->rotate(V3{0, 0, 1}, a[0]);
        else if (k == "align3d") s->align(v);
        else if (k == "reset_translation3d") s->reset_translation();
        else if (k == "reset_rotation3d") s->reset_rotation();
        else if (k == "set_new_origin3d") { auto* m = dynamic_cast<Mesh*>(s); if (!m) throw std::runtime_error("not a mesh"); m->set_new_origin(); }
        else throw std::runtime_error("unknown kinematic op " + k);
        return 0;
    }
    Object* o = O(h);
    if (k == "translate3d") o->translate(v);
    else if (k == "translate_to3d") o->translate_to(v);
    else if (k == "rotate3d") o->rotate(v, a[3]);
    else if (k == "xrotate3d") o->rotate(V3{1, 0, 0}, a[0]);
    else if (k == "yrotate3d") o->rotate(V3{0, 1, 0}, a[0]);
    else if (k == "zrotate3d") o->rotate(V3{0, 0, 1}, a[0]);
    else if (k == "align3d") o->align(v);
    else if (k == "reset_translation3d") o->reset_translation();
    else if (k == "reset_rotation3d") o->reset_rotation();
    else if (k == "set_new_origin3d") { auto* m = dynamic_cast<Mesh*>(o->shape); if (!m) throw std::runtime_error("not a mesh"); m->set_new_origin(); }
    else throw std::runtime_error("unknown kinematic op " + k);
    return 0;
    ORC_CATCH(-1)
}

// pose: pos[3], dir[9] row-major of a shape / object handle
int orc_pose(int h, double* pos, double* dir) {
    ORC_TRY
    Entry& e = g_reg.at(h);
    V3 p; M3 d;
    if (e.shape) { p = e.shape->pos; d = e.shape->dir; } else { p = O(h)->position(); d = O(h)->orientation(); }
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dir[3 * i + j] = d.m[i][j];
    return 0;
    ORC_CATCH(-1)
}

int orc_solve(int hsys, int hbeam, int r_max) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    sys->flatten();
    Entry& e = g_reg.at(hbeam);
    if (e.beam) solve_system(*sys, *e.beam, r_max);
    else if (e.gauss) solve_system(*sys, *e.gauss, r_max);
    else throw std::runtime_error("not a beam");
    return 0;
    ORC_CATCH(-1)
}

// solve_system!(system, beam; r_max, retrace=true) on a beam that may already hold a solution (System.jl:444-461)
int orc_solve_retrace(int hsys, int hbeam, int r_max) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    sys->flatten();
    Entry& e = g_reg.at(hbeam);
    if (e.beam) solve_system(*sys, *e.beam, r_max, true);
    else if (e.gauss) solve_system(*sys, *e.gauss, r_max, true);
    else throw std::runtime_error("not a beam");
    return 0;
    ORC_CATCH(-1)
}

// Beam tree export in BFS (level) order.  Layout per ray (24 doubles):
// pos3 dir3 n lambda t nrm3 obj part E0(6) polarized pad(3)
static void export_beam(System* sys, Beam* root, std::vector<double>& rays, std::vector<int>& beams) {
    std::deque<std::pair<Beam*, int>> q{{root, -1}};
    int idx = 0;
    while (!q.empty()) {
        auto [b, parent] = q.front(); q.pop_front();
        int me = idx++;
        beams.push_back(parent); beams.push_back((int)b->rays.size());
        for (auto& r : b->rays) {
            double rec[24] = {r.pos.x, r.pos.y, r.pos.z, r.dir.x, r.dir.y, r.dir.z, r.n, r.lambda, r.length(),
                              r.hit.n.x, r.hit.n.y, r.hit.n.z,
                              (double)(r.hit.valid ? leaf_index(sys, r.hit.object) : -1),
                              (double)(r.hit.valid ? part_index(r.hit.object, r.hit.shape) : -1),
                              r.E0[0].re, r.E0[0].im, r.E0[1].re, r.E0[1].im, r.E0[2].re, r.E0[2].im,
                              r.polarized ? 1.0 : 0.0, 0, 0, 0};
            rays.insert(rays.end(), rec, rec + 24);
        }
        for (auto* c : b->children) q.push_back({c, me});
    }
}
// returns number of beams; fills up to cap entries.  beams: (parent, nrays) pairs.
int orc_beam_export(int hsys, int hbeam, double* rays, int cap_rays, int* beams, int cap_beams, int* n_rays_out) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    std::vector<double> r; std::vector<int> b;
    export_beam(sys, g_reg.at(hbeam).beam, r, b);
    int nr = (int)r.size() / 24, nb = (int)b.size() / 2;
    *n_rays_out = nr;
    if (rays && nr <= cap_rays) std::memcpy(rays, r.data(), r.size() * sizeof(double));
    if (beams && nb <= cap_beams) std::memcpy(beams, b.data(), b.size() * sizeof(int));
    return nb;
    ORC_CATCH(-1)
}
// Gaussian tree export, BFS order.  beams: (parent, nseg); per beamlet gparams: lambda w0 E0re E0im
// length opl; rays: for each segment chief, waist, divergence records (24 doubles each).
int orc_gauss_export(int hsys, int hg, double* rays, int cap_rays, int* beams, int cap_beams, double* gparams, int* n_rays_out) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    std::vector<double> r, gp; std::vector<int> b;
    std::deque<std::pair<Gauss*, int>> q{{g_reg.at(hg).gauss, -1}};
    int idx = 0;
    while (!q.empty()) {
        auto [g, parent] = q.front(); q.pop_front();
        int me = idx++;
        b.push_back(parent); b.push_back((int)g->chief.rays.size());
        double gpr[6] = {g->lambda, g->w0, g->E0.re, g->E0.im, g->length(), g->opl()};
        gp.insert(gp.end(), gpr, gpr + 6);
        for (size_t s = 0; s < g->chief.rays.size(); s++) {
            Beam* tri[3] = {&g->chief, &g->waist, &g->divergence};
            for (auto* bm : tri) {
                if (s >= bm->rays.size()) { double z[24] = {0}; r.insert(r.end(), z, z + 24); continue; }
                Ray& ry = bm->rays[s];
                double rec[24] = {ry.pos.x, ry.pos.y, ry.pos.z, ry.dir.x, ry.dir.y, ry.dir.z, ry.n, ry.lambda, ry.length(),
                                  ry.hit.n.x, ry.hit.n.y, ry.hit.n.z,
                                  (double)(ry.hit.valid ? leaf_index(sys, ry.hit.object) : -1),
                                  (double)(ry.hit.valid ? part_index(ry.hit.object, ry.hit.shape) : -1), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                r.insert(r.end(), rec, rec + 24);
            }
        }
        for (auto* c : g->children) q.push_back({c, me});
    }
    int nr = (int)r.size() / 24, nb = (int)b.size() / 2;
    *n_rays_out = nr;
    if (rays && nr <= cap_rays) std::memcpy(rays, r.data(), r.size() * sizeof(double));
    if (beams && nb <= cap_beams) { std::memcpy(beams, b.data(), b.size() * sizeof(int)); if (gparams) std::memcpy(gparams, gp.data(), gp.size() * sizeof(double)); }
    return nb;
    ORC_CATCH(-1)
}

int orc_pd_field(int hpd, double* out /* n*n*2, column-major [i,j] interleaved re,im */) {
    ORC_TRY
    Object* pd = O(hpd);
    for (size_t i = 0; i < pd->field.size(); i++) { out[2 * i] = pd->field[i].re; out[2 * i + 1] = pd->field[i].im; }
    return pd->pd_n;
    ORC_CATCH(-1)
}
int orc_pd_empty(int hpd) { ORC_TRY Object* pd = O(hpd); for (auto& f : pd->field) f = Cx{0, 0}; return 0; ORC_CATCH(-1) }
// Photodetector.jl:109-116 + Trapz.jl: optical_power = trapz((x, y), |E|^2/(2 Z))
double orc_pd_power(int hpd) {
    ORC_TRY
    Object* pd = O(hpd);
    int n = pd->pd_n;
    auto coord = [&](int i) { double t = (n == 1) ? 0.0 : (double)i / (double)(n - 1); return (1 - t) * pd->pd_lo + t * pd->pd_hi; };
    // integrate over x (first dim) then y
    std::vector<double> col(n);
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int i = 0; i + 1 < n; i++) {
            double a = abs2(pd->field[(size_t)i + (size_t)n * j]) / (2 * kZvac);
            double b = abs2(pd->field[(size_t)i + 1 + (size_t)n * j]) / (2 * kZvac);
            s += (coord(i + 1) - coord(i)) * (a + b) / 2;
        }
        col[j] = s;
    }
    double p = 0;
    for (int j = 0; j + 1 < n; j++) p += (coord(j + 1) - coord(j)) * (col[j] + col[j + 1]) / 2;
    return p;
    ORC_CATCH(std::nan(""))
}
// ---- PSFDetector (PSFDetector.jl) ------------------------------------------------------------------
// records: 9 doubles per hit (hit xyz, dir xyz, opl, proj, k); returns the number of hits
int orc_psf_data(int hpsf, double* out, int cap) {
    ORC_TRY
    Object* o = O(hpsf);
    int n = (int)o->psf.size();
    for (int i = 0; i < n && i < cap; i++) {
        const PSFData& h = o->psf[i];
        double* r = out + 9 * i;
        r[0] = h.hit.x; r[1] = h.hit.y; r[2] = h.hit.z; r[3] = h.dir.x; r[4] = h.dir.y; r[5] = h.dir.z; r[6] = h.opl; r[7] = h.proj; r[8] = h.k;
    }
    return n;
    ORC_CATCH(-1)
}
int orc_psf_empty(int hpsf) { ORC_TRY O(hpsf)->psf.clear(); return 0; ORC_CATCH(-1) }
// calc_local_pos + calc_local_lims (PSFDetector.jl:91-141); center: 0 = :centroid, 1 = :bbox
static void psf_lims(Object* o, double crop, int center, double lims[4]) {
    const size_t n = o->psf.size();
    std::vector<double> xs(n), zs(n);
    V3 e1 = o->shape->dir.col(0), e3 = o->shape->dir.col(2);
    for (size_t i = 0; i < n; i++) {
        V3 loc = o->psf[i].hit - o->shape->pos;
        xs[i] = dot(loc, e1); zs[i] = dot(loc, e3);
    }
    double x0, z0;
    if (center == 0) {
        double w = 0, sx = 0, sz = 0;
        for (size_t i = 0; i < n; i++) { w += o->psf[i].proj; sx += o->psf[i].proj * xs[i]; sz += o->psf[i].proj * zs[i]; }
        x0 = sx / w; z0 = sz / w;
    } else {
        x0 = (*std::min_element(xs.begin(), xs.end()) + *std::max_element(xs.begin(), xs.end())) / 2;
        z0 = (*std::min_element(zs.begin(), zs.end()) + *std::max_element(zs.begin(), zs.end())) / 2;
    }
    double dx = 0, dz = 0;
    for (size_t i = 0; i < n; i++) { dx = std::max(dx, std::fabs(xs[i] - x0)); dz = std::max(dz, std::fabs(zs[i] - z0)); }
    lims[0] = x0 - dx * crop; lims[1] = x0 + dx * crop; lims[2] = z0 - dz * crop; lims[3] = z0 + dz * crop;
}
int orc_psf_lims(int hpsf, double crop, int center, double* lims) {
    ORC_TRY
    Object* o = O(hpsf);
    if (o->psf.empty()) throw std::runtime_error("PSFDetector holds no data");
    psf_lims(o, crop, center, lims);
    return 0;
    ORC_CATCH(-1)
}
// intensity(psf; n, crop_factor, center, x_min, x_max, z_min, z_max, x0_shift, z0_shift) (PSFDetector.jl:190-237).
// lims_in: NULL or 4 doubles (Inf entries = automatic, pairwise like the reference); out: xs[n], zs[n], I[n*n]
// column-major [i(x), j(z)].
int orc_psf_intensity(int hpsf, int n, double crop, int center, const double* lims_in, double x0_shift, double z0_shift,
                      double* xs, double* zs, double* I) {
    ORC_TRY
    Object* o = O(hpsf);
    if (o->psf.empty()) throw std::runtime_error("PSFDetector holds no data");
    double lims[4];
    psf_lims(o, crop, center, lims);
    if (lims_in) {
        if (lims_in[0] != kInf && lims_in[1] != kInf) { lims[0] = lims_in[0]; lims[1] = lims_in[1]; }
        if (lims_in[2] != kInf && lims_in[3] != kInf) { lims[2] = lims_in[2]; lims[3] = lims_in[3]; }
    }
    auto lin = [&](int i, double lo, double hi) {   // LinRange getindex (Base.lerpi)
        double t = (n == 1) ? 0.0 : (double)i / (double)(n - 1);
        return (1 - t) * lo + t * hi;
    };
    for (int i = 0; i < n; i++) { xs[i] = lin(i, lims[0], lims[1]) + x0_shift; zs[i] = lin(i, lims[2], lims[3]) + z0_shift; }
    V3 e1 = o->shape->dir.col(0), e2 = o->shape->dir.col(2), org = o->shape->pos;
#pragma omp parallel for schedule(static)   // Threads.@threads over j, PSFDetector.jl:220
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) {
            V3 p = org + xs[i] * e1 + zs[j] * e2;
            Cx acc{0, 0};
            for (const PSFData& h : o->psf) {
                double l = dot(p - h.hit, h.dir);
                acc = acc + h.proj * cis(h.k * (h.opl + l));
            }
            I[(size_t)i + (size_t)n * j] = abs2(acc);
        }
    }
    return 0;
    ORC_CATCH(-1)
}

int orc_spots(int hsd, double* out, int cap) {
    ORC_TRY
    Object* sd = O(hsd);
    int n = (int)sd->spots.size();
    for (int i = 0; i < n && i < cap; i++) { out[2 * i] = sd->spots[i][0]; out[2 * i + 1] = sd->spots[i][1]; }
    return n;
    ORC_CATCH(-1)
}
int orc_spots_empty(int hsd) { ORC_TRY O(hsd)->spots.clear(); return 0; ORC_CATCH(-1) }

// ---------------------------------------------------------------------------------------------
// Bulk entry points (CPU baseline + parity at scale).  Rays are independent; `nthreads` > 1 uses
// the "threaded driver" of BASELINE.md B2 (the reference's own solve_system! over a beam group is
// serial, System.jl:463-468).  Detector state is written per ray into the output arrays instead
// of Spotdetector.data so the loop is race free; the tracing itself is the reference algorithm.
//
// out_seg: per ray up to max_seg segment records of 16 doubles: pos3 dir3 n t nrm3 obj part pad3
// out_nseg: number of segments per ray; out_spot: (x,z) of the hit on Spotdetector `hsd` or NaN.
// returns total number of interactions (hits that reached interact3d).
long long orc_bulk_trace_rays(int hsys, int n, const double* pos, const double* dir, const double* lambda,
                              int r_max, int nthreads, int max_seg, double* out_seg, int* out_nseg,
                              int hsd, double* out_spot) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    sys->flatten();
    Object* sd = hsd >= 0 ? O(hsd) : nullptr;
    long long total = 0;
    if (nthreads < 1) nthreads = 1;
    std::string err;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64) reduction(+ : total)
    for (int i = 0; i < n; i++) {
        try {
            Beam b;
            {   // `dir` is the direction of an already constructed Ray (normalised once by Ray(pos, dir, lambda))
                Ray r0; r0.pos = V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}; r0.dir = V3{dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
                r0.lambda = lambda[i]; r0.n = 1.0;
                b.rays.push_back(r0);
            }
            // Spotdetector hits are recomputed below from the last ray, so give each thread a scratch copy
            // of nothing: interact3d(O_SPOT) pushes into sd->spots, which would race -> trace manually.
            BeamInteraction interaction;
            while ((int)b.rays.size() < r_max) {
                Ray& ray = b.rays.back();
                tracing_step(*sys, ray, interaction.valid ? interaction.hint : Hint{});
                if (!ray.hit.valid) break;
                total += 1;
                if (ray.hit.object->kind == O_SPOT) {
                    V3 hp = ray.pos + ray.length() * ray.dir;
                    V3 loc = hp - ray.hit.object->shape->pos;
                    if (out_spot && ray.hit.object == sd) {
                        out_spot[2 * i] = dot(loc, sd->shape->dir.col(0));
                        out_spot[2 * i + 1] = dot(loc, sd->shape->dir.col(2));
                    }
                    break;
                }
                if (ray.hit.object->kind == O_THIN_BS || ray.hit.object->kind == O_PLATE_BS || ray.hit.object->kind == O_CUBE_BS)
                    throw std::runtime_error("orc_bulk_trace_rays: branching systems are not supported in bulk mode");
                interaction = interact3d(*sys, ray.hit.object, b, ray);
                if (!interaction.valid) break;
                b.rays.push_back(interaction.ray);
            }
            if (out_nseg) out_nseg[i] = (int)b.rays.size();
            if (out_seg) {
                for (int s = 0; s < (int)b.rays.size() && s < max_seg; s++) {
                    Ray& r = b.rays[s];
                    double* o = out_seg + ((size_t)i * max_seg + s) * 16;
                    o[0] = r.pos.x; o[1] = r.pos.y; o[2] = r.pos.z; o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z;
                    o[6] = r.n; o[7] = r.length(); o[8] = r.hit.n.x; o[9] = r.hit.n.y; o[10] = r.hit.n.z;
                    o[11] = r.hit.valid ? leaf_index(sys, r.hit.object) : -1;
                    o[12] = r.hit.valid ? part_index(r.hit.object, r.hit.shape) : -1;
                }
            }
        } catch (const std::exception& e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty()) throw std::runtime_error(err);
    return total;
    ORC_CATCH(-1)
}

// Photodetector accumulation of many independent root beamlets (each traced through `hsys`), the
// pixel loop threaded over rows like Photodetector.jl:87.  Returns ray-interactions (3 per beamlet hit).
// g: per beamlet 14 doubles as in orc_new("GaussianBeamlet").
long long orc_bulk_trace_beamlets(int hsys, int n, const double* g, int r_max, int nthreads) {
    ORC_TRY
    System* sys = g_reg.at(hsys).system;
    sys->flatten();
    long long total = 0;
    omp_set_num_threads(nthreads < 1 ? 1 : nthreads);
    for (int i = 0; i < n; i++) {
        const double* d = g + 14 * (size_t)i;
        Gauss* gb = make_gauss(V3{d[0], d[1], d[2]}, V3{d[3], d[4], d[5]}, d[6], d[7], d[8], d[9], d[10], V3{d[11], d[12], d[13]});
        solve_system(*sys, *gb, r_max);
        std::deque<Gauss*> q{gb};
        while (!q.empty()) {
            Gauss* c = q.front(); q.pop_front();
            for (auto& r : c->chief.rays) if (r.hit.valid) total += 3;
            for (auto* ch : c->children) q.push_back(ch);
        }
        delete gb;
    }
    return total;
    ORC_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------
// Function-level access for known-answer tests
int orc_eval(const char* fn, const int* ih, int ni, const double* a, int na, double* out) {
    ORC_TRY
    std::string k(fn);
    (void)ni; (void)na;
    if (k == "sdf") { out[0] = SD(ih[0])->sdf(V3{a[0], a[1], a[2]}); return 1; }
    if (k == "normal3d") { V3 n = SD(ih[0])->normal3d(V3{a[0], a[1], a[2]}); out[0] = n.x; out[1] = n.y; out[2] = n.z; return 3; }
    if (k == "gradient_ad") {
        P3<Dual> q{{a[0], {1, 0, 0}}, {a[1], {0, 1, 0}}, {a[2], {0, 0, 1}}};
        Dual d = SD(ih[0])->sdf(q); out[0] = d.v; out[1] = d.p[0]; out[2] = d.p[1]; out[3] = d.p[2]; return 4;
    }
    if (k == "numeric_gradient") { V3 n = SD(ih[0])->numeric_gradient(V3{a[0], a[1], a[2]}); out[0] = n.x; out[1] = n.y; out[2] = n.z; return 3; }
    if (k == "world_to_sdf") { auto p = SD(ih[0])->w2s(P3<double>{a[0], a[1], a[2]}); out[0] = p.x; out[1] = p.y; out[2] = p.z; return 3; }
    if (k == "intersect3d_shape") {  // -> valid t nx ny nz
        Hit h = S(ih[0])->intersect(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]});
        out[0] = h.valid; out[1] = h.t; out[2] = h.n.x; out[3] = h.n.y; out[4] = h.n.z; return 5;
    }
    if (k == "intersect3d_object") {
        Ray r; r.pos = V3{a[0], a[1], a[2]}; r.dir = V3{a[3], a[4], a[5]};
        Object* o = O(ih[0]);
        Hit h = o->intersect(r);
        out[0] = h.valid; out[1] = h.t; out[2] = h.n.x; out[3] = h.n.y; out[4] = h.n.z; out[5] = part_index(o, h.shape); return 6;
    }
    if (k == "moeller_trumbore") { out[0] = Mesh::moeller_trumbore(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}, V3{a[6], a[7], a[8]}, V3{a[9], a[10], a[11]}, V3{a[12], a[13], a[14]}); return 1; }
    if (k == "mesh_counts") { Mesh* m = M(ih[0]); out[0] = (double)m->vertices.size(); out[1] = (double)m->faces.size(); out[2] = m->f32; out[3] = m->scale; return 4; }
    if (k == "mesh_faces") { Mesh* m = M(ih[0]); for (size_t i = 0; i < m->faces.size(); i++) for (int j = 0; j < 3; j++) out[3 * i + j] = m->faces[i][j]; return 3 * (int)m->faces.size(); }
    if (k == "mesh_vertices") { Mesh* m = M(ih[0]); for (size_t i = 0; i < m->vertices.size(); i++) { out[3 * i] = m->vertices[i].x; out[3 * i + 1] = m->vertices[i].y; out[3 * i + 2] = m->vertices[i].z; } return 3 * (int)m->vertices.size(); }
    if (k == "thickness_shape") { out[0] = S(ih[0])->thickness(); return 1; }
    if (k == "thickness_object") { out[0] = O(ih[0])->thickness(); return 1; }
    if (k == "reflection3d") { V3 r = reflection3d(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; return 3; }
    if (k == "refraction3d") { bool tir; V3 r = refraction3d(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}, a[6], a[7], tir); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = tir; return 4; }
    if (k == "fresnel_coefficients") { Cx rs, rp, ts, tp; fresnel_coefficients(a[0], a[1], rs, rp, ts, tp); double o[8] = {rs.re, rs.im, rp.re, rp.im, ts.re, ts.im, tp.re, tp.im}; std::memcpy(out, o, sizeof(o)); return 8; }
    if (k == "rotate3d") { M3 R = rotate3d(V3{a[0], a[1], a[2]}, a[3]); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[3 * i + j] = R.m[i][j]; return 9; }
    if (k == "align3d") { M3 R = align3d(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[3 * i + j] = R.m[i][j]; return 9; }
    if (k == "angle3d") { out[0] = angle3d(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}); return 1; }
    if (k == "polarization_matrix_apply") {  // in3 out3 n3 j11(re,im) j22(re,im) E(6) -> E'(6)
        Cx Ein[3] = {{a[13], a[14]}, {a[15], a[16]}, {a[17], a[18]}}, Eo[3];
        calculate_global_E0(V3{a[0], a[1], a[2]}, V3{a[3], a[4], a[5]}, V3{a[6], a[7], a[8]}, Cx{a[9], a[10]}, Cx{a[11], a[12]}, Ein, Eo);
        for (int i = 0; i < 3; i++) { out[2 * i] = Eo[i].re; out[2 * i + 1] = Eo[i].im; }
        return 6;
    }
    if (k == "gauss_parameters") { Gauss* g = g_reg.at(ih[0]).gauss; gauss_parameters(*g, a[0], out[0], out[1], out[2], out[3]); return 4; }
    if (k == "gauss_electric_field") { Gauss* g = g_reg.at(ih[0]).gauss; Cx e = electric_field(*g, a[0], a[1]); out[0] = e.re; out[1] = e.im; return 2; }
    // electric_field!(gauss, electric_field(gauss) * (a[0] + i a[1])) (Gaussian.jl: electric_field!): the E0 of the root beamlet is
    // multiplied; a following solve_system!(...; retrace = true) carries it through the stored tree (test/runtests.jl:2846-2849)
    if (k == "gauss_scale_E0") { Gauss* g = g_reg.at(ih[0]).gauss; g->E0 = g->E0 * Cx{a[0], a[1]}; out[0] = g->E0.re; out[1] = g->E0.im; return 2; }
    if (k == "gauss_length") { Gauss* g = g_reg.at(ih[0]).gauss; out[0] = g->length(); out[1] = g->opl(); return 2; }
    if (k == "refractive_index") { out[0] = RI(ih[0])(a[0]); return 1; }
    throw std::runtime_error("orc_eval: unknown fn " + k);
    ORC_CATCH(-1)
}

}  // extern "C"
