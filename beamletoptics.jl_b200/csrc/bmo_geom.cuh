// bmo_geom.cuh -- device-side geometry of the tracer: SDF evaluation (value and dual), sphere
// tracing, Moeller-Trumbore behind a BVH, per-object closest hit, trace_all / trace_one.
// Citations are relative to /root/reference/src.
#pragma once
#include "../../include/bmo.h"
#include "bmo_math.cuh"

namespace bmo {

constexpr double eps_srf = 1e-9;   // SDFs/AbstractSDF.jl:1
constexpr double eps_ray = 1e-10;  // :2
constexpr double eps_ins = 1.0;    // :3
constexpr int kMarchIter = 1000;   // :105,135

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
    const double* bounds;    // [n_poses][n_parts][4]
    const double* det_pose;  // [n_poses][n_objects][12] pos(3) dir(9 row-major)
    const double* lambdas;   // [n_lambda]
    int32_t n_prims, n_parts, n_objects, n_meshes, n_lambda, n_poses, zr, pad;
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
// min(maximum(d), 0) + norm(max.(d, 0))   (SphericalLensSDF.jl:64)
template <class T> BMO_D T cyl_(T d1, T d2, int zr) {
    return min_(max_(d1, d2), 0.0) + norm2_(max_(d1, 0.0), max_(d2, 0.0), zr);
}
template <class T> BMO_NI T prim_eval(const bmo_prim& pr, P3<T> q, int zr, Stats& st) {
    st.sdf++;
    P3<T> p = w2s(pr, q);
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
template <class T> BMO_D T member_eval(const bmo_prim* prims, int i, P3<T> q, int zr, Stats& st) {
    const bmo_prim& pr = prims[i];
    if (pr.type != BMO_PRIM_MENISCUS) return prim_eval(pr, q, zr, st);
    P3<T> p = w2s(pr, q);
    T cv = prim_eval(prims[i + 1], p, zr, st);
    T cy = prim_eval(prims[i + 2], p, zr, st);
    T cc = prim_eval(prims[i + 3], p, zr, st);
    return max_(min_(cv, cy), -cc);
}
BMO_D int member_advance(const bmo_prim* prims, int i) { return prims[i].type == BMO_PRIM_MENISCUS ? 4 : 1; }

// UnionSDF.jl:53-56  minimum over members (left fold)
BMO_NI double shape_sdf(const bmo_prim* prims, int first, int count, V3 p, int zr, Stats& st) {
    P3<double> q; q.x = p.x; q.y = p.y; q.z = p.z;
    double m = member_eval(prims, first, q, zr, st);
    for (int i = first + member_advance(prims, first); i < first + count; i += member_advance(prims, i))
        m = jl_min(m, member_eval(prims, i, q, zr, st));
    return m;
}
// AbstractSDF.jl:79-95 + UnionSDF.jl:86-91: normal of the arg-min member; ForwardDiff gradient,
// central differences (eps = 1e-8) if any component of the normalised gradient is NaN.
BMO_NI V3 shape_normal(const bmo_prim* prims, int first, int count, V3 p, int zr, Stats& st) {
    P3<double> q; q.x = p.x; q.y = p.y; q.z = p.z;
    int idx = first;
    if (count > member_advance(prims, first)) {
        double m = member_eval(prims, first, q, zr, st);
        for (int i = first + member_advance(prims, first); i < first + count; i += member_advance(prims, i)) {
            double v = member_eval(prims, i, q, zr, st);
            if (v < m) { m = v; idx = i; }
        }
    }
    P3<Dual> qd;
    qd.x = mkd(p.x, 1, 0, 0); qd.y = mkd(p.y, 0, 1, 0); qd.z = mkd(p.z, 0, 0, 1);
    Dual g = member_eval(prims, idx, qd, zr, st);
    V3 n = normalize(mk3(g.p0, g.p1, g.p2));
    if (!isnan(n.x) && !isnan(n.y) && !isnan(n.z)) return n;
    const double e = 1e-8;
    P3<double> a, b;
    V3 gr;
    a = q; b = q; a.x = p.x + e; b.x = p.x - e;
    gr.x = member_eval(prims, idx, a, zr, st) - member_eval(prims, idx, b, zr, st);
    a = q; b = q; a.y = p.y + e; b.y = p.y - e;
    gr.y = member_eval(prims, idx, a, zr, st) - member_eval(prims, idx, b, zr, st);
    a = q; b = q; a.z = p.z + e; b.z = p.z - e;
    gr.z = member_eval(prims, idx, a, zr, st) - member_eval(prims, idx, b, zr, st);
    return normalize(gr);
}

// AbstractSDF.jl:102-125.  The bounding-sphere test is result-identical: once the march point is
// outside the (inflated) bounding sphere and moving away, every later point p + s*d (s >= 0) stays
// outside it, so sdf >= margin > eps_ray for the rest of the reference's 1000 iterations => miss.
BMO_NI bool march_outside(const bmo_prim* prims, int first, int count, const double* bnd, V3 p, V3 d, int zr, Stats& st,
                         double& t, V3& n) {
    double dist = shape_sdf(prims, first, count, p, zr, st);
    double t0 = dist;
    const double R2 = bnd[3] * bnd[3];
    for (int i = 0; i < kMarchIter; i++) {
        p = p + dist * d;
        dist = shape_sdf(prims, first, count, p, zr, st);
        t0 += dist;
        if (dist < eps_ray) {
            n = shape_normal(prims, first, count, p, zr, st);
            t = t0;
            return true;
        }
        V3 v = mk3(p.x - bnd[0], p.y - bnd[1], p.z - bnd[2]);
        if (dot(v, v) > R2 && dot(v, d) > 0.0) return false;
        if (!(dist == dist)) return false;  // NaN can never satisfy dist < eps_ray again
    }
    return false;
}
// AbstractSDF.jl:132-159
BMO_D bool march_inside(const bmo_prim* prims, int first, int count, const double* bnd, V3 p, V3 d, int zr, Stats& st,
                        double& t, V3& n) {
    double t0 = 0;
    for (int i = 0; i < kMarchIter; i++) {
        p = p + eps_ins * d;
        t0 += eps_ins;
        double dist = shape_sdf(prims, first, count, p, zr, st);
        if (dist > 0) {
            double tt;
            if (!march_outside(prims, first, count, bnd, p, -d, zr, st, tt, n)) return false;
            t = t0 - tt;
            return true;
        }
    }
    return false;
}
// AbstractSDF.jl:166-181
BMO_NI bool sdf_intersect(const bmo_prim* prims, int first, int count, const double* bnd, V3 pos, V3 dir, int zr, Stats& st,
                         double& t, V3& n) {
    // Guaranteed miss: origin outside the bounding sphere and the line never enters it.
    {
        V3 v = mk3(pos.x - bnd[0], pos.y - bnd[1], pos.z - bnd[2]);
        double cc = dot(v, v) - bnd[3] * bnd[3];
        if (cc > 0.0) {
            double b = dot(v, dir);
            if (b >= 0.0) return false;
            if (b * b - dot(dir, dir) * cc < 0.0) return false;
        }
    }
    double s0 = shape_sdf(prims, first, count, pos, zr, st);
    if (s0 > eps_srf) return march_outside(prims, first, count, bnd, pos, dir, zr, st, t, n);
    V3 nn = shape_normal(prims, first, count, pos, zr, st);
    if (dot(dir, nn) <= 0) return march_inside(prims, first, count, bnd, pos, dir, zr, st, t, n);
    return false;
}

// ---- meshes -----------------------------------------------------------------------------------
BMO_D V3 load_vertex(const double* verts, int64_t i) { return mk3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]); }
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
BMO_NI bool mesh_intersect(const SysView& S, int mesh_id, int pose, V3 pos, V3 dir, Stats& st, double& t, V3& n) {
    const MeshView mv = S.meshes[mesh_id];
    const double* verts = S.vertices + 3 * ((int64_t)pose * S.n_vertices + mv.first_vertex);
    const int32_t* faces = S.faces + 3 * mv.first_face;
    double t0 = INFINITY;
    int64_t fid = -1;
    if (mv.n_nodes == 0 || S.n_poses > 1) {  // small meshes (and posed sweeps): reference order
        for (int64_t i = 0; i < mv.n_faces; i++) {
            st.tri++;
            double tt = moeller_trumbore(load_vertex(verts, faces[3 * i]), load_vertex(verts, faces[3 * i + 1]),
                                         load_vertex(verts, faces[3 * i + 2]), pos, dir, mv.f32);
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
                    double tt = moeller_trumbore(load_vertex(verts, faces[3 * f]), load_vertex(verts, faces[3 * f + 1]),
                                                 load_vertex(verts, faces[3 * f + 2]), pos, dir, mv.f32);
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
    n = face_normal(load_vertex(verts, faces[3 * fid]), load_vertex(verts, faces[3 * fid + 1]), load_vertex(verts, faces[3 * fid + 2]), mv.f32);
    return true;
}

// ---- shapes, objects, system --------------------------------------------------------------------
struct TraceCtx {
    const SysView* S;
    const bmo_prim* prims;  // prim table of this ray's pose (shared memory copy when n_poses == 1)
    int pose;
};
// intersect3d(shape, ray)
BMO_NI bool part_intersect(const TraceCtx& C, int part, V3 pos, V3 dir, Stats& st, double& t, V3& n) {
    const bmo_part& pt = C.S->parts[part];
    if (pt.shape_kind == BMO_SHAPE_SDF) {
        const double* bnd = C.S->bounds + 4 * ((int64_t)C.pose * C.S->n_parts + part);
        return sdf_intersect(C.prims, pt.first, pt.count, bnd, pos, dir, C.S->zr, st, t, n);
    }
    return mesh_intersect(*C.S, pt.first, C.pose, pos, dir, st, t, n);
}
// intersect3d(object, ray): AbstractRay.jl:118-155; PlateBeamsplitter.jl:160-187
BMO_D Hit object_intersect(const TraceCtx& C, int obj, V3 pos, V3 dir, Stats& st) {
    const bmo_object& ob = C.S->objects[obj];
    Hit best; best.part = -1; best.t = INFINITY; best.n = mk3(0, 0, 0);
    if (ob.kind == BMO_OBJ_PLATE_BS) {
        double tc, ts; V3 nc, ns;
        bool hc = part_intersect(C, ob.first_part + 1, pos, dir, st, tc, nc);
        bool hs = part_intersect(C, ob.first_part, pos, dir, st, ts, ns);
        if (!hc && !hs) return best;
        bool coat = !hs ? true : (!hc ? false : (jl_isapprox(tc, ts) ? true : (tc < ts)));
        if (coat) { best.t = tc; best.n = nc; best.part = ob.first_part + 1; }
        else { best.t = ts; best.n = ns; best.part = ob.first_part; }
        return best;
    }
    for (int k = 0; k < ob.n_parts; k++) {
        double t; V3 n;
        if (!part_intersect(C, ob.first_part + k, pos, dir, st, t, n)) continue;
        if (best.part < 0 || t < best.t) { best.t = t; best.n = n; best.part = ob.first_part + k; }
    }
    return best;
}
// System.jl:57-72
BMO_D Hit trace_all(const TraceCtx& C, V3 pos, V3 dir, Stats& st) {
    Hit res; res.part = -1; res.t = INFINITY; res.n = mk3(0, 0, 0);
    for (int o = 0; o < C.S->n_objects; o++) {
        Hit h = object_intersect(C, o, pos, dir, st);
        if (h.part < 0) continue;
        if (res.part < 0 || h.t < res.t) res = h;
    }
    return res;
}
// System.jl:74-110: the hinted shape is accepted without comparing against other objects
BMO_D Hit tracing_step(const TraceCtx& C, V3 pos, V3 dir, int hint_part, Stats& st) {
    if (hint_part >= 0) {
        Hit h; h.part = -1; h.t = INFINITY; h.n = mk3(0, 0, 0);
        if (part_intersect(C, hint_part, pos, dir, st, h.t, h.n)) { h.part = hint_part; return h; }
    }
    return trace_all(C, pos, dir, st);
}

}  // namespace bmo
