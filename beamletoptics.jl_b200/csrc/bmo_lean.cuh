// bmo_lean.cuh -- tracing_step! for LEAN systems (every SDF part a union of at most 4 plain primitives, every mesh small:
// bmo_sys::all_lean), restated for the register file of sm_100a.
//
// tracing_step<RK, false> (bmo_geom.cuh) keeps everything in registers and lets ptxas spill: ray origin, reciprocal
// direction, the best hit of the system and of the current object with their normals, the march state and the union loop
// add up to well over the 64-72 registers that 7-8 resident blocks per SM allow, and the spill code lands in the march
// loop (ncu source page of the C2 waves, profiles/r02g_*: 20 % of the executed instructions were local loads / stores, and
// the dirty stack lines quadrupled the DRAM writes of the kernel).  Here the cold state is placed by hand:
//   * per-thread scratch in shared memory ([LS_SLOTS][128] doubles per block, one column per thread): ray origin,
//     reciprocal direction, march end point of the best hit of the current object and of the system;
//   * hit normals are DEFERRED: the march returns (t, end point, arg-min member); the normal -- one dual-number evaluation --
//     is formed once, for the hit that tracing_step! returns.  The reference evaluates normal3d for every shape that is hit
//     (AbstractSDF.jl:117) and then keeps the closest (System.jl:62-67): the normals of the losers are never looked at, so
//     skipping them is result-identical;
//   * the bounding sphere of the shape is read where it is used (two places per march step at most) instead of living in
//     8 registers across the union loop.
// Semantics, operation order and rounding are those of tracing_step / sdf_intersect_t, which stay the reference for this file
// (same tests: every bitwise GPU test of a lens system runs through here).
#pragma once
#include "bmo_geom.cuh"

namespace bmo {

enum { LS_POS = 0, LS_INV = 3, LS_DIR = 6, LS_OB = 9, LS_RES = 12, LS_CAND = 15, LS_T = 18, LS_SLOTS = 19 };
constexpr int LB_ROWS = 5;    // member bounds [4] + the path-length factor of the ray
// shared-memory address of p, formed once (volatile: ptxas would otherwise re-derive it at every use)
BMO_D unsigned smem_u32(const void* p) { unsigned r; asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(r) : "l"(p)); return r; }
BMO_D double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
BMO_D void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }
BMO_D unsigned ls_slot(unsigned sc, int slot) { return sc + 8u * LB_STRIDE * slot; }
BMO_D V3 ls_load3(unsigned sc, int slot) { return mk3(lds_f64(ls_slot(sc, slot)), lds_f64(ls_slot(sc, slot + 1)), lds_f64(ls_slot(sc, slot + 2))); }
BMO_D void ls_store3(unsigned sc, int slot, V3 v) { sts_f64(ls_slot(sc, slot), v.x); sts_f64(ls_slot(sc, slot + 1), v.y); sts_f64(ls_slot(sc, slot + 2), v.z); }
// a table entry that must not be hoisted into a register for the whole march (generic address: shared or global tables)
BMO_D double ld_tab(const double* p) { double v; asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

// intersect3d(::AbstractSDF, ray) without the hit normal: t (scratch slot LS_T) and march end point (LS_CAND) of a hit.
// State machine, bounding-sphere arguments and operation order: sdf_intersect_t (bmo_geom.cuh).
// What the march keeps in registers is the march point, the last distance, one packed state word and the union loop; the ray
// direction, the running length t0, the bounding sphere and the path-length factor are read from shared memory where they are
// used (once per march step), because the primitive evaluation inside the union loop needs the registers.
// Returns (primitive evaluations << 4) | (arg-min member of the hit + 1), low nibble 0 on a miss.
// state word: bits 0-1 mode, bit 2 marching backwards, bits 3-15 iterations, bits 16-31 inside steps taken
BMO_D unsigned lean_march(const bmo_prim* prims, int first, int count, int zr, const double* bnd, unsigned sc, unsigned lb_addr) {
    enum { INIT = 0u, IN = 1u, OUT = 2u, BACK = 4u, IT1 = 8u, NIN1 = 1u << 16 };
    unsigned nsdf = 0;
    V3 p = ls_load3(sc, LS_POS);
    {   // guaranteed miss: origin outside the bounding sphere and the line never enters it
        const V3 dir = ls_load3(sc, LS_DIR);
        const double R = ld_tab(bnd + 3);
        const V3 v = mk3(p.x - ld_tab(bnd), p.y - ld_tab(bnd + 1), p.z - ld_tab(bnd + 2));
        const double cc = dot(v, v) - R * R;
        if (cc > 0.0) {
            const double b = dot(v, dir);
            if (b >= 0.0) return 0u;
            if (b * b - dot(dir, dir) * cc < 0.0) return 0u;
        }
        // |d| <= max(1, |d|^2): how far the march point moves per unit of step (directions are unit up to rounding)
        sts_f32(lb_addr + 4u * LB_STRIDE * 4, __double2float_ru(fmax(1.0, dot(dir, dir)) * (1.0 + 1e-9)));
    }
    unsigned stw = INIT;
    SdfShape sh;
    sh.prims = prims; sh.first = first; sh.count = count; sh.zr = zr;
    LeanBounds lb; lb.addr = lb_addr; lb.reset();
    for (;;) {
        int idx;
        const double dist = shape_eval<false>(sh, p, nsdf, idx, lb);
        const unsigned mode = stw & 3u;
        if (mode == OUT) {
            const double t0 = lds_f64(ls_slot(sc, LS_T)) + dist;
            if (!(dist < eps_ray)) {
                const V3 dir = ls_load3(sc, LS_DIR);
                const double R = ld_tab(bnd + 3);
                const V3 v = mk3(p.x - ld_tab(bnd), p.y - ld_tab(bnd + 1), p.z - ld_tab(bnd + 2));
                const double vd = dot(v, dir);
                const bool back = (stw & BACK) != 0;
                if (dot(v, v) > R * R && (back ? vd < 0.0 : vd > 0.0)) return nsdf << 4;
                stw += IT1;
                if (((stw >> 3) & 0x1fffu) >= (unsigned)kMarchIter || !(dist == dist)) return nsdf << 4;   // NaN can never satisfy dist < eps_ray again
                sts_f64(ls_slot(sc, LS_T), t0);
                p = p + (back ? -dist : dist) * dir; lb.moved_f(dist, lds_f32(lb_addr + 4u * LB_STRIDE * 4));
                continue;
            }
            sts_f64(ls_slot(sc, LS_T), (stw & BACK) ? (double)(stw >> 16) - t0 : t0);
            ls_store3(sc, LS_CAND, p);
            return (nsdf << 4) | (unsigned)(idx - first + 1);
        } else if (mode == INIT) {
            if (dist > eps_srf) {
                stw = OUT; sts_f64(ls_slot(sc, LS_T), dist);
                p = p + dist * ls_load3(sc, LS_DIR); lb.moved_f(dist, lds_f32(lb_addr + 4u * LB_STRIDE * 4));
                continue;
            }
        } else {  // IN
            if (dist > 0) {
                stw = (stw & 0xffff0000u) | OUT | BACK; sts_f64(ls_slot(sc, LS_T), dist);
                p = p + (-dist) * ls_load3(sc, LS_DIR); lb.moved_f(dist, lds_f32(lb_addr + 4u * LB_STRIDE * 4));
                continue;
            }
            stw += IT1;
            if (((stw >> 3) & 0x1fffu) >= (unsigned)kMarchIter) return nsdf << 4;
            p = p + eps_ins * ls_load3(sc, LS_DIR); stw += NIN1; lb.moved_f(eps_ins, lds_f32(lb_addr + 4u * LB_STRIDE * 4));
            continue;
        }
        // INIT on the surface: the normal decides between "heading out" (miss) and the inside march
        Stats tmp; tmp.sdf = 0; tmp.tri = 0;
        const V3 n = member_normal<false>(prims, idx, p, zr, tmp);
        nsdf += tmp.sdf;
        const V3 dir = ls_load3(sc, LS_DIR);
        if (!(dot(dir, n) <= 0)) return nsdf << 4;
        stw = IN | NIN1;
        p = p + eps_ins * dir; lb.moved_f(eps_ins, lds_f32(lb_addr + 4u * LB_STRIDE * 4));
    }
}

// Mesh.jl:244-267 for small meshes, face normal deferred: t and the face that was hit
BMO_NI bool mesh_hit_small(const MeshTabs S, int mesh_id, int pose, V3 pos, V3 dir, double& t, int& fid_out, unsigned& ntri) {
    const MeshView mv = S.meshes[mesh_id];
    const double* verts = S.vertices + 3 * ((int64_t)pose * S.n_vertices + mv.first_vertex);
    const int32_t* faces = S.faces + 3 * mv.first_face;
    double t0 = INFINITY;
    int fid = -1;
    const int nf = (int)mv.n_faces;
    for (int i = 0; i < nf; i++) {
        const double tt = moeller_trumbore(load_vertex(verts, __ldg(faces + 3 * i)), load_vertex(verts, __ldg(faces + 3 * i + 1)),
                                           load_vertex(verts, __ldg(faces + 3 * i + 2)), pos, dir, mv.f32);
        if (tt < t0) { t0 = tt; fid = i; }
    }
    ntri = (unsigned)nf;
    if (fid < 0) return false;
    t = t0; fid_out = fid;
    return true;
}
BMO_NI V3 mesh_face_normal(const MeshTabs S, int mesh_id, int pose, int fid) {
    const MeshView mv = S.meshes[mesh_id];
    const double* verts = S.vertices + 3 * ((int64_t)pose * S.n_vertices + mv.first_vertex);
    const int32_t* faces = S.faces + 3 * mv.first_face;
    return face_normal(load_vertex(verts, __ldg(faces + 3 * fid)), load_vertex(verts, __ldg(faces + 3 * fid + 1)), load_vertex(verts, __ldg(faces + 3 * fid + 2)), mv.f32);
}

// slab test with the origin and the reciprocal direction read from the scratch column (box_may_hit, bmo_geom.cuh)
BMO_D bool box_may_hit_ls(const double* bx, unsigned sc, V3 d, double t_best) {
    double tmin = 0.0, tmax = INFINITY;
    const double dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double o = lds_f64(ls_slot(sc, LS_POS + k));
        if (dd[k] == 0.0) {
            if (o < bx[k] || o > bx[3 + k]) return false;
        } else {
            const double inv = lds_f64(ls_slot(sc, LS_INV + k));
            double t1 = (bx[k] - o) * inv, t2 = (bx[3 + k] - o) * inv;
            if (t1 > t2) { const double tt = t1; t1 = t2; t2 = tt; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
        }
    }
    return tmin <= tmax * (1 + 1e-9) + 1e-12 && tmin * (1 - 1e-9) - 1e-9 <= t_best;
}

// One 64-bit word per part, made at staging time from bmo_part / bmo_object so that the part loop needs one shared-memory load
// where it used to chase parts[part].object -> objects[obj].kind through global memory:
//   low word:  bits 0-15 object, bit 16 the object is a plate beamsplitter, bit 17 the part is its coating,
//              bit 18 SDF shape (else mesh), bits 20-23 members of the union
//   high word: first prim record (SDF) or mesh id
BMO_D unsigned long long lean_part_info(const bmo_part& pt, const bmo_object& ob) {
    const unsigned lo = ((unsigned)pt.object & 0xffffu) | (ob.kind == BMO_OBJ_PLATE_BS ? 1u << 16 : 0u) |
                        ((pt.role == BMO_ROLE_COATING && ob.kind == BMO_OBJ_PLATE_BS) ? 1u << 17 : 0u) |
                        (pt.shape_kind == BMO_SHAPE_SDF ? 1u << 18 : 0u) | (((unsigned)pt.count & 15u) << 20);
    return ((unsigned long long)(unsigned)pt.first << 32) | lo;
}
BMO_D unsigned long long lds_u64(unsigned a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }

// tracing_step! (System.jl:57-110) for LEAN systems; see tracing_step (bmo_geom.cuh) for the rules it follows.
// sc / lb_addr: shared-memory addresses of this thread's scratch column and member-bounds column; pinfo: shared-memory
// address of the part words.
// [lo, hi): the parts trace_all looks at (retrace_system! restricts them, see tracing_step); hi < 0 means all parts.
BMO_D Hit tracing_step_lean(const TraceCtx& C, unsigned sc, unsigned lb_addr, unsigned pinfo, V3 pos, V3 dir, int hint_part, Stats& st,
                            const int lo = 0, const int hi = -1) {
    ls_store3(sc, LS_POS, pos);
    ls_store3(sc, LS_DIR, dir);
    ls_store3(sc, LS_INV, mk3(1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z));
    double res_t = INFINITY, ob_t = INFINITY;
    int res_part = -1, ob_part = -1, res_idx = 0, ob_idx = 0, cur_obj = -1;
    const int n_parts = hi < 0 ? C.n_parts : hi;
    for (int it = hint_part >= 0 ? lo - 1 : lo; it <= n_parts; it++) {
        const bool all = it >= lo;                   // false: the trace_one iteration on the hinted part
        const int part = all ? it : hint_part;
        const bool last = it == n_parts;
        const unsigned long long w = last ? 0ull : lds_u64(pinfo + 8u * (unsigned)part);
        const unsigned wl = (unsigned)w;
        const int obj = (all && !last) ? (int)(wl & 0xffffu) : -1;
        if (all && obj != cur_obj) {       // object boundary: trace_all's comparison (System.jl:62-67)
            if (ob_part >= 0 && (res_part < 0 || ob_t < res_t)) {
                res_t = ob_t; res_part = ob_part; res_idx = ob_idx;
                ls_store3(sc, LS_RES, ls_load3(sc, LS_OB));
            }
            ob_part = -1; ob_t = INFINITY;
            cur_obj = obj;
        }
        if (last) break;
        if (all && part == hint_part) continue;      // trace_one has just missed this shape: the same ray misses it again
        const bool is_sdf = (wl >> 18) & 1u;
        {
            double t_best = INFINITY;      // a plate beamsplitter prefers its coating on approximate equality: never culled against the best hit
            if (all && !((wl >> 16) & 1u)) {
                if (res_part >= 0) t_best = res_t;
                if (ob_part >= 0 && ob_t < t_best) t_best = ob_t;
            }
            if (!box_may_hit_ls(C.bounds + NBOUND * part + 4, sc, dir, t_best)) continue;
        }
        double t;
        int hidx;
        const int first = (int)(w >> 32);
        if (is_sdf) {
            const unsigned r = lean_march(C.prims, first, (int)((wl >> 20) & 15u), C.zr, C.bounds + NBOUND * part, sc, lb_addr);
            st.sdf += r >> 4;
            if (!(r & 15u)) continue;
            t = lds_f64(ls_slot(sc, LS_T)); hidx = first + (int)(r & 15u) - 1;
        } else {
            unsigned ntri = 0;
            double tm = 0.0;
            int fid = 0;
            const bool hit = mesh_hit_small(C.M, first, C.pose, ls_load3(sc, LS_POS), dir, tm, fid, ntri);
            st.tri += ntri;
            if (!hit) continue;
            t = tm; hidx = fid;
        }
        if (!all) {                        // trace_one: the hinted shape is accepted without looking at anything else
            res_t = t; res_part = part; res_idx = hidx;
            if (is_sdf) ls_store3(sc, LS_RES, ls_load3(sc, LS_CAND));
            break;
        }
        bool take = ob_part < 0 || t < ob_t;
        if (((wl >> 17) & 1u) && ob_part >= 0) take = jl_isapprox(t, ob_t) ? true : (t < ob_t);   // parts = (substrate, coating)
        if (take) {
            ob_t = t; ob_part = part; ob_idx = hidx;
            if (is_sdf) ls_store3(sc, LS_OB, ls_load3(sc, LS_CAND));
        }
    }
    Hit res; res.part = res_part; res.t = res_t; res.n = mk3(0, 0, 0);
    if (res_part >= 0) {   // normal3d of the one hit that is returned
        const unsigned long long w = lds_u64(pinfo + 8u * (unsigned)res_part);
        if (((unsigned)w >> 18) & 1u) {
            Stats tmp; tmp.sdf = 0; tmp.tri = 0;
            res.n = member_normal<false>(C.prims, res_idx, ls_load3(sc, LS_RES), C.zr, tmp);
            st.sdf += tmp.sdf;
        } else {
            res.n = mesh_face_normal(C.M, (int)(w >> 32), C.pose, res_idx);
        }
    }
    return res;
}

}  // namespace bmo
