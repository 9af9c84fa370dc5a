// bmo_interact.cuh -- device-side interact3d physics: reflection, Snell refraction / TIR, Fresnel
// coefficients and the 3-D polarisation ray-tracing update, Gaussian beamlet parameters.
// Citations are relative to /root/reference/src.
#pragma once
#include "bmo_geom.cuh"

namespace bmo {

// Utils/OpticUtils.jl:7-9
BMO_D V3 reflection3d(V3 dir, V3 normal) { return dir - (2 * dot(dir, normal)) * normal; }

// Utils/OpticUtils.jl:31-45.  err is set when the reference would throw ArgumentError.
BMO_D V3 refraction3d(V3 dir, V3 normal, double n1, double n2, bool& tir, bool& err) {
    if (!jl_isapprox(norm(dir), 1.0) || !jl_isapprox(norm(normal), 1.0)) err = true;
    double n = n1 / n2;
    double cosi = -dot(normal, dir);
    double sint2 = (n * n) * (1 - cosi * cosi);
    if (sint2 > 1.0) { tir = true; return reflection3d(dir, normal); }
    tir = false;
    double cost = sqrt(1 - sint2);
    double f = n * cosi - cost;
    return mk3(n * dir.x + f * normal.x, n * dir.y + f * normal.y, n * dir.z + f * normal.z);
}
// AbstractTypes/AbstractRay.jl:234-253  refraction3d(ray, n2): flips the normal when exiting
BMO_D V3 refraction3d_ray(V3 dir, V3 nrm, double n_ray, double n2, bool& tir, bool& err) {
    if (!(dot(dir, nrm) < 0)) nrm = nrm * -1.0;
    return refraction3d(dir, nrm, n_ray, n2, tir, err);
}
// Utils/LinearAlgebraUtils.jl:103-108
BMO_D double angle3d(V3 target, V3 reference) {
    double arg = jl_clamp(dot(target, reference) / (norm(target) * norm(reference)), -1.0, 1.0);
    return acos(arg);
}
// Utils/LinearAlgebraUtils.jl:6-8, atol = eps()
BMO_D bool isparallel3d(V3 a, V3 b) {
    double d = fabs(dot(normalize(a), normalize(b)));
    return fabs(d - 1.0) <= 2.220446049250313e-16;
}
// Utils/OpticUtils.jl:121-131
BMO_NI void fresnel_coefficients(double theta, double n, Cx& rs, Cx& rp, Cx& ts, Cx& tp) {
    double sn, cost;
    sincos(theta, &sn, &cost);
    Cx n2s2 = csqrt_(mkc(n * n - sn * sn, 0.0));
    rs = (cost - n2s2) / (cost + n2s2);
    rp = ((-(n * n)) * cost + n2s2) / ((n * n) * cost + n2s2);
    ts = rs + 1.0;
    tp = mkc(2 * n * cost, 0.0) / ((n * n) * cost + n2s2);
}
BMO_D bool is_internally_reflected(Cx rp, Cx rs) {  // :144-146
    return fabs(abs2(rs) - 1) <= 1e-6 && fabs(abs2(rp) - 1) <= 1e-6;
}
// PolarizedRays.jl:165-207 with J = diag(j11, j22, 1).  The reference's random vector for exactly
// normal incidence (LinearAlgebraUtils.jl:35-41) is pinned to Gram-Schmidt of (0.3, 0.5, 0.8): the
// result does not depend on it whenever |j11| == |j22| (mirrors, splitters, normal-incidence Fresnel).
BMO_NI void calculate_global_E0(V3 in_dir, V3 out_dir, V3 normal, Cx j11, Cx j22, const Cx* Ein, Cx* Eout) {
    V3 v = !isparallel3d(in_dir, out_dir) ? out_dir : normal;
    if (isparallel3d(in_dir, normal)) {
        V3 nw = mk3(0.3, 0.5, 0.8);
        double nn = norm(in_dir);
        nw = nw - ((dot(nw, in_dir) * in_dir) / (nn * nn));
        v = normalize(nw);
    }
    V3 s = normalize(cross(in_dir, v));
    V3 p1 = cross(in_dir, s);
    V3 oc[3];
    oc[0] = s;
    V3 negout = -out_dir;
    double ni = norm(in_dir), no = norm(negout);
    bool approx_neg = norm(in_dir - negout) <= 1.4901161193847656e-8 * (ni > no ? ni : no);
    if (isparallel3d(in_dir, out_dir) && !approx_neg) { oc[1] = p1; oc[2] = in_dir; }
    else { oc[1] = cross(out_dir, s); oc[2] = out_dir; }
    Cx J[3] = {j11, j22, mkc(1, 0)};
    V3 rows[3] = {s, p1, in_dir};
    Cx P[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Cx OJ0 = comp(oc[0], i) * J[0], OJ1 = comp(oc[1], i) * J[1], OJ2 = comp(oc[2], i) * J[2];
#pragma unroll
        for (int j = 0; j < 3; j++)
            P[i][j] = (OJ0 * comp(rows[0], j) + OJ1 * comp(rows[1], j)) + OJ2 * comp(rows[2], j);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) Eout[i] = (P[i][0] * Ein[0] + P[i][1] * Ein[1]) + P[i][2] * Ein[2];
}
// PolarizedRays.jl:54-56
BMO_D bool e0_orthogonal(V3 dir, const Cx* E0) {
    double re = dir.x * E0[0].re + dir.y * E0[1].re + dir.z * E0[2].re;
    double im = dir.x * E0[0].im + dir.y * E0[1].im + dir.z * E0[2].im;
    return sqrt(re * re + im * im) <= 1e-14;
}

// Result of a single-ray interaction (BeamInteraction of Beam.jl:74-77)
struct RayOut {
    V3 pos, dir;
    double n;
    int hint;    // part index or -1
    bool valid;  // false <=> interact3d returned nothing
    bool err;    // the reference would throw ArgumentError (non-unit dir / normal)
    bool warn;   // E0 not orthogonal to dir at atol 1e-14 (PolarizedRays.jl:54-56): flagged, tracing continues
    Cx E0[3];
};

// Lenses.jl:46-77 for a plain Ray, inlined: the out-of-line interact_refractive below takes its result by reference,
// which pins the caller's RayOut (and the E0 array it is handed) to local memory; plain rays and Gaussian triples
// (K2 MODE 0 / 2) keep everything in registers with this copy of its non-polarised branch (same operations, same order).
BMO_D void interact_refractive_plain(V3 rpos, V3 rdir, double rn, double t, V3 nrm, double n_opt, double n_sys, int self_part, RayOut& o) {
    V3 normal = nrm;
    double n1, n2;
    o.hint = -1;
    o.err = false; o.warn = false;
    o.valid = true;
    o.pos = rpos + t * rdir;
    if (dot(rdir, nrm) < 0) { n1 = rn; n2 = n_opt; o.hint = self_part; }
    else { n1 = n_opt; n2 = n_sys; normal = -normal; }
    bool tir;
    o.dir = refraction3d(rdir, normal, n1, n2, tir, o.err);
    if (tir) { o.hint = self_part; n2 = n_opt; }
    o.n = n2;
}
// Mirrors.jl:39-48 for a plain Ray, inlined for the same reason
BMO_D void interact_mirror_plain(V3 rpos, V3 rdir, double rn, double t, V3 nrm, RayOut& o) {
    o.valid = true; o.err = false; o.warn = false; o.hint = -1;
    o.pos = rpos + t * rdir;
    o.dir = reflection3d(rdir, nrm);
    o.n = rn;
}
// OpticalComponents/Lenses.jl:46-126  (n_opt = refractive_index(optic, lambda))
BMO_NI void interact_refractive(V3 rpos, V3 rdir, double rn, const Cx* rE0, bool polarized, double t, V3 nrm, double n_opt,
                               double n_sys, int self_part, RayOut& o) {
    V3 normal = nrm;
    double n1, n2;
    o.hint = -1;
    o.err = false; o.warn = false;
    o.valid = true;
    o.pos = rpos + t * rdir;
    if (dot(rdir, nrm) < 0) { n1 = rn; n2 = n_opt; o.hint = self_part; }
    else { n1 = n_opt; n2 = n_sys; normal = -normal; }
    if (!polarized) {
        bool tir;
        o.dir = refraction3d(rdir, normal, n1, n2, tir, o.err);
        if (tir) { o.hint = self_part; n2 = n_opt; }
        o.n = n2;
        return;
    }
    double thi = angle3d(rdir, -normal);
    Cx rs, rp, ts, tp;
    fresnel_coefficients(thi, n2 / n1, rs, rp, ts, tp);
    Cx j11, j22;
    if (is_internally_reflected(rp, rs)) {
        o.hint = self_part;
        n2 = n_opt;
        o.dir = reflection3d(rdir, normal);
        j11 = -rs; j22 = rp;
    } else {
        bool tir;
        o.dir = refraction3d(rdir, normal, n1, n2, tir, o.err);
        j11 = ts; j22 = tp;
    }
    calculate_global_E0(rdir, o.dir, nrm, j11, j22, rE0, o.E0);
    if (!e0_orthogonal(o.dir, o.E0)) o.warn = true;
    o.n = n2;
}
// OpticalComponents/Mirrors.jl:39-69
BMO_NI void interact_mirror(V3 rpos, V3 rdir, double rn, const Cx* rE0, bool polarized, double t, V3 nrm, RayOut& o) {
    o.valid = true; o.err = false; o.warn = false; o.hint = -1;
    o.pos = rpos + t * rdir;
    o.dir = reflection3d(rdir, nrm);
    o.n = rn;
    if (polarized) {
        calculate_global_E0(rdir, o.dir, nrm, mkc(-1, 0), mkc(1, 0), rE0, o.E0);
        if (!e0_orthogonal(o.dir, o.E0)) o.warn = true;
    }
}
// Polarizers/PolarizationFilter.jl:31-47 with JonesCalculus.jl:29-46: the Jones matrix is given in the global
// frame of the unrotated element; P = R J R' follows the element's orientation R and is projected into the
// plane transverse to the ray, P := Q P Q with Q = I - d d'; E0' = P E0.  The direction does not change.
// A ray whose |E0'| is approximately the cutoff is terminated (interact3d returns nothing).
BMO_NI void interact_polfilter(V3 rpos, V3 rdir, double rn, const Cx* rE0, double t, const double* Rm /* row-major 3x3 */,
                              const double* J /* row-major 3x3, then cutoff */, RayOut& o) {
    o.err = false; o.warn = false; o.hint = -1;
    o.pos = rpos + t * rdir;
    o.dir = rdir;
    o.n = rn;
    double RJ[3][3], P[3][3], Q[3][3], QP[3][3];
    const double dv[3] = {rdir.x, rdir.y, rdir.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) RJ[i][j] = (Rm[3 * i] * J[j] + Rm[3 * i + 1] * J[3 + j]) + Rm[3 * i + 2] * J[6 + j];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            P[i][j] = (RJ[i][0] * Rm[3 * j] + RJ[i][1] * Rm[3 * j + 1]) + RJ[i][2] * Rm[3 * j + 2];
            Q[i][j] = (i == j ? 1.0 : 0.0) - dv[i] * dv[j];
        }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) QP[i][j] = (Q[i][0] * P[0][j] + Q[i][1] * P[1][j]) + Q[i][2] * P[2][j];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double p0 = (QP[i][0] * Q[0][0] + QP[i][1] * Q[1][0]) + QP[i][2] * Q[2][0];
        double p1 = (QP[i][0] * Q[0][1] + QP[i][1] * Q[1][1]) + QP[i][2] * Q[2][1];
        double p2 = (QP[i][0] * Q[0][2] + QP[i][1] * Q[1][2]) + QP[i][2] * Q[2][2];
        o.E0[i] = (p0 * rE0[0] + p1 * rE0[1]) + p2 * rE0[2];
    }
    const double nrm = sqrt((abs2(o.E0[0]) + abs2(o.E0[1])) + abs2(o.E0[2]));
    o.valid = !jl_isapprox(nrm, J[9]);
    if (o.valid && !e0_orthogonal(o.dir, o.E0)) o.warn = true;
}
// ThinBeamsplitter.jl:73-106: children restart as Ray(pos, dir, lambda): n = 1, dir re-normalised
BMO_NI void bs_children(V3 rpos, V3 rdir, const Cx* rE0, bool polarized, double t, V3 nrm, double refl, double trans,
                       RayOut& tr, RayOut& rf) {
    V3 pos = rpos + t * rdir;
    tr.valid = rf.valid = true; tr.err = rf.err = false; tr.warn = rf.warn = false; tr.hint = rf.hint = -1;
    tr.pos = pos; rf.pos = pos;
    tr.n = 1.0; rf.n = 1.0;
    V3 rd = reflection3d(rdir, nrm);
    if (polarized) {
        calculate_global_E0(rdir, rdir, nrm, mkc(trans, 0), mkc(trans, 0), rE0, tr.E0);
        calculate_global_E0(rdir, rd, nrm, mkc(-refl, 0), mkc(refl, 0), rE0, rf.E0);
    }
    tr.dir = normalize(rdir);
    rf.dir = normalize(rd);
    if (polarized) {
        if (!e0_orthogonal(tr.dir, tr.E0)) tr.warn = true;
        if (!e0_orthogonal(rf.dir, rf.E0)) rf.warn = true;
    }
}

// Gaussian.jl:298-353 at chief point p0 of the segment whose rays are (c, w, d)
BMO_NI void gauss_parameters(V3 p0, V3 c_dir, double c_n, V3 w_pos, V3 w_dir, V3 d_pos, V3 d_dir, double lambda,
                            double& w, double& R, double& psi, double& w0) {
    double il = nan("");
    {
        double denom = dot(c_dir, d_dir);
        if (fabs(denom) > 1e-6) il = dot(p0 - d_pos, c_dir) / denom;
    }
    V3 y0 = d_pos + il * d_dir - p0;
    double y_d = norm(y0);
    y0 = y0 / y_d;
    double m_d = tan(kHalfPi - angle3d(y0, d_dir));
    il = nan("");
    {
        double denom = dot(c_dir, w_dir);
        if (fabs(denom) > 1e-6) il = dot(p0 - w_pos, c_dir) / denom;
    }
    y0 = w_pos + il * w_dir - p0;
    double y_w = norm(y0);
    y0 = y0 / y_w;
    double m_w = tan(kHalfPi - angle3d(y0, w_dir));
    double H = fabs(c_n * (y_w * m_d - y_d * m_w));
    if (!(fabs(H - lambda / kPi) <= 1e-6)) H = lambda / kPi;
    double E_kt = y_d * m_d + y_w * m_w;
    double F_kt = sqrt(m_d * m_d + m_w * m_w);
    w = sqrt(y_d * y_d + y_w * y_w);
    R = E_kt / (w * w);
    double z = E_kt / (F_kt * F_kt);
    psi = -atan2(1.0, sqrt(1 / (R * z) - 1));
    w0 = H / (c_n * F_kt);
    if (isnan(R)) R = 0.0;
    if (isnan(psi)) psi = 0.0;
    if (isnan(w0)) w0 = w;
    if (R < 0) psi = -psi;
}

}  // namespace bmo
