"""K5: batched kinematics on the device (bmo_system_set_kinematics / bmo_system_apply_poses) against the host mirror of
the reference's kinematic API (translate3d!/rotate3d!: AbstractShape.jl:56-94, AbstractShapeTrait.jl:88-128,
UnionSDF.jl:63-82, Mesh.jl:78-96, ObjectGroups.jl:21-47, rotate3d LinearAlgebraUtils.jl:55-65).

CPU part: the micro-op expansion of KinProgram, replayed by a small interpreter written here, reproduces the host
objects' poses bit for bit.  GPU part: the device kernel produces the same prim records, vertices and detector poses as
re-flattening the moved host system pose by pose (bit for bit), conservative part bounds, and the C5 sweep through the
device poses equals the sweep through host-flattened poses and the oracle."""
import copy
import ctypes as C
import math

import numpy as np
import pytest

from tests import scenes
from tests import scenes2 as s2
from tests.scenes import INCH


def _scene(bmo):
    """Every kind of kinematic node: ObjectGroup (nested), MultiShape objects (doublet, cube and plate splitters), UnionSDF
    lenses, a meniscus, a sphere (orientation pinned), SDF mirror, Float64 meshes (detectors, retroreflector), a Float32 mesh."""
    dl = bmo.SphericalDoubletLens(*scenes.AC254, 1.6456, 1.7168)
    men = bmo.SphericalLens(0.03, 0.05, 2e-3, INCH, 1.6); men.translate3d_([0.0, 0.04, 0.0])
    ball = bmo.Prism(bmo.SphereSDF(4e-3), 1.5); ball.translate3d_([0.02, 0.0, 0.01])
    cbs = bmo.CubeBeamsplitter(INCH, 1.5); cbs.translate3d_([0.0, 0.1, 0.0]); cbs.zrotate3d_(0.3)
    pbs = bmo.RectangularPlateBeamsplitter(36e-3, 25e-3, 3e-3, 1.5); pbs.translate3d_([0.05, 0.1, 0.0]); pbs.zrotate3d_(math.radians(45))
    m1 = bmo.RoundPlanoMirror(INCH, 5e-3); m1.translate3d_([0.1, 0.1, 0.0])
    retro = bmo.Retroreflector(0.02); retro.translate3d_([-0.1, 0.1, 0.0])
    v, f = s2.uv_sphere_f32(10.0, 8, 6)
    f32 = bmo.Mirror(bmo.Mesh(v, f, scale=float(np.float32(1e-3)), f32=True)); f32.translate3d_([0.0, 0.2, 0.05])
    pd = bmo.Photodetector(5e-3, 16); pd.translate3d_([0.0, 0.3, 0.0])
    sd = bmo.Spotdetector(5e-3); sd.translate3d_([0.1, 0.3, 0.0])
    inner = bmo.ObjectGroup([cbs, pbs])
    arm = bmo.ObjectGroup([inner, m1, retro])
    optics = bmo.ObjectGroup([dl, men, ball])
    system = bmo.System([optics, arm, f32, pd, sd])
    return dict(system=system, dl=dl, men=men, ball=ball, cbs=cbs, pbs=pbs, m1=m1, retro=retro, f32=f32, pd=pd, sd=sd, inner=inner, arm=arm, optics=optics)


def _poses(P):
    rng = np.random.default_rng(5)
    return dict(off_arm=rng.normal(size=(P, 3)) * 1e-2, th_arm=rng.normal(size=P) * 0.4, th_dl=rng.normal(size=P) * 0.2,
                off_m1=rng.normal(size=(P, 3)) * 1e-6, tgt_f32=rng.normal(size=(P, 3)) * 0.1, th_f32=rng.normal(size=P),
                th_cbs=rng.normal(size=P), th_men=rng.normal(size=P) * 0.1, th_opt=rng.normal(size=P) * 0.3, th_pd=rng.normal(size=P) * 0.2)


AX1 = (0.0, 0.0, 1.0)
AX2 = tuple(float(x) for x in np.array([1.0, 2.0, -0.5]) / np.linalg.norm([1.0, 2.0, -0.5]))


def _record(prog, sc, q):
    """The kinematic calls of the sweep, recorded for all poses."""
    prog.translate3d_(sc["arm"], q["off_arm"])
    prog.rotate3d_(sc["arm"], AX1, q["th_arm"])
    prog.rotate3d_(sc["dl"], AX2, q["th_dl"])
    prog.translate3d_(sc["m1"], q["off_m1"])
    prog.translate_to3d_(sc["f32"], q["tgt_f32"])
    prog.rotate3d_(sc["f32"], AX2, q["th_f32"])
    prog.rotate3d_(sc["cbs"], AX1, q["th_cbs"])
    prog.rotate3d_(sc["men"], AX2, q["th_men"])
    prog.rotate3d_(sc["optics"], AX1, q["th_opt"])
    prog.rotate3d_(sc["pd"], AX1, q["th_pd"])
    prog.translate_to3d_(sc["pbs"], q["tgt_f32"])


def _apply_host(sc, q, p):
    """The same calls on (a copy of) the host objects for pose p."""
    sc["arm"].translate3d_(q["off_arm"][p])
    sc["arm"].rotate3d_(AX1, float(q["th_arm"][p]))
    sc["dl"].rotate3d_(AX2, float(q["th_dl"][p]))
    sc["m1"].translate3d_(q["off_m1"][p])
    sc["f32"].translate_to3d_(q["tgt_f32"][p])
    sc["f32"].rotate3d_(AX2, float(q["th_f32"][p]))
    sc["cbs"].rotate3d_(AX1, float(q["th_cbs"][p]))
    sc["men"].rotate3d_(AX2, float(q["th_men"][p]))
    sc["optics"].rotate3d_(AX1, float(q["th_opt"][p]))
    sc["pd"].rotate3d_(AX1, float(q["th_pd"][p]))
    sc["pbs"].translate_to3d_(q["tgt_f32"][p])


def _prim_table(flat):
    n = flat.n_prims
    return np.array([[pr.pos[k] for k in range(3)] + [pr.tdir[k] for k in range(9)] for pr in flat._prims[:n]])


def _interpret(prog, flat, pose):
    """Reference interpreter of the micro-op program (plain Python floats, the operation order bmo_pose.cu documents)."""
    nodes = prog.nodes
    pos = [[float(x) for x in nd.pos] for nd in nodes]
    dirs = [[float(x) for x in nd.dir] for nd in nodes]
    verts = flat._verts.copy()
    mesh_rng = {}
    for i in range(flat.n_meshes):
        m = flat._meshes[i]
        mesh_rng[i] = (int(m.first_vertex), int(m.n_vertices), bool(m.f32))
    r32 = lambda x: float(np.float32(x))
    params = np.stack(prog.params, axis=1)[pose] if prog.params else np.zeros((0, 9))

    def matmul(R, D):
        return [R[3 * i] * D[j] + R[3 * i + 1] * D[3 + j] + R[3 * i + 2] * D[6 + j] for i in range(3) for j in range(3)]

    def matvec(R, v):
        return [R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2] for i in range(3)]

    def translate(a, off):
        for j in range(a, a + nodes[a].size):
            f32 = nodes[j].kind == 4 and mesh_rng[nodes[j].index][2]
            pos[j] = [pos[j][k] + off[k] for k in range(3)]
            if f32:
                pos[j] = [r32(x) for x in pos[j]]
            if nodes[j].kind == 4:
                fv, nv, _ = mesh_rng[nodes[j].index]
                w = verts[fv:fv + nv] + np.array(off)
                verts[fv:fv + nv] = w.astype(np.float32).astype(np.float64) if f32 else w

    for kind, a, b, k in prog.ops:
        q = [float(x) for x in params[k]]
        nd = nodes[a]
        if kind == prog.TRANSLATE:
            translate(a, q[:3])
        elif kind == prog.TRANSLATE_TO:
            pa = pos[nd.pos_ref]
            translate(a, [q[i] - pa[i] for i in range(3)])
        elif kind == prog.ROT_FRAME:
            dirs[a] = matmul(q, dirs[a])
        elif kind == prog.ROT_LEAF:
            if nd.kind == 3:
                if not (nd.flags & 1):
                    dirs[a] = matmul(q, dirs[a])
            else:
                fv, nv, f32 = mesh_rng[nd.index]
                d = verts[fv:fv + nv] - np.array(pos[a])
                if f32:
                    d = d.astype(np.float32).astype(np.float64)
                r = np.empty_like(d)
                for i in range(3):
                    r[:, i] = q[3 * i] * d[:, 0] + q[3 * i + 1] * d[:, 1] + q[3 * i + 2] * d[:, 2]
                w = r + np.array(pos[a])
                verts[fv:fv + nv] = w.astype(np.float32).astype(np.float64) if f32 else w
                dirs[a] = matmul(q, dirs[a])
                if f32:
                    dirs[a] = [r32(x) for x in dirs[a]]
        elif kind == prog.PIVOT:
            pa, pb = pos[nd.pos_ref], pos[nodes[b].pos_ref]
            v = [pa[i] - pb[i] for i in range(3)]
            rv = matvec(q, v)
            translate(a, [rv[i] - v[i] for i in range(3)])
    return pos, dirs, verts


def test_kin_program_expansion_reproduces_host_kinematics(bmo):
    P = 5
    sc = _scene(bmo)
    from bmo_b200.flatten import FlatSystem
    flat = FlatSystem(sc["system"], [1e-6])
    q = _poses(P)
    prog = bmo.KinProgram(flat, P)
    _record(prog, sc, q)
    kinds = [nd.kind for nd in prog.nodes]
    assert set(kinds) == {0, 1, 2, 3, 4}
    for i, nd in enumerate(prog.nodes):      # pre-order layout: children tile the subtree
        c = i + 1
        while c < i + nd.size:
            c += prog.nodes[c].size
        assert c == i + nd.size
    for p in range(P):
        moved = copy.deepcopy(sc)
        _apply_host(moved, q, p)
        ref = FlatSystem(moved["system"], [1e-6])
        pos, dirs, verts = _interpret(prog, flat, p)
        got = _prim_table(flat)
        for j, nd in enumerate(prog.nodes):
            if nd.kind == 3:
                got[nd.index, 0:3] = pos[j]
                got[nd.index, 3:12] = [dirs[j][3 * b + a] for a in range(3) for b in range(3)]    # tdir = transpose(dir)
        want = _prim_table(ref)
        top = [nd.index for nd in prog.nodes if nd.kind == 3]
        assert np.array_equal(got[top], want[top]), p
        assert np.array_equal(verts, ref._verts), p
        for j, nd in enumerate(prog.nodes):
            if nd.object >= 0:
                o = ref._objs[nd.object]
                assert [o.pos[k] for k in range(3)] == pos[j] and [o.dir[k] for k in range(9)] == dirs[j], (p, j)
        # group state follows as well
        assert list(moved["arm"].position()) == pos[prog.node_of[id(sc["arm"])]]
        assert [x for row in moved["optics"].orientation() for x in row] == dirs[prog.node_of[id(sc["optics"])]]


@pytest.mark.gpu
def test_device_poses_match_host_flattening_bitwise(bmo):
    from bmo_b200.flatten import FlatSystem
    P = 7
    sc = _scene(bmo)
    q = _poses(P)
    dsys = bmo.upload_system(sc["system"], [1e-6])
    prog = bmo.KinProgram(dsys.flat, P)
    _record(prog, sc, q)
    prog.apply(dsys)
    for p in range(P):
        moved = copy.deepcopy(sc)
        _apply_host(moved, q, p)
        ref = FlatSystem(moved["system"], [1e-6])
        prims, verts, bounds, det = bmo.get_pose_tables(dsys, p)
        got = np.array([[pr.pos[k] for k in range(3)] + [pr.tdir[k] for k in range(9)] + [pr.par[k] for k in range(1, 4)] + [pr.type] for pr in prims[:ref.n_prims]])
        want = np.array([[pr.pos[k] for k in range(3)] + [pr.tdir[k] for k in range(9)] + [pr.par[k] for k in range(1, 4)] + [pr.type] for pr in ref._prims[:ref.n_prims]])
        assert np.array_equal(got, want), p
        assert np.array_equal(verts, ref._verts), p
        for oi in range(ref.n_objects):
            o = ref._objs[oi]
            assert np.array_equal(det[oi], np.array([o.pos[k] for k in range(3)] + [o.dir[k] for k in range(9)])), (p, oi)
        # bounds: conservative and close to the flattener's
        hb = np.array([list(pt.bound) for pt in ref._parts])
        assert np.abs(bounds[:, 0:3] - hb[:, 0:3]).max() <= 1e-9
        assert (bounds[:, 3] >= hb[:, 3] - 1e-12).all() and np.abs(bounds[:, 3] - hb[:, 3]).max() <= 1e-6
        assert (bounds[:, 4:7] <= hb[:, 4:7] + 1e-12).all() and (bounds[:, 7:10] >= hb[:, 7:10] - 1e-12).all()
        assert np.abs(bounds[:, 4:] - hb[:, 4:]).max() <= 2e-3     # meniscus boxes are taken in the frame of the meniscus (looser)


@pytest.mark.gpu
def test_traces_through_device_poses_equal_traces_through_host_poses(bmo):
    """Rays traced through pose p of the device-made tables == rays traced through the re-flattened host system."""
    P = 4
    sc = _scene(bmo)
    # the two splitters between the retroreflector and m1 would form a cavity: every round trip multiplies the beams by
    # 16 and neither the reference nor the device trace would ever end.  Out of the plane of the bundle they stay part of
    # the kinematic tree (their tables are compared bit for bit above) without closing the loop.
    sc["m1"].translate3d_([0.0, 0.0, 1.0]); sc["retro"].translate3d_([0.0, 0.0, -1.0])
    q = _poses(P)
    for key in ("off_arm", "tgt_f32"):
        q[key] *= 0.05                      # keep the optics roughly in the path of the bundle
    for key in ("th_arm", "th_f32", "th_cbs", "th_opt", "th_dl", "th_men", "th_pd"):
        q[key] *= 0.1
    dsys = bmo.upload_system(sc["system"], [707e-9])
    prog = bmo.KinProgram(dsys.flat, P)
    _record(prog, sc, q)
    prog.apply(dsys)
    pos, d = scenes.fibonacci_disc(256)
    lam_id = np.zeros(256, np.int32)
    for p in range(P):
        res = bmo.trace_rays(dsys, pos, d, lam_id, pose_id=np.full(256, p, np.int32), r_max=20)
        moved = copy.deepcopy(sc)
        _apply_host(moved, q, p)
        ref = bmo.solve_system_(moved["system"], bmo.RayBundle(pos, d, 707e-9), r_max=20)
        a, b = res.segments(), ref.segments()
        assert np.array_equal(res.beams()["nseg"], ref.beams()["nseg"])
        for key in ("pos", "dir", "t", "nrm", "n", "obj"):
            assert np.array_equal(a[key], b[key]), (p, key)
        assert res.n_segments > 2 * 256


@pytest.mark.gpu
def test_c5_sweep_with_device_poses(bmo, orc):
    P, n = 12, 48
    B = s2.MZI_BEAM
    sc = s2.mzi(bmo, pd_n=n)
    g = bmo.GaussianBeamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], P0=B["P0"], support=B["support"])
    base = sc["m1"].position()
    shifts = np.array([s2.mzi_shift(p, P) for p in range(P)])
    dev = bmo.solve_pose_sweep_device(sc["system"], g, P, lambda prog: prog.translate3d_(sc["m1"], shifts), sc["pd"])
    assert sc["m1"].position() == base                      # the host objects are not moved

    def apply_pose(p):
        sc["m1"].translate_to3d_(base)
        sc["m1"].translate3d_(s2.mzi_shift(p, P))
    host = bmo.solve_pose_sweep(sc["system"], g, P, apply_pose, sc["pd"])
    sc["m1"].translate_to3d_(base)
    assert np.array_equal(dev["fields"], host["fields"]) and np.array_equal(dev["power"], host["power"])
    for p in (0, 5, 11):
        o = s2.mzi_oracle(pd_n=n)
        o["m1"].translate3d_(s2.mzi_shift(p, P))
        og = orc.gaussian_beamlet(B["pos"], B["dir"], B["lam"], B["w0"], M2=B["M2"], P0=B["P0"], support=B["support"])
        orc.solve_system_(o["system"], og)
        ref = o["pd"].pd_field(n)
        assert np.linalg.norm((dev["fields"][p] - ref).ravel()) / np.linalg.norm(ref.ravel()) <= 1e-8
