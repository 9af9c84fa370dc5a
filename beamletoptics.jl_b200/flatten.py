"""Flattener: `Leaves(system.objects)` -> structure-of-arrays tables of include/bmo.h.

Order is semantic (SURVEY hard part 16): objects keep the pre-order leaf order of
`system.objects` (tie-break of trace_all, System.jl:57-72) and parts keep the order of
`shape(object)` (AbstractRay.jl:130-155).  Refractive indices are evaluated here, on the host, once
per (part, distinct wavelength) (Lenses.jl:37-38); an untabulated wavelength raises KeyError like
DiscreteRefractiveIndex does (RefractiveIndexUtils.jl:31).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import components as co
from . import linalg as la
from . import shapes as sh

SHAPE_SDF, SHAPE_MESH = 0, 1
ROLE_SINGLE, ROLE_FRONT, ROLE_BACK, ROLE_SUBSTRATE, ROLE_COATING = range(5)
OBJ_REFRACTIVE, OBJ_MIRROR, OBJ_THIN_BS, OBJ_PLATE_BS, OBJ_CUBE_BS, OBJ_DOUBLET, OBJ_PD, OBJ_SPOT, OBJ_STOP, OBJ_PSF, OBJ_POLFILTER = range(11)

_KIND = {"refractive": OBJ_REFRACTIVE, "mirror": OBJ_MIRROR, "thin_bs": OBJ_THIN_BS, "plate_bs": OBJ_PLATE_BS,
         "cube_bs": OBJ_CUBE_BS, "doublet": OBJ_DOUBLET, "pd": OBJ_PD, "spot": OBJ_SPOT, "stop": OBJ_STOP, "psf": OBJ_PSF, "polfilter": OBJ_POLFILTER}
_ROLES = {"plate_bs": (ROLE_SUBSTRATE, ROLE_COATING), "cube_bs": (ROLE_FRONT, ROLE_BACK, ROLE_COATING),
          "doublet": (ROLE_FRONT, ROLE_BACK)}

BOUND_REL, BOUND_ABS = 1e-9, 1e-6   # inflation of the bounding spheres / boxes (see bmo_geom.cuh sdf_intersect_f, tracing_step)


def _prim_record(s, ext):
    p = L.bmo_prim()
    p.type = s.type if isinstance(s, sh.PrimSDF) else sh.MENISCUS
    p.pos[:] = s.pos
    p.tdir[:] = [s.tdir[i][j] for i in range(3) for j in range(3)]
    p.par[:] = s.par if isinstance(s, sh.PrimSDF) else (0.0, 0.0, 0.0, 0.0)
    if isinstance(s, sh.AsphericalSurfaceSDF):      # parameter block in tables.ext
        p.ext_first, p.ext_count = len(ext), len(s.ext)
        ext.extend(s.ext)
    return p


def _emit_sdf(shape, prims, ext, index_of=None):
    """Append the prim records of one top-level SDF shape; returns (first, count).  index_of: id(member) -> prim record."""
    first = len(prims)
    members = shape.sdfs if isinstance(shape, sh.UnionSDF) else [shape]
    for m in members:
        if index_of is not None:
            index_of[id(m)] = len(prims)
        if isinstance(m, sh.MeniscusLensSDF):
            prims.append(_prim_record(m, ext))
            for child in (m.convex, m.cylinder, m.concave):
                prims.append(_prim_record(child, ext))
        elif isinstance(m, sh.PrimSDF):
            prims.append(_prim_record(m, ext))
        else:
            raise TypeError(f"unsupported SDF member {type(m).__name__}")
    return first, len(prims) - first


class FlatSystem:
    """Host tables + the bookkeeping needed to map device ids back to host objects."""

    def __init__(self, system, lambdas, norm_zero_rule=1):
        self.lambdas = [float(x) for x in lambdas]
        self.system = system
        self.prim_index, self.mesh_index = {}, {}     # id(host shape) -> prim record / mesh index (kinematic tree)
        leaves = [o for o in system.leaves() if not isinstance(o, co.NonInteractableObject)]
        if not leaves:
            raise ValueError("system has no traceable objects")
        self.objects = leaves           # device object index -> host object
        self.part_owner = []            # device part index -> host sub-object (Lens of a doublet, coating, ...)
        prims, parts, objs, meshes = [], [], [], []
        verts, faces, ntab, jones, ext = [], [], [], [], []
        nv = nf = 0
        for oi, o in enumerate(leaves):
            kind = _KIND[o.kind]
            sub = o.parts if o.multi else [o]
            roles = _ROLES.get(o.kind, (ROLE_SINGLE,))
            rec = L.bmo_object()
            rec.kind, rec.first_part, rec.n_parts, rec.pd_n = kind, len(parts), len(sub), 0
            det_shape = o.shape if not o.multi else None
            if det_shape is not None:
                rec.pos[:] = det_shape.pos
                rec.dir[:] = [det_shape.dir[i][j] for i in range(3) for j in range(3)]
            if kind == OBJ_PD:
                rec.pd_n, rec.pd_lo, rec.pd_hi = o.n, o.lo, o.hi
            if kind == OBJ_POLFILTER:     # GlobalJonesBasis (3x3, row-major) + cutoff -> tables.jones[pd_n]
                rec.pd_n = len(jones)
                jones.append([float(x) for row in o.JMat for x in row] + [float(o.cutoff)])
            objs.append(rec)
            for s_obj, role in zip(sub, roles):
                shape = s_obj.shape
                pr = L.bmo_part()
                pr.object, pr.role, pr.n_row = oi, role, -1
                pr.reflectance = getattr(s_obj, "reflectance", 0.0)
                pr.transmittance = getattr(s_obj, "transmittance", 0.0)
                if isinstance(shape, sh.AbstractSDF):
                    pr.shape_kind = SHAPE_SDF
                    pr.first, pr.count = _emit_sdf(shape, prims, ext, self.prim_index)
                elif isinstance(shape, sh.Mesh):
                    pr.shape_kind = SHAPE_MESH
                    pr.first, pr.count = len(meshes), 1
                    self.mesh_index[id(shape)] = len(meshes)
                    m = L.bmo_mesh()
                    m.first_vertex, m.n_vertices = nv, shape.vertices.shape[0]
                    m.first_face, m.n_faces = nf, shape.faces.shape[0]
                    m.f32 = int(shape.f32)
                    meshes.append(m)
                    verts.append(np.ascontiguousarray(shape.vertices, dtype=np.float64))
                    faces.append(np.ascontiguousarray(shape.faces, dtype=np.int32))
                    nv += shape.vertices.shape[0]
                    nf += shape.faces.shape[0]
                else:
                    raise TypeError(f"unsupported shape {type(shape).__name__}")
                c, r = shape.local_bound()
                lo, hi = sh._box_of(shape.box_points())        # world AABB of everything with sdf <= 0 / of the mesh
                pad = [BOUND_ABS + BOUND_REL * max(abs(lo[k]), abs(hi[k])) for k in range(3)]
                pr.bound[:] = (c[0], c[1], c[2], r * (1 + BOUND_REL) + BOUND_ABS,
                               lo[0] - pad[0], lo[1] - pad[1], lo[2] - pad[2], hi[0] + pad[0], hi[1] + pad[1], hi[2] + pad[2])
                if hasattr(s_obj, "refractive_index"):
                    pr.n_row = len(ntab)
                    ntab.append([float(s_obj.refractive_index(lam)) for lam in self.lambdas])
                parts.append(pr)
                self.part_owner.append(s_obj)
        self._prims = (L.bmo_prim * max(len(prims), 1))(*prims)
        self._parts = (L.bmo_part * len(parts))(*parts)
        self._objs = (L.bmo_object * len(objs))(*objs)
        self._meshes = (L.bmo_mesh * max(len(meshes), 1))(*meshes)
        self._verts = np.concatenate(verts) if verts else np.zeros((0, 3))
        self._faces = np.concatenate(faces) if faces else np.zeros((0, 3), dtype=np.int32)
        self._lams = np.array(self.lambdas, dtype=np.float64)
        self._ntab = np.array(ntab, dtype=np.float64).reshape(len(ntab), len(self.lambdas)) if ntab else np.zeros((0, len(self.lambdas)))
        self.n_prims, self.n_parts, self.n_objects, self.n_meshes = len(prims), len(parts), len(objs), len(meshes)
        t = L.bmo_tables()
        t.n_prims, t.prims = len(prims), self._prims
        t.n_parts, t.parts = len(parts), self._parts
        t.n_objects, t.objects = len(objs), self._objs
        t.n_meshes, t.meshes = len(meshes), self._meshes
        t.n_vertices, t.vertices = nv, self._verts.ctypes.data_as(C.POINTER(C.c_double))
        t.n_faces, t.faces = nf, self._faces.ctypes.data_as(C.POINTER(C.c_int32))
        t.n_lambda, t.lambdas = len(self.lambdas), self._lams.ctypes.data_as(C.POINTER(C.c_double))
        t.n_rows, t.n_table = len(ntab), self._ntab.ctypes.data_as(C.POINTER(C.c_double))
        self._jones = np.array(jones, dtype=np.float64).reshape(len(jones), 10) if jones else np.zeros((0, 10))
        t.n_jones, t.jones = len(jones), self._jones.ctypes.data_as(C.POINTER(C.c_double))
        self._ext = np.array(ext, dtype=np.float64) if ext else np.zeros(1)
        t.n_ext, t.ext = len(ext), self._ext.ctypes.data_as(C.POINTER(C.c_double))
        t.n_system = float(system.n)
        t.norm_zero_rule = int(norm_zero_rule)
        self.tables = t

    # pose tables of this flattening, used for batched pose sweeps (bmo_system_set_poses)
    def pose_arrays(self):
        prims = np.frombuffer(self._prims, dtype=np.uint8).copy() if self.n_prims else np.zeros(0, dtype=np.uint8)
        bounds = np.array([list(p.bound) for p in self._parts], dtype=np.float64)
        det_pos = np.array([list(o.pos) for o in self._objs], dtype=np.float64)
        det_dir = np.array([list(o.dir) for o in self._objs], dtype=np.float64)
        return prims[:self.n_prims * C.sizeof(L.bmo_prim)], self._verts.copy(), bounds, det_pos, det_dir

    # ---- K5: kinematic tree of the system for batched pose updates on the device (bmo_system_set_kinematics) ----
    def kinematics(self):
        """(nodes, prim_bounds, node_of): bmo_kin_node array in pre-order, [n_prims][10] local bounds of the top-level prim
        records, and id(host object / group / shape) -> node index.  Mirrors who owns a pose in the reference:
        ObjectGroup (center, dir), MultiShape objects (position = position of a pivot part), UnionSDF (own pose + members
        with world poses), primitives, meshes."""
        nodes, node_of = [], {}
        obj_index = {id(o): i for i, o in enumerate(self.objects)}

        def new(kind, host, pos=(0.0, 0.0, 0.0), dirm=la.IDENTITY, index=-1, flags=0, obj=-1):
            nd = L.bmo_kin_node()
            nd.kind, nd.size, nd.pos_ref, nd.index, nd.flags, nd.object = kind, 1, len(nodes), index, flags, obj
            nd.pos[:] = [float(x) for x in pos]
            nd.dir[:] = [float(dirm[i][j]) for i in range(3) for j in range(3)]
            node_of[id(host)] = len(nodes)
            nodes.append(nd)
            return len(nodes) - 1

        def emit_shape(shape, obj=-1):
            if isinstance(shape, sh.UnionSDF):
                i = new(2, shape, shape.pos, shape.dir, obj=obj)
                for m in shape.sdfs:
                    emit_shape(m)
                nodes[i].size = len(nodes) - i
                return i
            if isinstance(shape, sh.Mesh):
                return new(4, shape, shape.pos, shape.dir, index=self.mesh_index[id(shape)], obj=obj)
            if isinstance(shape, (sh.PrimSDF, sh.MeniscusLensSDF)):
                sphere = isinstance(shape, sh.PrimSDF) and shape.type == sh.SPHERE
                return new(3, shape, shape.pos, shape.dir, index=self.prim_index[id(shape)], flags=1 if sphere else 0, obj=obj)
            raise TypeError(f"unsupported shape {type(shape).__name__}")

        def emit(o):
            if isinstance(o, co.NonInteractableObject):
                return None
            if isinstance(o, co.ObjectGroup):
                i = new(0, o, o.center, o.dir)
                for c in o.parts:
                    emit(c)
                nodes[i].size = len(nodes) - i
                return i
            if o.multi:
                i = new(1, o)
                kids = [emit(c) for c in o.parts]
                nodes[i].size = len(nodes) - i
                # position(object): the part whose position the object reports (first part; the coating of a plate splitter)
                pivot = 1 if o.kind == "plate_bs" else 0
                nodes[i].pos_ref = nodes[kids[pivot]].pos_ref
                return i
            j = emit_shape(o.shape, obj_index.get(id(o), -1))
            node_of[id(o)] = j
            return j

        for o in self.system.objects:
            emit(o)
        bounds = np.zeros((max(self.n_prims, 1), 10))
        for part_i in range(self.n_parts):
            shape = self.part_owner[part_i].shape
            if not isinstance(shape, sh.AbstractSDF):
                continue
            for m in (shape.sdfs if isinstance(shape, sh.UnionSDF) else [shape]):
                # bounds in the member's own frame: evaluate them with the member moved to the origin frame
                pos, dirm, tdir = m.pos, m.dir, m.tdir
                m.pos, m.dir, m.tdir = (0.0, 0.0, 0.0), la.IDENTITY, la.IDENTITY
                try:
                    c, r = m.local_bound()
                    lo, hi = sh._box_of(m.box_points())
                finally:
                    m.pos, m.dir, m.tdir = pos, dirm, tdir
                bounds[self.prim_index[id(m)]] = [c[0], c[1], c[2], r, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]]
        arr = (L.bmo_kin_node * len(nodes))(*nodes)
        return arr, np.ascontiguousarray(bounds), node_of

    def object_index(self, obj):
        for i, o in enumerate(self.objects):
            if o is obj:
                return i
        raise KeyError("object is not part of the flattened system")
