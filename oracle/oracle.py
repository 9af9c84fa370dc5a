"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

The oracle is a CPU restatement of BeamletOptics.jl's trace hot path (see oracle/orc_math.hpp for
the parity status).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Objects are created from the *reference's*
constructor-level arguments so the product's flattener is not shared with the checker.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")   # ORACLE_LIB: another build of the checker (the UBSan one)


def build(force=False):
    """Compile oracle/liboracle.so with the Makefile next to this file."""
    srcs = [os.path.join(_HERE, f) for f in ("bmo_oracle.cpp", "orc_math.hpp", "orc_shapes.hpp", "orc_optics.hpp", "orc_asphere.hpp")]
    if os.environ.get("ORACLE_LIB"):
        return _LIB_PATH
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_new.argtypes = [C.c_char_p, _dp, C.c_int, _ip, C.c_int]
        L.orc_refindex.argtypes = [C.c_int, _dp, C.c_int]
        L.orc_kin.argtypes = [C.c_int, C.c_char_p, _dp]
        L.orc_pose.argtypes = [C.c_int, _dp, _dp]
        L.orc_eval.argtypes = [C.c_char_p, _ip, C.c_int, _dp, C.c_int, _dp]
        L.orc_beam_export.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, C.c_int, _ip]
        L.orc_gauss_export.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, C.c_int, _dp, _ip]
        L.orc_pd_field.argtypes = [C.c_int, _dp]
        L.orc_pd_power.argtypes = [C.c_int]
        L.orc_pd_power.restype = C.c_double
        L.orc_spots.argtypes = [C.c_int, _dp, C.c_int]
        L.orc_psf_data.argtypes = [C.c_int, _dp, C.c_int]
        L.orc_psf_lims.argtypes = [C.c_int, C.c_double, C.c_int, _dp]
        L.orc_psf_intensity.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_bulk_trace_rays.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _ip, C.c_int, _dp]
        L.orc_bulk_trace_rays.restype = C.c_longlong
        L.orc_bulk_trace_beamlets.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.orc_bulk_trace_beamlets.restype = C.c_longlong
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _d(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    return a, a.ctypes.data_as(_dp), a.size


def _i(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())
    return a, a.ctypes.data_as(_ip), a.size


def _chk(rv):
    if rv < 0:
        raise OracleError(lib().orc_last_error().decode())
    return rv


def set_norm_zero_rule(rule):
    """1 = ZERO (default, pinned by test/runtests.jl:1309-1314), 0 = NAN -- see orc_math.hpp norm2_/norm3_."""
    lib().orc_set_norm_zero_rule(int(rule))


class Handle:
    """A shape / object / system / beam living inside the oracle library."""

    def __init__(self, h):
        self.h = int(h)

    # --- kinematic API (names of the reference with `!` -> `_`) ---
    def _kin(self, op, args=None):
        if args is None:
            _chk(lib().orc_kin(self.h, op.encode(), None))
        else:
            a, p, _ = _d(args)
            _chk(lib().orc_kin(self.h, op.encode(), p))
        return self

    def translate3d_(self, v): return self._kin("translate3d", v)
    def translate_to3d_(self, v): return self._kin("translate_to3d", v)
    def rotate3d_(self, axis, th): return self._kin("rotate3d", list(axis) + [th])
    def xrotate3d_(self, th): return self._kin("xrotate3d", [th, 0, 0])
    def yrotate3d_(self, th): return self._kin("yrotate3d", [th, 0, 0])
    def zrotate3d_(self, th): return self._kin("zrotate3d", [th, 0, 0])
    def align3d_(self, v): return self._kin("align3d", v)
    def reset_translation3d_(self): return self._kin("reset_translation3d")
    def reset_rotation3d_(self): return self._kin("reset_rotation3d")
    def set_new_origin3d_(self): return self._kin("set_new_origin3d")

    def pose(self):
        pos = np.zeros(3)
        d = np.zeros(9)
        _chk(lib().orc_pose(self.h, pos.ctypes.data_as(_dp), d.ctypes.data_as(_dp)))
        return pos, d.reshape(3, 3)

    def position(self): return self.pose()[0]
    def orientation(self): return self.pose()[1]
    def part(self, i): return Handle(_chk(lib().orc_part(self.h, int(i))))
    def shape(self): return Handle(_chk(lib().orc_shape_of(self.h)))

    def eval(self, fn, args=(), nout=16, extra_handles=()):
        ih, ihp, ni = _i([self.h] + [x.h for x in extra_handles])
        a, ap, na = _d(args if len(args) else [0.0])
        out = np.zeros(nout)
        n = _chk(lib().orc_eval(fn.encode(), ihp, ni, ap, na, out.ctypes.data_as(_dp)))
        return out[:n] if n <= nout else out

    # --- detectors ---
    def pd_field(self, n):
        out = np.zeros(2 * n * n)
        _chk(lib().orc_pd_field(self.h, out.ctypes.data_as(_dp)))
        z = out[0::2] + 1j * out[1::2]
        return z.reshape(n, n, order="F")  # [i, j] column-major like the reference's Matrix

    def pd_empty(self): _chk(lib().orc_pd_empty(self.h))
    def pd_power(self): return lib().orc_pd_power(self.h)

    # PSFDetector (PSFDetector.jl)
    def psf_data(self):
        n = _chk(lib().orc_psf_data(self.h, None, 0))
        out = np.zeros((max(n, 1), 9))
        _chk(lib().orc_psf_data(self.h, out.ctypes.data_as(_dp), n))
        return out[:n]

    def psf_empty(self): _chk(lib().orc_psf_empty(self.h))

    def psf_lims(self, crop_factor=1.0, center="centroid"):
        out = np.zeros(4)
        _chk(lib().orc_psf_lims(self.h, float(crop_factor), 0 if center == "centroid" else 1, out.ctypes.data_as(_dp)))
        return out

    def psf_intensity(self, n=100, crop_factor=1.0, center="centroid", x_min=np.inf, x_max=np.inf, z_min=np.inf, z_max=np.inf,
                      x0_shift=0.0, z0_shift=0.0):
        lims = np.array([x_min, x_max, z_min, z_max], dtype=np.float64)
        xs, zs, I = np.zeros(n), np.zeros(n), np.zeros(n * n)
        _chk(lib().orc_psf_intensity(self.h, int(n), float(crop_factor), 0 if center == "centroid" else 1, lims.ctypes.data_as(_dp),
                                     float(x0_shift), float(z0_shift), xs.ctypes.data_as(_dp), zs.ctypes.data_as(_dp), I.ctypes.data_as(_dp)))
        return xs, zs, I.reshape(n, n, order="F")

    def spots(self):
        n = _chk(lib().orc_spots(self.h, None, 0))
        out = np.zeros(2 * max(n, 1))
        _chk(lib().orc_spots(self.h, out.ctypes.data_as(_dp), n))
        return out[:2 * n].reshape(n, 2)

    def spots_empty(self): _chk(lib().orc_spots_empty(self.h))


def new(kind, d=(), ih=()):
    a, ap, na = _d(d if len(d) else [0.0])
    hh = [x.h if isinstance(x, Handle) else int(x) for x in ih]
    b, bp, nb = _i(hh if len(hh) else [0])
    return Handle(_chk(lib().orc_new(kind.encode(), ap, na, bp, len(hh))))


def refindex(n):
    """n: float (constant), (lambdas, ns) tuple (DiscreteRefractiveIndex) or ('sellmeier', B1,B2,B3,C1,C2,C3)."""
    if isinstance(n, (int, float)):
        a, ap, na = _d([float(n)])
        return Handle(_chk(lib().orc_refindex(0, ap, na)))
    if isinstance(n, tuple) and len(n) == 2:
        a, ap, na = _d(list(n[0]) + list(n[1]))
        return Handle(_chk(lib().orc_refindex(1, ap, na)))
    if isinstance(n, tuple) and n[0] == "sellmeier":
        a, ap, na = _d(n[1:])
        return Handle(_chk(lib().orc_refindex(2, ap, na)))
    raise TypeError(n)


def feval(fn, args, nout=16):
    a, ap, na = _d(args)
    ih, ihp, ni = _i([0])
    out = np.zeros(nout)
    n = _chk(lib().orc_eval(fn.encode(), ihp, 0, ap, na, out.ctypes.data_as(_dp)))
    return out[:n]


def mesh(vertices, faces, f32=False):
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int32)
    return new("Mesh", v.ravel(), [v.shape[0], f.shape[0], int(f32)] + list(f.ravel()))


def load_stl(path):
    """Mesh(load(path)) of a binary STL (Mesh.jl:48-70): Mesh{Float32} scaled by 1e-3."""
    lib().orc_load_stl.argtypes = [C.c_char_p]
    return Handle(_chk(lib().orc_load_stl(os.fsencode(path))))


def system(objects):
    return new("System", ih=objects)


def beam(pos, dir, lam=1000e-9):
    return new("Beam", list(pos) + list(dir) + [lam])


def polarized_beam(pos, dir, lam, E0):
    e = np.asarray(E0, dtype=np.complex128)
    return new("PolarizedBeam", list(pos) + list(dir) + [lam] + [x for c in e for x in (c.real, c.imag)])


def gaussian_beamlet(pos, dir, lam=1e-6, w0=1e-3, M2=1.0, P0=1e-3, z0=0.0, support=(1.0, 0.0, 0.0)):
    return new("GaussianBeamlet", list(pos) + list(dir) + [lam, w0, M2, P0, z0] + list(support))


def solve_system_(sys, beam_handle, r_max=100, retrace=False):
    """solve_system!(system, beam; r_max, retrace).  retrace=True re-validates a stored solution first
    (System.jl:188-428); on a fresh beam both give the same result."""
    if retrace:
        _chk(lib().orc_solve_retrace(sys.h, beam_handle.h, int(r_max)))
    else:
        _chk(lib().orc_solve(sys.h, beam_handle.h, int(r_max)))


RAY_FIELDS = 24


def _unpack(rec):
    return dict(pos=rec[:, 0:3], dir=rec[:, 3:6], n=rec[:, 6], lam=rec[:, 7], t=rec[:, 8], nrm=rec[:, 9:12],
                obj=rec[:, 12].astype(int), part=rec[:, 13].astype(int),
                E0=rec[:, 14:20:2] + 1j * rec[:, 15:20:2], polarized=rec[:, 20] != 0)


def beam_export(sys, b):
    """-> list of beams in BFS order: dict(parent=int, rays=dict of arrays)."""
    nr = C.c_int(0)
    nb = _chk(lib().orc_beam_export(sys.h, b.h, None, 0, None, 0, C.byref(nr)))
    rays = np.zeros((nr.value, RAY_FIELDS))
    beams = np.zeros((nb, 2), dtype=np.int32)
    _chk(lib().orc_beam_export(sys.h, b.h, rays.ctypes.data_as(_dp), nr.value, beams.ctypes.data_as(_ip), nb, C.byref(nr)))
    out, o = [], 0
    for p, n in beams:
        out.append(dict(parent=int(p), rays=_unpack(rays[o:o + n])))
        o += n
    return out


def gauss_export(sys, g):
    """-> list of beamlets in BFS order: dict(parent, lam, w0, E0, length, opl, chief, waist, div)."""
    nr = C.c_int(0)
    nb = _chk(lib().orc_gauss_export(sys.h, g.h, None, 0, None, 0, None, C.byref(nr)))
    rays = np.zeros((nr.value, RAY_FIELDS))
    beams = np.zeros((nb, 2), dtype=np.int32)
    gp = np.zeros((nb, 6))
    _chk(lib().orc_gauss_export(sys.h, g.h, rays.ctypes.data_as(_dp), nr.value, beams.ctypes.data_as(_ip), nb,
                                gp.ctypes.data_as(_dp), C.byref(nr)))
    out, o = [], 0
    for k, (p, n) in enumerate(beams):
        blk = rays[o:o + 3 * n].reshape(n, 3, RAY_FIELDS)
        out.append(dict(parent=int(p), lam=gp[k, 0], w0=gp[k, 1], E0=gp[k, 2] + 1j * gp[k, 3], length=gp[k, 4], opl=gp[k, 5],
                        chief=_unpack(blk[:, 0]), waist=_unpack(blk[:, 1]), div=_unpack(blk[:, 2])))
        o += 3 * n
    return out


def bulk_trace_rays(sys, pos, dir, lam, r_max=100, nthreads=1, max_seg=8, spot=None, want_segments=True):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    dir = np.ascontiguousarray(dir, dtype=np.float64)
    n = pos.shape[0]
    lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lam, dtype=np.float64), (n,)))
    seg = np.full((n, max_seg, 16), np.nan) if want_segments else None
    nseg = np.zeros(n, dtype=np.int32)
    sp = np.full((n, 2), np.nan)
    total = lib().orc_bulk_trace_rays(sys.h, n, pos.ctypes.data_as(_dp), dir.ctypes.data_as(_dp), lam.ctypes.data_as(_dp),
                                      int(r_max), int(nthreads), int(max_seg),
                                      seg.ctypes.data_as(_dp) if want_segments else None, nseg.ctypes.data_as(_ip),
                                      spot.h if spot is not None else -1, sp.ctypes.data_as(_dp))
    _chk(total)
    return dict(interactions=int(total), seg=seg, nseg=nseg, spot=sp)


def bulk_trace_beamlets(sys, g, r_max=100, nthreads=1):
    g = np.ascontiguousarray(g, dtype=np.float64)
    total = lib().orc_bulk_trace_beamlets(sys.h, g.shape[0], g.ctypes.data_as(_dp), int(r_max), int(nthreads))
    _chk(total)
    return int(total)
