"""Multi-GPU layer of the trace path: one process per GPU, `torch.distributed` for the
plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Rays, beamlets and poses are independent, the system tables are small and replicated, so the path
shards without a data-path collective: every rank traces its slice of the bundle on its own GPU and
owns the corresponding slice of the outputs.  The single exchange step is the Photodetector: each
rank accumulates *its* beamlets onto a full n x n complex128 partial field and the partial fields
are summed with one all-reduce of 2 n^2 doubles (the reference adds the beamlet fields serially into
`pd.field`, Photodetector.jl:103; addition order across ranks changes the sum at the 1e-16 level).
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import beams as bm


class FieldComm:
    """bmo_comm (include/bmo.h): the NCCL communicator behind the C ABI, one per process and GPU.  The 128-byte id
    made by rank 0 (`FieldComm.unique_id()`) has to reach every rank by some host transport; `from_torch` uses the
    process group torchrun set up, a Julia host would use Distributed / MPI / a file."""

    def __init__(self, rank, world_size, id_bytes, device=0):
        self.rank, self.world_size, self.device = int(rank), int(world_size), int(device)
        buf = (C.c_uint8 * L.COMM_ID_BYTES).from_buffer_copy(bytes(id_bytes))
        h = C.c_void_p()
        L.check(L.lib().bmo_comm_init(L.context(device), self.world_size, self.rank, buf, C.byref(h)))
        self.h = h

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * L.COMM_ID_BYTES)()
        L.check(L.lib().bmo_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch(cls, device=None, group=None):
        import torch.distributed as dist
        rank, ws = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(rank, ws, box[0], world()[2] if device is None else device)

    def nccl_version(self):
        v = C.c_int32(0)
        L.check(L.lib().bmo_comm_info(self.h, None, None, C.byref(v)))
        return int(v.value)

    def allreduce(self, field, n_complex=None, sync=True):
        """Sum `field` over all ranks in place (bmo_pd_allreduce).  field: complex128 / float64 numpy array (host, staged
        through the GPU by the library), or an integer device pointer together with n_complex."""
        if isinstance(field, int):
            L.check(L.lib().bmo_pd_allreduce(self.h, C.c_void_p(field), int(n_complex), L.INPUT_DEVICE | (L.COMM_SYNC if sync else 0)))
            return field
        assert field.dtype in (np.complex128, np.float64) and (field.flags["C_CONTIGUOUS"] or field.flags["F_CONTIGUOUS"])
        n = field.size if field.dtype == np.complex128 else field.size // 2
        L.check(L.lib().bmo_pd_allreduce(self.h, L.ptr(field), int(n), 0))
        return field

    def free(self):
        if self.h:
            L.lib().bmo_comm_free(self.h)
            self.h = None


_default_comm = None


def default_comm(group=None):
    """The process-wide FieldComm (created on first use from the torch.distributed group; NCCL backends only)."""
    global _default_comm
    if _default_comm is None:
        _default_comm = FieldComm.from_torch(group=group)
    return _default_comm


def world():
    """(rank, world_size, local_rank) of this process; (0, 1, 0) outside torchrun."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_slice(n, rank, world_size):
    """Contiguous slice [start, stop) of rank `rank`: sizes differ by at most one, order preserved
    (sequential lens stacks: equal count = equal work)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_interleaved(n, rank, world_size):
    """Round-robin indices rank, rank + W, ... for branching workloads (C4): the branch depth varies
    with the ray position, so interleaving equalises the work where contiguous blocks would not."""
    return np.arange(rank, n, world_size)


def shard_bundle(bundle, rank, world_size, interleaved=False):
    """This rank's part of a RayBundle / BeamletBundle (+ the global indices it holds)."""
    n = len(bundle)
    idx = shard_interleaved(n, rank, world_size) if interleaved else np.arange(*shard_slice(n, rank, world_size))
    if isinstance(bundle, bm.RayBundle):
        part = bm.RayBundle(bundle.pos[idx], bundle.dir[idx], bundle.lam[idx], None if bundle.E0 is None else bundle.E0[idx], normalize=False)
    elif isinstance(bundle, bm.BeamletBundle):
        part = bm.BeamletBundle(bundle.rays[idx], bundle.lam[idx], bundle.w0[idx], bundle.E0[idx])
    else:
        raise TypeError(f"cannot shard {type(bundle).__name__}")
    return part, idx


def allreduce_field(field, group=None, comm=None):
    """Sum the ranks' partial Photodetector fields in place.  With a FieldComm (or an NCCL process group, for which the
    process-wide FieldComm is used) the sum runs through the C ABI (bmo_pd_allreduce: NCCL inside libbmo.so); the
    torch.distributed path remains for gloo (CPU tests).  `field`: complex128 numpy array or a torch tensor (float64 view
    of re/im pairs, or complex128) that lives where the backend wants it."""
    import torch
    import torch.distributed as dist
    if comm is None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl":
        comm = default_comm(group)
    if comm is not None:
        if isinstance(field, np.ndarray):
            return comm.allreduce(field)
        t = torch.view_as_real(field) if field.is_complex() else field
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float64
        torch.cuda.current_stream(t.device).synchronize()      # the library reduces on its own stream
        comm.allreduce(int(t.data_ptr()), t.numel() // 2, sync=True)
        return field
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return field
    if isinstance(field, np.ndarray):
        assert field.dtype == np.complex128
        flat = np.ascontiguousarray(field.T if field.flags["F_CONTIGUOUS"] and field.ndim == 2 else field)
        t = torch.from_numpy(flat.view(np.float64).reshape(-1))
        if dist.get_backend(group) == "nccl":
            g = t.cuda()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            t.copy_(g.cpu())
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out = flat.T if flat is not field and field.ndim == 2 and field.flags["F_CONTIGUOUS"] else flat
        if out is not field:
            field[...] = out
        return field
    t = torch.view_as_real(field) if field.is_complex() else field
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return field


def gather_rows(local, idx, n_total, group=None):
    """Assemble per-ray outputs (e.g. Spotdetector hits) of all ranks in global ray order on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (np.asarray(idx), np.asarray(local)), group=group)
    out = np.full((n_total,) + np.asarray(local).shape[1:], np.nan, dtype=np.asarray(local).dtype)
    for i, v in parts:
        out[i] = v
    return out


def solve_system_sharded(system, bundle, r_max=100, interleaved=False, keep_segments=False, group=None, comm=None):
    """solve_system! of a bundle across all ranks: this rank traces its shard on GPU LOCAL_RANK; every
    Photodetector of `system` ends up with the field of the *whole* bundle on every rank (all-reduce);
    Spotdetector data stay rank-local (disjoint slices).  Returns (TraceResult of the shard, indices)."""
    from . import components as co
    from .solver import solve_system_
    rank, ws, local = world()
    part, idx = shard_bundle(bundle, rank, ws, interleaved)
    pds = [o for o in system.leaves() if isinstance(o, co.Photodetector)]
    before = [pd.field.copy() for pd in pds]
    for pd in pds:                       # accumulate this rank's beamlets onto a zeroed partial field
        pd.field[...] = 0
    res = solve_system_(system, part, r_max=r_max, device=local, keep_segments=keep_segments) if len(part) else None
    for pd, old in zip(pds, before):
        allreduce_field(pd.field, group, comm)
        pd.field += old                  # the reference's `+=` onto whatever the detector already held
    return res, idx
