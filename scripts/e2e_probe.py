"""timeout 120 torchrun --nproc-per-node N scripts/e2e_probe.py  (ALWAYS under `timeout`).
The end-to-end C2 step (bmo_trace_rays_spots with host buffers) of every rank at once, with the input buffer in ordinary pinned
memory and in write-combined pinned memory, next to the raw copy rates of both kinds of memory: where does the time go when
several ranks share the host?  One line per variant on rank 0 (max over ranks)."""
import ctypes as C
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as ge
m = ge.load_package()
from bmo_b200 import _lib as L
from tests import scenes
n = 1 << 20
stream = torch.cuda.current_stream()
L.set_stream(stream.cuda_stream, local)
sc = scenes.doublet_spot(m)
dsys = m.upload_system(sc["system"], [707e-9], device=local)
pos, d = scenes.fibonacci_disc(n)
cudart = C.CDLL([p for p in (os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12"), "libcudart.so.12", "libcudart.so") if p.startswith("lib") or os.path.exists(p)][0])
cudart.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = cudart.cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, rc
    return p


def as_array(p, shape, dtype):
    nb = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * nb).from_address(p.value), dtype=dtype).reshape(shape)


PINNED, WC = 0, 4                                   # cudaHostAllocDefault, cudaHostAllocWriteCombined
bufs = {}
for name, flags in (("pinned", PINNED), ("write-combined", WC)):
    p = host_alloc(pos.nbytes, flags); as_array(p, pos.shape, np.float64)[...] = pos
    bufs[name] = p
dir1 = host_alloc(24, PINNED); as_array(dir1, (3,), np.float64)[...] = d[0]
xz = host_alloc(n * 16, PINNED)
dev_in = torch.empty(pos.nbytes, dtype=torch.uint8, device="cuda"); dev_out = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
s2 = torch.cuda.Stream()


def step(inp):
    h = C.c_void_p()
    L.check(L.lib().bmo_trace_rays_spots(dsys.h, n, inp, dir1, None, None, None, 100, L.UNIFORM_DIR, None, xz, C.byref(h)))
    L.lib().bmo_result_free(h)


def copies(inp, h2d=True, d2h=True):
    if h2d: cudart.cudaMemcpyAsync(C.c_void_p(dev_in.data_ptr()), inp, pos.nbytes, 1, C.c_void_p(stream.cuda_stream))
    if d2h: cudart.cudaMemcpyAsync(xz, C.c_void_p(dev_out.data_ptr()), n * 16, 2, C.c_void_p(s2.cuda_stream))


def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if ws > 1: dist.barrier()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([sum(ts) / len(ts) * 1e3], device="cuda")
    if ws > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name in ("pinned", "write-combined"):
    inp = bufs[name]
    r = {"e2e step": timed(lambda: step(inp)), "H2D 25 MB": timed(lambda: copies(inp, True, False)), "D2H 16.8 MB": timed(lambda: copies(inp, False, True)),
         "H2D + D2H": timed(lambda: copies(inp, True, True))}
    if rank == 0:
        print(f"[{ws} ranks] input in {name:15s}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in r.items()), flush=True)
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
