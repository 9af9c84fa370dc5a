#!/usr/bin/env python
"""Per-call wall time of bmo_retrace / bmo_trace_rays (segments kept) on the C2 bundle, with the stream-ordered pool's
reserved / used bytes after each call.  Usage: python scripts/retrace_probe.py [n_iter]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

m = ge.load_package()
from tests import scenes  # noqa: E402

rt = C.CDLL("libcudart.so")
pool = C.c_void_p()
rt.cudaDeviceGetDefaultMemPool(C.byref(pool), 0)


def pool_stat():
    out = []
    for attr in (5, 7):   # cudaMemPoolAttrReservedMemCurrent, cudaMemPoolAttrUsedMemCurrent
        v = C.c_uint64()
        rt.cudaMemPoolGetAttribute(pool, attr, C.byref(v))
        out.append(v.value / 2**20)
    return out


n = 1 << 20
sc = scenes.doublet_spot(m)
dsys = m.upload_system(sc["system"], [707e-9])
pos, d = scenes.fibonacci_disc(n)
pos_d, dir_d = torch.from_numpy(pos).cuda(), torch.from_numpy(d).cuda()
lam_d = torch.zeros(n, dtype=torch.int32, device="cuda")
prev = m.trace_rays(dsys, (pos_d.data_ptr(), n), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100, keep_segments=True, device_inputs=True)
prev.keep = True
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in ("fresh", "retrace", "fresh", "retrace"):
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        flush.fill_(it)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "fresh":
            r = m.trace_rays(dsys, (pos_d.data_ptr(), n), dir_d.data_ptr(), lam_d.data_ptr(), None, None, 100, keep_segments=True, device_inputs=True)
        else:
            r = m.retrace(dsys, prev, 100, keep_segments=True)
        t1 = time.perf_counter()
        r.free()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res, used = pool_stat()
        print(f"{name:8s} {it:2d}  call {1e3 * (t1 - t0):8.3f} ms  free+sync {1e3 * (t2 - t1):8.3f} ms  pool reserved {res:9.1f} MiB used {used:9.1f} MiB", flush=True)
