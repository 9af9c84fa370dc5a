"""Time the two Photodetector kernels (fast vs reference order) on the C3 scene.  GPU only."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build_libbmo()
m = ge.load_package()
from bmo_b200 import _lib as L
from tests import scenes2 as s2
import ctypes as C
import torch
k, n = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sc = s2.expander(m, n)
lat = s2.beamlet_lattice(k, aperture=8e-3 * k / 256)
b = m.BeamletBundle.from_params(lat["pos"], lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=1e-3 / 65536, support=lat["support"])
dsys = m.upload_system(sc["system"], [lat["lam"]])
res = m.trace_beamlets(dsys, b.rays, np.zeros(k * k, np.int32), b.w0, b.E0)
pdi = dsys.flat.object_index(sc["pd"])
f = torch.zeros(n * n * 2, dtype=torch.float64, device="cuda")
for name, flags in (("fast", L.INPUT_DEVICE), ("reference_order", L.INPUT_DEVICE | L.PD_REFERENCE_ORDER)):
    for it in range(3):
        L.counters_reset()
        L.check(L.lib().bmo_pd_accumulate(dsys.h, res.h, pdi, 0, C.c_void_p(f.data_ptr()), flags))
        c = L.counters()
    print(f"{name:16s} k={k} n={n}: kernel {c['pd_field_ms']:.3f} ms, call {c['pd_ms']:.3f} ms, {c['px_beamlets'] / (c['pd_field_ms'] * 1e-3):.4g} px-beamlets/s")
