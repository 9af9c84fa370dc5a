"""C4 (mesh scene, PolarizedRays, branching) at a given size: host-profile of the wave loop; run under ncu for the launch list."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
m = ge.load_package()
from bmo_b200 import _lib as L
from tests import scenes2 as s2
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = s2.mesh_scene(m)
pos, d, E0 = s2.jittered_lattice(nr)
dsys = m.upload_system(sc["system"], [1e-6])
lam = np.zeros(nr, np.int32)
Ef = np.ascontiguousarray(np.broadcast_to(E0, (nr, 3)))
dd = np.ascontiguousarray(np.broadcast_to(d, (nr, 3)))
for it in range(reps):
    L.counters_reset()
    t0 = time.perf_counter()
    r = m.trace_rays(dsys, pos, dd, lam, Ef, None, 100, keep_segments=bool(int(os.environ.get("KEEP", "0"))))
    wall = time.perf_counter() - t0
    c = L.counters()
    print(f"rep {it}: rays {nr} beams {r.n_beams} interactions {r.interactions} waves {r.waves} trace_ms {c['trace_ms']:.1f} wall {wall*1e3:.1f} ms "
          f"k1_ms {c['trace_step_ms']:.1f} ({c['trace_step_launches']} launches) tri {c['tri_tests']} sdf {c['sdf_evals']}", flush=True)
    r.free()
