import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
m = ge.load_package()
from oracle import oracle as orc
from tests import scenes2 as s2
n = 1 << 21
sc = s2.mesh_scene(m)
pos, d, E0 = s2.jittered_lattice(n)
res = m.solve_system_(sc["system"], m.RayBundle(pos, d, 1e-6, E0=E0), r_max=100)
b, seg = res.beams(), res.segments()
bad_rows = np.nonzero(np.isnan(seg["E0"]).any(axis=1))[0]
print("rows with NaN E0:", len(bad_rows), "of", len(seg["t"]))
# map rows to beams -> roots
first = b["first"]; nb = res.n_beams
beam_of_row = np.searchsorted(first, bad_rows, side="right") - 1
roots = set()
for bi in np.unique(beam_of_row)[:2000]:
    r = bi
    while b["parent"][r] >= 0: r = b["parent"][r]
    roots.add(int(r))
roots = sorted(roots)
print("distinct root rays:", len(roots), roots[:10])
osc = s2.mesh_scene_oracle()
for r in roots[:3]:
    print("root", r, "pos", pos[r])
    ob = orc.polarized_beam(pos[r], d, 1e-6, E0)
    orc.solve_system_(osc["system"], ob)
    tree = orc.beam_export(osc["system"], ob)
    for t in tree:
        print("  oracle beam: obj", t["rays"]["obj"], "t", t["rays"]["t"], "E0 nan:", np.isnan(t["rays"]["E0"]).any(axis=1))
    # gpu beams of this root
    for bi in range(nb):
        pass
    ids = [r]
    k = 0
    while k < len(ids):
        ids += [int(c) for c in np.nonzero(b["parent"] == ids[k])[0]]
        k += 1
    for bi in ids:
        f, ns = int(first[bi]), int(b["nseg"][bi])
        print("  gpu beam", bi, "obj", seg["obj"][f:f+ns], "t", seg["t"][f:f+ns], "E0 nan:", np.isnan(seg["E0"][f:f+ns]).any(axis=1), "dir", seg["dir"][f:f+ns][-1], "nrm", seg["nrm"][f:f+ns])
