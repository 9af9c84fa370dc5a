import sys; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as ge
m = ge.load_package()
from oracle import oracle as orc
from tests import scenes2 as s2
sc, osc = s2.mesh_scene(m), s2.mesh_scene_oracle()
pos, d, E0 = s2.jittered_lattice(96)
for i in (20, 21):
    bundle = m.RayBundle(pos[i:i+1], d, 1e-6, E0=E0)
    res = m.solve_system_(sc["system"], bundle, r_max=100)
    b, seg = res.beams(), res.segments()
    print('GPU ray', i)
    for bi in range(res.n_beams):
        f, ns = int(b["first"][bi]), int(b["nseg"][bi])
        print('  beam', bi, 'parent', b['parent'][bi], 'slot', b['slot'][bi], 'status', b['status'][bi], 'obj', seg['obj'][f:f+ns].tolist(), 't', np.round(seg['t'][f:f+ns],6).tolist())
    print('  bfs', res.bfs_order())
    ob = orc.polarized_beam(pos[i], d, 1e-6, E0)
    orc.solve_system_(osc["system"], ob)
    print('ORC')
    for t in orc.beam_export(osc["system"], ob):
        print('  parent', t['parent'], 'obj', t['rays']['obj'].tolist(), 't', np.round(t['rays']['t'],6).tolist())
