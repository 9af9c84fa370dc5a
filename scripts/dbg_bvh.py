import sys; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as ge
m = ge.load_package()
from oracle import oracle as orc
from tests import scenes2 as s2
v, f = s2.uv_sphere_f32(50.0, 96, 96)
ball = m.Mirror(m.Mesh(v, f, f32=True)); ball.translate3d_([0.0, 0.3, 0.0]); ball.xrotate3d_(0.3)
oball = orc.new("Mirror", ih=[orc.mesh(v, f, f32=True)]); oball.translate3d_([0.0, 0.3, 0.0]); oball.xrotate3d_(0.3)
ov = oball.shape().eval("mesh_vertices", nout=3*v.shape[0])
print('vertex parity', np.array_equal(ov, ball.shape.vertices.ravel()), np.abs(ov-ball.shape.vertices.ravel()).max())
rng = np.random.default_rng(1)
n = 4096
pos = np.zeros((n, 3)); pos[:, [0, 2]] = (rng.random((n, 2)) - 0.5) * 0.12
d = np.tile([0.0, 1.0, 0.0], (n, 1)); d[:, [0, 2]] += (rng.random((n, 2)) - 0.5) * 0.05
bundle = m.RayBundle(pos, d, 1e-6)
res = m.solve_system_(m.System([ball]), bundle, r_max=3)
ref = orc.bulk_trace_rays(orc.system([oball]), bundle.pos, bundle.dir, 1e-6, r_max=3, max_seg=4)
b, seg = res.beams(), res.segments()
first = b["first"]
t_gpu, t_ref = seg["t"][first], ref["seg"][:, 0, 7]
hit = np.isfinite(t_ref)
dt = np.abs(t_gpu[hit]-t_ref[hit]); dn = np.abs(seg["nrm"][first][hit]-ref["seg"][:,0,8:11][hit]).max(axis=1)
print('n hit', hit.sum(), 't differs', (dt>0).sum(), 'max dt', dt.max(), 'normal differs', (dn>0).sum(), 'max dn', dn.max())
idx = np.where(dt>0)[0][:5]; print(dt[idx], dn[idx])
