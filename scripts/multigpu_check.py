"""torchrun --nproc-per-node N scripts/multigpu_check.py : the sharded solve (one rank per GPU, NCCL all-reduce of the
detector field) reproduces the single-GPU solve of the whole bundle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import __graft_entry__ as ge
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    ge.build_libbmo()
dist.barrier()
m = ge.load_package()
from tests import scenes, scenes2 as s2
# C3: beamlet lattice -> Photodetector, field of the whole bundle on every rank
k, n = 12, 160
lat = s2.beamlet_lattice(k)
bundle = m.BeamletBundle.from_params(lat["pos"], lat["dir"], lat["lam"], lat["w0"], M2=lat["M2"], P0=lat["P0"], support=lat["support"])
sc = s2.expander(m, n)
res, idx = m.solve_system_sharded(sc["system"], bundle)
sharded = sc["pd"].field.copy()
sc1 = s2.expander(m, n)
m.solve_system_(sc1["system"], bundle, device=local)
single = sc1["pd"].field
rel = np.linalg.norm((sharded - single).ravel()) / np.linalg.norm(single.ravel())
# C2: ray slices, spot diagram gathered in ray order
nr = 10007
pos, d = scenes.fibonacci_disc(nr)
rb = m.RayBundle(pos, d, 707e-9)
sc2 = scenes.doublet_spot(m)
r2, idx2 = m.solve_system_sharded(sc2["system"], rb)
spots = m.parallel.gather_rows(sc2["spot"].data, idx2, nr)
sc3 = scenes.doublet_spot(m)
m.solve_system_(sc3["system"], rb, device=local, keep_segments=False)
same = np.array_equal(spots, sc3["spot"].data)
ok = torch.tensor([float(rel <= 1e-12 and same)], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multigpu_check world={ws}: detector field sharded vs single rel L2 = {rel:.2e}; spot diagram identical: {same}; all ranks ok: {bool(ok.item())}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
