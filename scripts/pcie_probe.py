"""timeout 120 torchrun --nproc-per-node N scripts/pcie_probe.py  (ALWAYS under `timeout`: a collective mismatch hangs every rank) : host<->device copy bandwidth per rank, alone and with all ranks copying
at once (pinned buffers, H2D and D2H on separate streams like the pipelined trace call)."""
import os, sys, time
import torch, torch.distributed as dist
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 64
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda"); d_out = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(mode, reps=20, together=True):
    torch.cuda.synchronize()
    if ws > 1 and together: dist.barrier()      # every rank calls run() in the "all" phase; in the "solo" phase only rank 0 does
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return MB * reps / 1024 / dt      # GiB/s per direction

for who in ("solo", "all"):
    for mode in ("h2d", "d2h", "both"):
        if who == "solo":
            v = run(mode, together=False) if rank == 0 else 0.0
            if ws > 1: dist.barrier()
        else:
            v = run(mode)
        t = torch.tensor([v], device="cuda")
        if ws > 1:
            lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); hi = t.clone(); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        else:
            lo = hi = t
        if rank == 0:
            print(f"{who:4s} {mode:4s}: rank0 {v:6.1f} GiB/s per direction" + (f"   (min {lo.item():.1f} max {hi.item():.1f} over ranks)" if who == "all" else ""), flush=True)
if ws > 1:
    dist.destroy_process_group()
