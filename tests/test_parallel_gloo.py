"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo process group (127.0.0.1)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_bundle(bmo):
    from bmo_b200 import parallel as par
    for n in (0, 1, 7, 8, 1 << 20, 10_000_019):
        for w in (1, 2, 3, 8):
            sl = [par.shard_slice(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(sl[r][1] == sl[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1
            if n <= 1 << 20:
                il = np.concatenate([par.shard_interleaved(n, r, w) for r in range(w)])
                assert np.array_equal(np.sort(il), np.arange(n))


def test_shard_bundle_keeps_rows(bmo):
    from bmo_b200 import parallel as par
    rng = np.random.default_rng(0)
    n = 37
    rb = bmo.RayBundle(rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), 1e-6)
    got = np.full((n, 3), np.nan)
    for r in range(4):
        part, idx = par.shard_bundle(rb, r, 4, interleaved=True)
        assert np.array_equal(part.dir, rb.dir[idx])          # not re-normalised: bit-identical rows
        got[idx] = part.pos
    assert np.array_equal(got, rb.pos)
    bb = bmo.BeamletBundle.from_params(rng.normal(size=(n, 3)), (0.0, 1.0, 0.0), 1e-6, 1e-3)
    part, idx = par.shard_bundle(bb, 1, 2)
    assert np.array_equal(part.rays, bb.rays[idx]) and np.array_equal(part.E0, bb.E0[idx])


def _worker(rank, world_size, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from bmo_b200 import parallel as par
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    rng = np.random.default_rng(100 + rank)
    n = 24
    partial = np.asfortranarray(rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))   # pd.field layout: column-major
    mine = partial.copy()
    par.allreduce_field(partial)
    np.save(os.path.join(tmp, f"sum{rank}.npy"), partial)
    np.save(os.path.join(tmp, f"part{rank}.npy"), mine)
    # per-ray outputs come back in global ray order on every rank
    total = 11
    idx = par.shard_interleaved(total, rank, world_size)
    rows = np.stack([idx * 1.0, idx * 10.0 + rank], axis=1)
    full = par.gather_rows(rows, idx, total)
    np.save(os.path.join(tmp, f"rows{rank}.npy"), full)
    assert par.world() == (rank, world_size, rank)
    dist.barrier()
    dist.destroy_process_group()


def test_field_allreduce_and_gather_world2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(2)]
    for r in range(2):
        got = np.load(tmp_path / f"sum{r}.npy")
        assert np.array_equal(got, parts[0] + parts[1])        # two addends: order-independent, bit-exact
        rows = np.load(tmp_path / f"rows{r}.npy")
        assert np.array_equal(rows[:, 0], np.arange(11.0))
        assert np.array_equal(rows[:, 1], np.arange(11) * 10.0 + (np.arange(11) % 2))
