"""CPU, world_size 2 over gloo: the multi-GPU host logic - the `-j` style partition of the iteration
range (utilities.c:1526) and the rank-major gather order - reproduces the single-process result.
The per-rank compute is stood in for by the oracle (this is a test of the plumbing, on CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, out):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from sts_b200.sharding import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    n = 1 << 17
    prm = O.make_params(n=n, apen_m=5, serial_m=8)
    raw = O.lcg_bytes(lo, hi - lo, n)                       # rank r reads only its own slice of the file
    pv, st, w, lay = O.run(prm, raw, hi - lo, nthreads=1)
    t = torch.from_numpy(pv)
    gathered = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        np.save(out, torch.cat(gathered, 0).numpy())
    # the assessment needs no gather at all: per-column tallies are additive, one small all-reduce
    a = O.assess(lay, pv, 4)
    tally = torch.from_numpy(np.concatenate([a["sample"][:, None], a["toolow"][:, None], a["freq"]], axis=1).copy())
    dist.all_reduce(tally, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out + ".tally.npy", tally.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, oracle):
    from sts_b200.sharding import shard_range
    total, world = 4, 2
    assert [shard_range(total, r, world) for r in range(world)] == [(0, 2), (2, 4)]
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]   # -j: iterations // jobs each
    out = str(tmp_path / "g.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    got = np.load(out)
    n = 1 << 17
    prm = oracle.make_params(n=n, apen_m=5, serial_m=8)
    pv, st, w, lay = oracle.run(prm, oracle.lcg_bytes(0, total, n), total, nthreads=2)
    assert got.shape == pv.shape and (got == pv).all()
    a = oracle.assess(lay, pv, 4)
    tally = np.load(out + ".tally.npy")
    assert (tally[:, 0] == a["sample"]).all() and (tally[:, 1] == a["toolow"]).all() and (tally[:, 2:] == a["freq"]).all()
