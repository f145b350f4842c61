#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun, one rank per GPU): every rank tests its `-j` slice of the LCG
stream sequence, the p-values are gathered to rank 0 over NCCL, and rank 0 verifies that the gathered
array equals what a single engine computes for all the streams, in job order."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sts_b200  # noqa: E402
from sts_b200.sharding import shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lo, hi = shard_range(total, rank, world)
B = hi - lo
eng = sts_b200.Engine(device=local, max_streams=max(B, 1))
uid = [sts_b200.Engine.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(uid[0], rank, world)
eng.generate_lcg(lo, B)
eng.submit_resident(B, lo)
eng.wait()
eng.comm_barrier()
allpv = eng.gather_pvals(0, world)
ok = True
if rank == 0:
    assert allpv.shape == (world * B, eng.layout.npv), allpv.shape
    for r in range(world):
        a, b = shard_range(total, r, world)
        eng.generate_lcg(a, b - a)
        eng.submit_resident(b - a, a)
        eng.wait()
        pv, _, _ = eng.results()
        same = (allpv[r * B:(r + 1) * B] == pv).all()
        print("rank %d slice [%d,%d): gathered == local recompute: %s" % (r, a, b, same))
        ok &= bool(same)
    print("MULTI-GPU OK" if ok else "MULTI-GPU MISMATCH")
eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
