#!/usr/bin/env python3
"""BASELINE config 3 on N GPUs, checked (run under torchrun, one rank per GPU; tests/test_gpu_multi.py does):
`total` LCG streams of 2^20 bits sharded over the ranks by the reference's `-j` rule (utilities.c:1526), every
batch's p-values gathered to rank 0 over NCCL (gather mode).  Rank 0 then verifies the gathered array, which is in
iteration (= job) order like the reference's `-m a` after reading every job's .pvalues file:
  * against a single engine on its own GPU testing all `total` streams: bit-identical;
  * against the CPU oracle on streams sampled from every rank's slice: p-values within 1e-9 relative
    (CumulativeSums: absolute 1e-13 on top, see tests/test_gpu_parity.py).
Prints one JSON line on rank 0; exit status 0 only if everything agreed.

usage: torchrun --nproc-per-node N tools/config3_check.py [--total 16384] [--batch 1024] [--oracle-streams 16]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sts_b200  # noqa: E402
from sts_b200.sharding import run_sharded, shard_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=16384)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--oracle-streams", type=int, default=16)
ap.add_argument("--generator", type=int, default=1, help="tools/generators number made on the device: 1, 2, 4 or 5")
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = args.batch
eng = sts_b200.Engine(device=local, max_streams=B, pipeline_depth=3)
uid = [sts_b200.Engine.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(uid[0], rank, world)

marks = []
torch.cuda.synchronize()
dist.barrier()
allpv = run_sharded(eng, args.total, rank, world, B, collect=True, timer=lambda: marks.append(time.perf_counter()),
                    generator=args.generator)
dist.barrier()
sec = torch.tensor([marks[1] - marks[0]], dtype=torch.float64, device="cuda")
dist.all_reduce(sec, op=dist.ReduceOp.MAX)
eng.close()

ok = True
if rank == 0:
    per = args.total // world
    n = 1 << 20
    assert allpv.shape == (world * per, 188), allpv.shape
    # (1) one engine, all streams, no communicator
    single = np.empty_like(allpv)
    with sts_b200.Engine(device=local, max_streams=B, pipeline_depth=2) as one:
        for b in range(world * per // B):
            one.generate(args.generator, b * B, B)
            one.submit_resident(B, b * B)
            one.wait()
            single[b * B:(b + 1) * B] = one.results(copy=False)[0]
        lay = one.layout
        # the sampled streams for the oracle come from the device generator (byte-identical to tools/generators 1, see
        # test_device_lcg_is_the_reference_generator): the CPU generator has no skip-ahead and would walk 10^10 steps
        k = max(world, args.oracle_streams)
        idx = sorted(set(int(x) for x in np.linspace(0, world * per - 1, k)) |
                     set(shard_range(args.total, r, world)[0] for r in range(world)) |		# first and last iteration of every job
                     set(shard_range(args.total, r, world)[1] - 1 for r in range(world)))
        raws = []
        for i in idx:
            one.generate(args.generator, i, 1)
            raws.append(one.download_input(1))
        raw = np.concatenate(raws)
    bad_rows = np.flatnonzero((allpv != single).any(axis=1))
    identical = bad_rows.size == 0
    # (2) the oracle on streams spread over all slices, in -j order
    from oracle import oracle as O
    opv = O.run(O.make_params(), raw, len(idx), nthreads=os.cpu_count() or 1)[0]
    got = allpv[idx]
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(got - opv) / np.abs(opv)
    rel[got == opv] = 0.0
    c3 = slice(lay.pv_off[3], lay.pv_off[3] + 2)
    rel[:, c3][np.abs(got[:, c3] - opv[:, c3]) <= 1e-13] = 0.0
    worst = float(rel.max())
    ok = identical and worst <= 1e-9
    print(json.dumps({"config": "BASELINE config 3", "generator": args.generator, "ranks": world, "total_streams": world * per, "batch": B,
                      "seconds_sharded": float(sec.item()), "gbit_per_s": world * per * n / float(sec.item()) / 1e9,
                      "gathered_identical_to_single_gpu": bool(identical), "first_bad_rows": bad_rows[:8].tolist(),
                      "oracle_streams": idx, "oracle_max_rel_err": worst, "ok": bool(ok)}))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
