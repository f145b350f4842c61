"""Partition of the iteration range over ranks: the reference's `-j jobnum` rule, and the sharded run built on it.

A job tests `iterations` consecutive streams starting at byte jobnum * n * iterations / 8 of the data
file (utilities.c:1500-1527, README.md:129-213), i.e. every job gets the same count and job r follows
job r-1.  With G ranks and I total iterations each rank gets I // G (the remainder is not tested, as in
the reference where it needs one more job)."""
import numpy as np


def shard_range(total_iterations, rank, nranks):
    per = total_iterations // nranks
    return rank * per, (rank + 1) * per


def run_sharded(eng, total, rank, world, batch, collect=True, first_stream=0, inflight=None, timer=None, generator=1):
    """BASELINE config 3: `total` streams of tools/generators `generator` (1 = LCG, the default; 2, 4, 5 also jump ahead:
    stsgpu_generate), made on the device by skip-ahead from `first_stream`, rank r testing the `-j r` slice in batches of `batch` streams with the engine in gather mode
    (stsgpu_gather_begin: every batch ends with the exchange of its p-values to rank 0 over NCCL).

    Returns, on rank 0 with collect=True, the p-values of all world * (total // world) streams in iteration (= job)
    order, as the reference's `-m a` sees them after reading every job's .pvalues file; otherwise None.
    `timer`, if given, is called right before the first submit and right after the last wait (host clocks around the
    pipelined region; the caller adds its barriers)."""
    lo, hi = shard_range(total, rank, world)
    per = hi - lo
    assert per % batch == 0 and per > 0, "the slice of a rank must be whole batches"
    nb = per // batch
    depth = inflight or eng.params.pipeline_depth or 1
    out = np.empty((world * per, eng.layout.npv), dtype=np.float64) if (collect and rank == 0) else None
    eng.gather_begin(0)
    done = 0

    def finish():
        nonlocal done
        eng.wait()
        g = eng.gather_pvals(0, world, copy=False)
        if out is not None:
            for r in range(world):		# rank r's batch `done` holds iterations r*per + done*batch ...
                out[r * per + done * batch:r * per + (done + 1) * batch] = g[r * batch:(r + 1) * batch]
        done += 1

    if timer:
        timer()
    pending = 0
    for b in range(nb):
        eng.generate(generator, first_stream + lo + b * batch, batch)
        eng.submit_resident(batch, lo + b * batch)
        pending += 1
        if pending == depth:
            finish()
            pending -= 1
    while pending:
        finish()
        pending -= 1
    if timer:
        timer()
    eng.gather_begin(-1)
    return out
