"""Partition of the iteration range over ranks: the reference's `-j jobnum` rule.

A job tests `iterations` consecutive streams starting at byte jobnum * n * iterations / 8 of the data
file (utilities.c:1500-1527, README.md:129-213), i.e. every job gets the same count and job r follows
job r-1.  With G ranks and I total iterations each rank gets I // G (the remainder is not tested, as in
the reference where it needs one more job)."""


def shard_range(total_iterations, rank, nranks):
    per = total_iterations // nranks
    return rank * per, (rank + 1) * per
