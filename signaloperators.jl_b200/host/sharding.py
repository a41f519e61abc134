"""Batch sharding: whole signals -> ranks/devices, contiguous index
ranges, no data-path collective.  The same split `sigops_plan_run` uses across the
devices of one context (`ninst*d/nd .. ninst*(d+1)/nd`)."""


def shard_range(ninst, rank, world):
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return ninst * rank // world, ninst * (rank + 1) // world
