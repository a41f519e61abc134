"""N>1 host logic on CPU: two `gloo` ranks shard a batch the way bench.py / the library
do (contiguous instance ranges, no data-path collective), each materialises its shard
from the same plan bytes (numpy plan emulator standing in for the device), and the
gathered result must equal the oracle on the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ninst, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plan_emulator import Emulator
    from signalops import Amplify, Filt, Lowpass, Signal, dB, kHz
    from signalops.lowering import lower
    from signalops.sharding import shard_range
    lo, hi = shard_range(ninst, rank, world)
    outs = []
    blobs = set()
    for i in range(lo, hi):
        x = np.random.default_rng(100 + i).standard_normal((2000, 2))
        plan = lower(Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB))
        blobs.add(plan.tobytes())
        outs.append(Emulator(plan.tobytes()).run(plan.input_arrays)[0])
    assert len(blobs) <= 1                          # one plan for the whole shard
    mine = torch.from_numpy(np.stack(outs)) if outs else torch.zeros((0, 2000, 2), dtype=torch.float64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.shape[0]]))          # bookkeeping only
    elapsed = torch.tensor([float(rank + 1)])
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)                 # "max over ranks" timing rule
    dist.barrier()
    ret[rank] = (lo, hi, mine.numpy(), [int(s) for s in sizes], float(elapsed))
    dist.destroy_process_group()


def test_two_rank_batch_shard_matches_oracle():
    import oracle
    from signalops import Amplify, Filt, Lowpass, Signal, dB, kHz
    world, ninst = 2, 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ninst, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]
    covered = []
    for r in range(world):
        lo, hi, data, sizes, tmax = ret[r]
        assert sizes == [ret[q][1] - ret[q][0] for q in range(world)] and tmax == float(world)
        for k, i in enumerate(range(lo, hi)):
            x = np.random.default_rng(100 + i).standard_normal((2000, 2))
            want, _ = oracle.sink(Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB))
            assert np.max(np.abs(data[k] - want)) <= 1e-9 * np.sqrt(np.mean(want ** 2))
            covered.append(i)
    assert covered == list(range(ninst))


def test_shard_range_partitions_exactly():
    from signalops.sharding import shard_range
    for n in (0, 1, 7, 256, 4096):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
