"""CPU, world_size 2 over gloo: the host logic of the N > 1 path (frame sharding, variable-length key
all-gather).  The kernels themselves are single-GPU and covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG


def _worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rt3d import multi
    rng = np.random.default_rng(100 + rank)
    keys = torch.from_numpy(np.unique(rng.integers(0, 5000, size=300 + 700 * rank)).astype(np.int64))
    parts = multi.all_gather_varlen(keys)
    merged = torch.unique(multi.box_wide_keys(keys))      # stand-in for ops.keys_sort_unique on the CPU
    start, stop = multi.shard_range(257, rank, world)
    out[rank] = (keys.tolist(), [p.tolist() for p in parts], merged.tolist(), (start, stop))
    empty = multi.all_gather_varlen(torch.empty(0, dtype=torch.int64) if rank == 0 else keys[:3])
    assert [p.numel() for p in empty] == [0, 3]
    # a step's worth of segments (one per rig frame) in two collectives
    lens = [5 + rank, 0, 17 * (rank + 1)]
    flat = torch.arange(sum(lens), dtype=torch.int64) + 1000 * rank
    ragged = multi.all_gather_ragged(flat, lens)
    for r, segs in enumerate(ragged):
        want_lens = [5 + r, 0, 17 * (r + 1)]
        assert [s.numel() for s in segs] == want_lens
        assert torch.equal(torch.cat(segs), torch.arange(sum(want_lens), dtype=torch.int64) + 1000 * r)
    dist.barrier()
    dist.destroy_process_group()


def test_varlen_all_gather_and_sharding():
    world, port = 2, 29613
    mgr = mp.get_context("spawn").Manager()   # no fork() of this multi-threaded process
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    own = [out[r][0] for r in range(world)]
    for r in range(world):
        assert out[r][1] == own                                     # every rank sees every rank's keys, in rank order
        assert out[r][2] == sorted(set(own[0]) | set(own[1]))       # union, sorted, identical on all ranks
    ranges = [out[r][3] for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 257 and ranges[0][1] == ranges[1][0]
    assert abs((ranges[0][1] - ranges[0][0]) - (ranges[1][1] - ranges[1][0])) <= 1


def test_shard_range_covers_everything():
    from rt3d import multi
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [multi.stream_owner(c, 4) for c in range(8)] == [0, 1, 2, 3, 0, 1, 2, 3]
