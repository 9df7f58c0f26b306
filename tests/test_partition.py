"""Host-side logic of the multi-GPU path, on CPU: the byte-balanced partitioner (pure host arithmetic inside the C-ABI
library) and a world_size-2 `gloo` run of the sharding scheme bench.py / the batched entry points use — every rank takes its
contiguous shard, processes it independently (here with the CPU oracle standing in for the device), and the gather is a
concatenation in rank order with no data-path collective."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest

from compu_b200 import batch
from helpers import oracle_inflate


def test_partition_properties():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 7, 100, 5000):
        sizes = rng.integers(0, 1 << 20, n).astype(np.uint64)
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes)
        for parts in (1, 2, 3, 8):
            cuts = batch.partition_by_bytes(offs, parts)
            assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == parts + 1
            assert all(cuts[i] <= cuts[i + 1] for i in range(parts))
            if n >= 100:
                total = int(offs[-1])
                for p in range(parts):
                    share = int(offs[cuts[p + 1]]) - int(offs[cuts[p]])
                    assert abs(share - total / parts) <= (1 << 20) + 1, "shard %d of %d is unbalanced" % (p, parts)


def test_partition_skewed_sizes():
    # one huge unit among small ones: cuts stay monotone and cover everything exactly once
    sizes = np.array([10, 10, 1 << 30, 10, 10, 10], dtype=np.uint64)
    offs = np.zeros(len(sizes) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(sizes)
    cuts = batch.partition_by_bytes(offs, 4)
    covered = []
    for p in range(4):
        covered += list(range(cuts[p], cuts[p + 1]))
    assert covered == list(range(len(sizes)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(99)  # same on every rank: the job description is replicated, the work is sharded
        datas = [bytes(rng.integers(97, 105, int(rng.integers(1, 30000)), dtype=np.uint8)) for _ in range(64)]
        streams = [zlib.compress(d, 6) for d in datas]
        offs = np.zeros(len(streams) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in streams])
        cuts = batch.partition_by_bytes(offs, world)
        a, b = cuts[rank], cuts[rank + 1]
        outs, st, lens = oracle_inflate(streams[a:b], [len(d) for d in datas[a:b]], 15)
        # gather: lengths and status only (O(#units) scalars); the payload of a rank stays on that rank's side of the gather
        gathered = [None] * world
        dist.all_gather_object(gathered, (a, b, [int(x) for x in lens], [int(x) for x in st],
                                          [zlib.crc32(o) for o in outs]))
        dist.barrier()
        if rank == 0:
            cover, crcs = [], []
            for (ga, gb, glens, gst, gcrc) in gathered:
                cover += list(range(ga, gb))
                assert all(s == 2 for s in gst)
                crcs += gcrc
            ok = cover == list(range(len(streams))) and crcs == [zlib.crc32(d) for d in datas]
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
