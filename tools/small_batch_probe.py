#!/usr/bin/env python
"""Latency of SMALL batches through cz_inflate_batch (1 / 16 / 100 / 1 000 zlib streams of 64 KiB, best of 5, the Python
wrapper's packing included). Development tool; run on a B200 from the repo root."""
import time, zlib, sys, os
sys.path.insert(0, os.getcwd())
from compu_b200 import batch
alice = open("tests/golden/alice29.txt", "rb").read() * 50
for n in (1, 16, 100, 1000):
    datas = [alice[i * 65536:(i + 1) * 65536] for i in range(n)]
    streams = [zlib.compress(d, 6) for d in datas]
    caps = [65536] * n
    best = 1e9
    for it in range(5):
        t0 = time.perf_counter()
        outs, st, lens, cons = batch.inflate_batch(streams, caps, 15)
        best = min(best, time.perf_counter() - t0)
    assert outs == datas and (st == 2).all()
    print("n=%d streams of 64 KiB: %.2f ms (python wrapper included)" % (n, best * 1e3))
