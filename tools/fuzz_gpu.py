#!/usr/bin/env python
"""One-off wider fuzz of the GPU paths against the oracle / zlib (development tool; the committed tests use fixed seeds).
usage: fuzz_gpu.py [first_seed] [n_seeds]"""
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from compu_b200 import batch  # noqa: E402
from helpers import assert_inflate_parity, fuzz_cases, make_data, oracle_inflate  # noqa: E402


def main():
    s0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    alice = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read()
    n_inf = n_def = 0
    for seed in range(s0, s0 + ns):
        for wbits in (15, 31, -15, 47):
            datas, streams, caps = fuzz_cases(seed * 7 + wbits, wbits, 150, alice,
                                              sizes=(0, 1, 2, 5, 100, 1000, 5000, 20000, 65536, 70000, 300000))
            ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
            outs, st, lens, cons = batch.inflate_batch(streams, caps, wbits)
            assert_inflate_parity(outs, st, ref_outs, ref_st, "seed %d wbits %d" % (seed, wbits))
            n_inf += len(streams)
        rng = random.Random(seed)
        units = [make_data(rng, rng.randrange(5), rng.choice([0, 1, 7, 300, 5000, 70000, 1 << 20, 3 << 20]), alice) for _ in range(24)]
        for wb in (15, 31, -15):
            level, strategy = rng.choice([0, 1, 3, 6, 6, 6, 9]), rng.choice([0, 0, 0, 1, 2, 3, 4])
            comp, st = batch.deflate_batch(units, level=level, window_bits=wb, strategy=strategy,
                                           segment_bytes=rng.choice([0, 65536, 1 << 18, 1 << 20]))
            assert (st == 2).all()
            for c, u in zip(comp, units):
                assert zlib.decompress(c, wb) == u, "seed %d wb %d level %d strategy %d" % (seed, wb, level, strategy)
            outs, ist, _, _ = batch.inflate_batch(comp, [len(u) for u in units], wb)
            assert (ist == 2).all() and outs == units
            n_def += len(units)
    print("fuzz ok: %d inflate cases vs oracle, %d deflate units round-tripped (zlib + GPU decode)" % (n_inf, n_def))


if __name__ == "__main__":
    main()
