#!/usr/bin/env python
"""Fuzz of the block-parallel path on the CPU emulator (development tool; the fixed-seed cases are tests/test_sim_runs.py):
random data classes, containers, levels and chunk sizes; every clean stream must decode bit-exactly, every damaged or truncated
variant must either be declined or decode to exactly what zlib makes of it. CUSIM_TOKW=0 / 1 pins the phase A kernel.
usage: [CUSIM_TOKW=1] python tools/fuzz_sim_runs.py [rounds] [seed]"""
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import simlib  # noqa: E402
from helpers import make_data, zcomp  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    alice = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read()
    n = 0
    for it in range(rounds):
        d = make_data(rng, rng.randrange(5), rng.choice([70000, 150000, 300000]), alice)
        wb = rng.choice([15, 31, -15, 47])
        s = zcomp(d, rng.choice([1, 6, 9]), (31 if rng.random() < 0.5 else 15) if wb == 47 else wb)
        variants = [(s, True)]
        for _ in range(2):
            b = bytearray(s)
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            variants.append((bytes(b), False))
        variants.append((s[:rng.randrange(len(s) // 2, len(s))], False))
        outs, ok, _, _, _ = simlib.sim_inflate_runs([v for v, _ in variants], [len(d)] * len(variants), wb,
                                                    chunk_bytes=rng.choice([4096, 8192]), seed=it)
        for (v, clean), o, k in zip(variants, outs, ok):
            n += 1
            if clean:
                assert k and o == d, "clean stream declined or wrong (round %d)" % it
            elif k:
                try:
                    ref = zlib.decompressobj(wb).decompress(v)
                except zlib.error:
                    ref = None
                assert ref is not None and o == ref, "damaged stream mis-decoded (round %d)" % it
    print("fuzz ok: %d streams" % n)


if __name__ == "__main__":
    main()
