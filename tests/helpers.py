"""Shared test helpers: seeded inputs, zlib reference compression, oracle batch inflate."""
import ctypes
import random
import zlib

import numpy as np

import oracle
from conftest import read_golden


def zcomp(data, level=6, wbits=15, strategy=0):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    return c.compress(data) + c.flush()


def make_data(rng, kind, n, alice):
    if kind == 0:
        o = rng.randrange(0, max(1, len(alice) - n))
        return alice[o:o + n]
    if kind == 1:
        return bytes(rng.getrandbits(8) for _ in range(n))
    if kind == 2:
        return bytes([rng.choice(b"ab")]) * n
    if kind == 3:
        return (b"abcdefgh" * (n // 8 + 1))[:n]
    b = bytearray()
    while len(b) < n:
        if rng.random() < 0.5:
            o = rng.randrange(0, len(alice) - 300)
            b += alice[o:o + rng.randrange(1, 300)]
        else:
            b += bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 200)))
    return bytes(b[:n])


def fuzz_cases(seed, wbits, count, alice, sizes=(0, 1, 2, 5, 100, 1000, 5000, 20000, 70000)):
    """Streams with the edge cases the domain has: empty, truncated, corrupted, small output slot, trailing bytes."""
    rng = random.Random(seed)
    datas, streams, caps = [], [], []
    for _ in range(count):
        kind = rng.randrange(5)
        n = rng.choice(sizes)
        d = make_data(rng, kind, n, alice)
        lvl = rng.choice([0, 1, 3, 6, 9])
        strat = rng.choice([0, 0, 0, 1, 2, 3, 4])
        wb = wbits if wbits != 47 else rng.choice([15, 31])
        s = zcomp(d, lvl, wb, strat)
        mode = rng.randrange(6)
        cap = len(d)
        if mode == 1 and len(s) > 4:
            s = s[:rng.randrange(1, len(s))]
        elif mode == 2 and len(s) > 8:
            b = bytearray(s)
            k = rng.randrange(len(b))
            b[k] ^= 1 << rng.randrange(8)
            s = bytes(b)
        elif mode == 3 and n > 2:
            cap = rng.randrange(0, n)
        elif mode == 4:
            s = s + b"trailing garbage"
        elif mode == 5:
            cap = n + rng.randrange(1, 100)
        datas.append(d)
        streams.append(s)
        caps.append(cap)
    return datas, streams, caps


def pack(chunks):
    offs = np.zeros(len(chunks) + 1, dtype=np.uint64)
    if chunks:
        offs[1:] = np.cumsum([len(c) for c in chunks], dtype=np.uint64)
    buf = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy()
    return buf, offs


def oracle_inflate(streams, caps, wbits, threads=0):
    L = oracle.lib()
    n = len(streams)
    inbuf, in_off = pack(streams)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(np.asarray(caps, dtype=np.uint64))
    out = np.zeros(int(out_off[-1]) + 64, dtype=np.uint8)
    ol = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.oz_inflate_batch(n, p(inbuf), p(in_off), p(out), p(out_off), p(ol), p(st), wbits, threads)
    return [out[int(out_off[i]):int(out_off[i]) + int(ol[i])].tobytes() for i in range(n)], st, ol


def assert_inflate_parity(outs, st, ref_outs, ref_st, tag=""):
    """Bit-exact bytes whenever the stream finished or filled its slot; identical status codes; and every partial
    output is a prefix-compatible piece of the oracle's."""
    for i in range(len(outs)):
        assert st[i] == ref_st[i], "%s stream %d: status %d, oracle %d" % (tag, i, st[i], ref_st[i])
        if st[i] in (1, 2):
            assert outs[i] == ref_outs[i], "%s stream %d: bytes differ" % (tag, i)
        else:
            k = min(len(outs[i]), len(ref_outs[i]))
            assert outs[i][:k] == ref_outs[i][:k], "%s stream %d: partial output is not a prefix" % (tag, i)
