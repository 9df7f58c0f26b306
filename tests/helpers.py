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


# ---------------------------------------------------------------------------------------------------------------------
# A small RFC 1951 reader used to LOOK INSIDE encoder output (block types, literal / match mix, distances) — the zlib module
# only says whether a stream decodes. Pure Python, for inputs up to a few hundred KB.
_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289,
              16385, 24577]
_DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


class _Bits:
    def __init__(self, data):
        self.v = int.from_bytes(data, "little")
        self.pos = 0
        self.nbits = 8 * len(data)

    def get(self, n):
        assert self.pos + n <= self.nbits, "inspector ran past the end of the stream"
        r = (self.v >> self.pos) & ((1 << n) - 1)
        self.pos += n
        return r


def _canon(lengths):
    """code lengths -> {(length, code): symbol}"""
    table, code = {}, 0
    for ln in range(1, 16):
        for sym, l in enumerate(lengths):
            if l == ln:
                table[(ln, code)] = sym
                code += 1
        code <<= 1
    return table


def _decode_sym(b, table):
    code = 0
    for ln in range(1, 16):
        code = (code << 1) | b.get(1)
        s = table.get((ln, code))
        if s is not None:
            return s
    raise AssertionError("inspector: invalid Huffman code")


def inspect_deflate(raw):
    """Walks a raw-deflate stream (no container). Returns (blocks, out_len, end_byte) where every block is a dict
    {type: 0|1|2, final, literals, matches, dists: set, min_len, max_len, stored_len}."""
    b = _Bits(raw)
    blocks, out_len = [], 0
    while True:
        final, typ = b.get(1), b.get(2)
        blk = {"type": typ, "final": final, "literals": 0, "matches": 0, "dists": set(), "min_len": None, "max_len": None, "stored_len": 0}
        if typ == 0:
            b.pos = (b.pos + 7) & ~7
            ln, nln = b.get(16), b.get(16)
            assert ln ^ 0xffff == nln
            b.pos += 8 * ln
            blk["stored_len"] = ln
            out_len += ln
        else:
            assert typ in (1, 2), "reserved block type"
            if typ == 1:
                lt = _canon([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8)
                dt = _canon([5] * 32)
            else:
                hlit, hdist, hclen = b.get(5) + 257, b.get(5) + 1, b.get(4) + 4
                order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
                cl = [0] * 19
                for i in range(hclen):
                    cl[order[i]] = b.get(3)
                ct = _canon(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = _decode_sym(b, ct)
                    if s < 16:
                        lens.append(s)
                    elif s == 16:
                        lens += [lens[-1]] * (3 + b.get(2))
                    elif s == 17:
                        lens += [0] * (3 + b.get(3))
                    else:
                        lens += [0] * (11 + b.get(7))
                lt, dt = _canon(lens[:hlit]), _canon(lens[hlit:hlit + hdist])
            while True:
                s = _decode_sym(b, lt)
                if s < 256:
                    blk["literals"] += 1
                    out_len += 1
                elif s == 256:
                    break
                else:
                    ln = _LEN_BASE[s - 257] + b.get(_LEN_EXTRA[s - 257])
                    ds = _decode_sym(b, dt)
                    dist = _DIST_BASE[ds] + b.get(_DIST_EXTRA[ds])
                    blk["matches"] += 1
                    blk["dists"].add(dist)
                    blk["min_len"] = ln if blk["min_len"] is None else min(blk["min_len"], ln)
                    blk["max_len"] = ln if blk["max_len"] is None else max(blk["max_len"], ln)
                    out_len += ln
        blocks.append(blk)
        if final:
            break
    return blocks, out_len, (b.pos + 7) // 8


def gzip_member(payload_raw, data, fextra=None, fname=None, fcomment=None, fhcrc=False, bad_hcrc=False, ftext=False):
    """Hand-built RFC 1952 member around a raw-deflate payload, with any combination of optional header fields."""
    flg = (1 if ftext else 0) | (2 if fhcrc else 0) | (4 if fextra is not None else 0) | (8 if fname is not None else 0) | \
          (16 if fcomment is not None else 0)
    h = bytes([0x1f, 0x8b, 8, flg, 0x12, 0x34, 0x56, 0x78, 0, 3])
    if fextra is not None:
        h += len(fextra).to_bytes(2, "little") + fextra
    if fname is not None:
        h += fname + b"\0"
    if fcomment is not None:
        h += fcomment + b"\0"
    if fhcrc:
        c = zlib.crc32(h) & 0xffff
        if bad_hcrc:
            c ^= 0x0101
        h += c.to_bytes(2, "little")
    return h + payload_raw + (zlib.crc32(data) & 0xffffffff).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")
