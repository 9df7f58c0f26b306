"""Encoder kernel-logic tests on the CPU emulator (tests/cusim): the SAME deflate_kernels.cuh source nvcc compiles, checked
(1) for decodability and bit-exact round trip with zlib (the reference decoder's L0), (2) byte for byte against a sequential
host run of the same decisions (tests/model), (3) for ratio against zlib level 6. CPU only."""
import ctypes
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

import simlib
from helpers import make_data

_HERE = os.path.dirname(os.path.abspath(__file__))


def model_lib():
    d = os.path.join(_HERE, "model")
    subprocess.check_call(["make", "-C", d, "-s"])
    L = ctypes.CDLL(os.path.join(d, "libdeflate_model.so"))
    L.model_deflate_segment.restype = ctypes.c_long
    return L


def model_segment(L, data, level=6, strategy=0):
    src = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    out = np.zeros(len(data) + len(data) // 8 + 1024, dtype=np.uint8)
    r = L.model_deflate_segment(src.ctypes.data_as(ctypes.c_void_p), len(data), out.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_uint64(len(out)), level, strategy, None, None, None)
    assert r >= 0
    return out[:r].tobytes()


def dec(stream, wbits):
    d = zlib.decompressobj(wbits)
    out = d.decompress(stream)
    assert d.eof and d.unused_data == b""
    return out


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_sim_deflate_roundtrip_containers(alice, wbits):
    units = [alice[:70000], b"", b"a", alice[1000:1100], b"X" * 10 + b"Y" * 10, bytes(range(256)) * 40]
    streams, st, lens, checks, _ = simlib.sim_deflate(units, seg_bytes=32768, window_bits=wbits)
    assert list(st) == [2] * len(units)
    for u, s, i in zip(units, streams, range(len(units))):
        assert dec(s, wbits) == u
        assert checks[2 * i] == zlib.adler32(u) and checks[2 * i + 1] == zlib.crc32(u)
    if wbits == 15:
        assert streams[0][:2] == b"\x78\x9c"
    if wbits == 31:
        assert streams[0][:10] == bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3])


@pytest.mark.parametrize("seed", [1, 2, 3, 5, 10, 20])  # match kernel: 1 tiled, 2 thread-per-position, 3 candidate pairs, 5 sweep (default), 10 sweep over single links... (seed & 4: no second-link array)
def test_sim_deflate_matches_sequential_model(alice, seed):
    L = model_lib()
    rng = random.Random(11)
    # the 230 000-byte unit crosses the 64 KiB wrap of the 16-bit head table and several sweeps
    units = [make_data(rng, k, n, alice) for k, n in [(0, 40000), (1, 3000), (2, 5000), (3, 9000), (4, 20000), (0, 17)]]
    units.append((alice + alice[:80000])[:230000])
    streams, st, lens, _, seg_sizes = simlib.sim_deflate(units, seg_bytes=1 << 20, window_bits=-15, piece_mode=1, seed=seed)
    assert list(st) == [2] * len(units)
    for u, s in zip(units, streams):
        assert s == model_segment(L, u), "kernel chain differs from the sequential run of the same decisions"
    assert [int(x) for x in seg_sizes] == [len(s) for s in streams]


@pytest.mark.parametrize("level,strategy", [(0, 0), (1, 0), (9, 0), (6, 1), (6, 2), (6, 3), (6, 4)])
def test_sim_deflate_levels_and_strategies(alice, level, strategy):
    rng = random.Random(level * 10 + strategy)
    units = [alice[:30000], make_data(rng, 1, 2000, alice), make_data(rng, 2, 3000, alice), make_data(rng, 4, 8000, alice)]
    streams, st, _, _, _ = simlib.sim_deflate(units, seg_bytes=16384, level=level, strategy=strategy, window_bits=15)
    assert list(st) == [2] * len(units)
    for u, s in zip(units, streams):
        assert dec(s, 15) == u
    # every match kernel makes the same decisions: identical bytes
    for seed in (2, 3, 5, 20):
        streams2, st2, _, _, _ = simlib.sim_deflate(units, seg_bytes=16384, level=level, strategy=strategy, window_bits=15, seed=seed)
        assert streams2 == streams and list(st2) == list(st)


def test_sim_deflate_packed_and_capacity(alice):
    units = [alice[i * 9000:(i + 1) * 9000] for i in range(5)]
    caps = [20000, 100, 20000, 20000, 20000]  # unit 1 cannot fit
    streams, st, lens, _, _ = simlib.sim_deflate(units, seg_bytes=4096, window_bits=-15, piece_mode=1, packed=1, caps=caps)
    assert list(st) == [2, 1, 2, 2, 2]
    assert lens[1] > 100
    for i in (0, 2, 3, 4):
        d = zlib.decompressobj(-15)
        assert d.decompress(streams[i]) == units[i]  # pieces never set BFINAL: no eof, but every byte is produced


def test_sim_deflate_ratio_vs_zlib6(alice):
    # whole file as one segment: within 3 % of zlib level 6 (north star); the survey's known answer is 54 398 B raw
    streams, st, _, _, _ = simlib.sim_deflate([alice], seg_bytes=1 << 20, window_bits=-15)
    z = zlib.compressobj(6, zlib.DEFLATED, -15)
    ref = len(z.compress(alice) + z.flush())
    assert ref == 54398
    assert dec(streams[0], -15) == alice
    assert len(streams[0]) <= ref * 1.03, (len(streams[0]), ref)


def test_sim_segment_join_is_one_valid_stream(alice):
    # header | seg0 | seg1 | ... | 03 00 | trailer decodes with the reference decoder's L0 in one go, and each segment
    # also inflates standalone (the basis of segment-parallel decode)
    streams, st, _, _, seg_sizes = simlib.sim_deflate([alice[:100000]], seg_bytes=16384, window_bits=31)
    s = streams[0]
    assert dec(s, 31) == alice[:100000]
    o = 10
    for i, sz in enumerate(seg_sizes):
        piece = s[o:o + int(sz)]
        assert piece[-4:] == b"\x00\x00\xff\xff"
        assert zlib.decompressobj(-15).decompress(piece) == alice[i * 16384:min((i + 1) * 16384, 100000)]
        o += int(sz)
    assert s[o:o + 2] == b"\x03\x00"


def test_length_limited_huffman_is_complete_and_within_limit():
    """Kraft sum of the code lengths must be exactly 1 (zlib rejects anything else: "invalid code lengths set" /
    "invalid literal/lengths set") and no length may exceed the limit — including skewed distributions that force the
    overflow repair (Fibonacci-like frequencies make the unrestricted tree as deep as the alphabet)."""
    L = model_lib()
    rng = np.random.default_rng(5)
    cases = []
    for n, maxbits in ((19, 7), (30, 15), (286, 15)):
        fib = [1, 1]
        while len(fib) < n:
            fib.append(min(fib[-1] + fib[-2], 1 << 22))
        cases.append((n, maxbits, np.array(fib[:n], dtype=np.uint32)))
        cases.append((n, maxbits, np.array(fib[:n][::-1], dtype=np.uint32)))
        for _ in range(300):
            k = int(rng.integers(1, n + 1))
            f = np.zeros(n, dtype=np.uint32)
            idx = rng.choice(n, k, replace=False)
            shape = rng.choice(["flat", "geo", "zipf", "rand"])
            if shape == "flat":
                f[idx] = rng.integers(1, 4, k)
            elif shape == "geo":
                f[idx] = np.minimum(1.7 ** np.minimum(np.arange(k), 40), float(1 << 22)).astype(np.uint32)
            elif shape == "zipf":
                f[idx] = np.maximum(1, (100000 / (1 + np.arange(k)) ** 2).astype(np.uint32))
            else:
                f[idx] = rng.integers(1, 1 << 16, k)
            cases.append((n, maxbits, f))
    for n, maxbits, f in cases:
        lens = np.zeros(n, dtype=np.uint8)
        L.model_huff_lengths(f.ctypes.data_as(ctypes.c_void_p), n, maxbits, lens.ctypes.data_as(ctypes.c_void_p))
        used = lens[lens > 0]
        assert used.max() <= maxbits
        assert (lens[f > 0] > 0).all()
        kraft = sum(1 << (maxbits - int(l)) for l in used)
        assert kraft == 1 << maxbits, (n, maxbits, kraft, f.tolist())


def test_sim_deflate_near_random_1mib_segment():
    # regression: near-random data drives the code-length code past 7 bits (the repair path of the length-limited Huffman)
    rng = np.random.default_rng(698)
    parts = []
    for i in range(64):
        parts.append(rng.integers(0, 256, 4096, dtype=np.uint8).tobytes() if i % 10 else bytes(rng.integers(97, 123, 4096, dtype=np.uint8)))
    data = b"".join(parts) * 2
    streams, st, _, _, _ = simlib.sim_deflate([data], seg_bytes=1 << 20, window_bits=-15, piece_mode=1)
    assert st[0] == 2
    assert zlib.decompressobj(-15).decompress(streams[0]) == data
