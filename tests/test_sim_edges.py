"""Parity edges the reference's own tests leave open, on the CPU emulator (the SAME kernel sources nvcc compiles):
gzip optional header fields (RFC 1952: FEXTRA / FNAME / FCOMMENT / FHCRC good and bad, FTEXT, a header cut at every byte),
what each ZlibStrategy may emit (/root/reference/src/encoder/zlib_common.rs:5-16), and the compressed size per level against
the oracle's codec at the SAME level. The GPU versions of these tests are in test_gpu_edges.py."""
import zlib

import pytest

import simlib
from helpers import assert_inflate_parity, gzip_member, inspect_deflate, oracle_inflate, zcomp

INFLATE_D = (1, -2)  # warp-per-stream kernel, two-phase (default) path


def header_cases(alice):
    data = alice[:3000]
    raw = zcomp(data, 6, -15)
    cases = [
        gzip_member(raw, data),
        gzip_member(raw, data, fname=b"alice29.txt"),
        gzip_member(raw, data, fcomment=b"a comment, with \xff bytes"),
        gzip_member(raw, data, fextra=b"AB\x04\x00abcd"),
        gzip_member(raw, data, fextra=b""),
        gzip_member(raw, data, fextra=bytes(range(256)) * 3, fname=b"n", fcomment=b"c", ftext=True),
        gzip_member(raw, data, fhcrc=True),
        gzip_member(raw, data, fhcrc=True, bad_hcrc=True),
        gzip_member(raw, data, fextra=b"xy\x01\x00z", fname=b"name", fcomment=b"comment", fhcrc=True),
        gzip_member(raw, data, fextra=b"xy\x01\x00z", fname=b"name", fcomment=b"comment", fhcrc=True, bad_hcrc=True),
        gzip_member(raw, data, fname=b"", fcomment=b""),
    ]
    # reserved flag bits, wrong method, wrong magic
    bad = bytearray(cases[0]); bad[3] |= 0x20; cases.append(bytes(bad))
    bad = bytearray(cases[0]); bad[2] = 7; cases.append(bytes(bad))
    bad = bytearray(cases[0]); bad[1] = 0x8c; cases.append(bytes(bad))
    return data, cases


def test_sim_gzip_optional_header_fields(alice):
    data, cases = header_cases(alice)
    caps = [len(data)] * len(cases)
    for wbits in (31, 47):
        ref_outs, ref_st, _ = oracle_inflate(cases, caps, wbits)
        assert list(ref_st[:7]) == [2] * 7 and ref_st[7] == -3 and ref_st[8] == 2 and ref_st[9] == -3
        for D in INFLATE_D:
            outs, st, ol, cons, _ = simlib.sim_inflate(cases, caps, wbits, D=D, seed=5)
            assert_inflate_parity(outs, st, ref_outs, ref_st, "wbits %d D %d" % (wbits, D))


def test_sim_gzip_header_truncated_at_every_byte(alice):
    data = alice[:600]
    raw = zcomp(data, 6, -15)
    full = gzip_member(raw, data, fextra=b"xy\x03\x00abc", fname=b"name.txt", fcomment=b"comment", fhcrc=True)
    hdr_len = len(full) - len(raw) - 8
    streams = [full[:k] for k in range(0, hdr_len + 6)]
    # the same cuts with a wrong header CRC: the verdict must come exactly when the second CRC byte arrives
    badm = gzip_member(raw, data, fextra=b"xy\x03\x00abc", fname=b"name.txt", fcomment=b"comment", fhcrc=True, bad_hcrc=True)
    streams += [badm[:k] for k in range(hdr_len - 3, hdr_len + 3)]
    caps = [len(data)] * len(streams)
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, 31)
    for D in INFLATE_D:
        outs, st, ol, cons, _ = simlib.sim_inflate(streams, caps, 31, D=D, seed=9)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "D %d" % D)


def strategy_inputs(alice):
    return [alice[:60000], b"ab" * 9000 + alice[:2000] + b"\0" * 5000, bytes((i * 7 + (i >> 5)) & 0xff for i in range(20000))]


def check_strategy_semantics(streams, datas, strategy):
    """What the strategy allows in the emitted blocks (zlib's deflate.h / deflate.c semantics, which compu passes through:
    /root/reference/src/encoder/zlib_ng.rs:70-76)."""
    for k, (s, d) in enumerate(zip(streams, datas)):
        assert zlib.decompress(s, -15) == d
        blocks, n, end = inspect_deflate(s)
        assert n == len(d) and end == len(s)
        for b in blocks:
            if strategy == 4:   # Fixed: "prevents the use of dynamic Huffman codes"
                assert b["type"] != 2
            if strategy == 2:   # HuffmanOnly: "force Huffman encoding only (no string match)"
                assert b["matches"] == 0
            if strategy == 3:   # Rle: "limit match distances to one (run-length encoding)"
                assert b["dists"] <= {1}
        if strategy in (2, 3, 0, 1) and k == 0:
            assert any(b["type"] == 2 for b in blocks)  # these strategies still use dynamic codes on text
        if strategy == 3:
            assert sum(b["matches"] for b in blocks) > 0 or b"\0\0\0\0" not in d  # runs ARE matched


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
def test_sim_strategy_semantics(alice, strategy):
    datas = strategy_inputs(alice)
    streams, st, lens, checks, _ = simlib.sim_deflate(datas, seg_bytes=65536, level=6, strategy=strategy, window_bits=-15)
    assert list(st) == [2] * len(datas)
    check_strategy_semantics(streams, datas, strategy)


# Size relative to zlib 1.3 at the SAME level, one 64 KiB segment of alice29.txt per stream. Measured on this build (the
# emulator and the GPU agree byte for byte): L1 0.940, L2 0.956, L3 0.983, L4 0.981, L5 1.001, L6 1.008, L7 1.005, L8 1.000,
# L9 0.997 — the low levels search more than zlib's deflate_fast does. The bound is the north star's 3 % at every level.
LEVEL_BOUND = {1: 1.03, 3: 1.03, 6: 1.03, 9: 1.03}


@pytest.mark.parametrize("level", [1, 3, 6, 9])
def test_sim_size_per_level_against_oracle_level(alice, level):
    datas = [alice[i * 65536:(i + 1) * 65536] for i in range(2)]
    streams, st, lens, checks, _ = simlib.sim_deflate(datas, seg_bytes=65536, level=level, window_bits=-15)
    assert list(st) == [2, 2]
    ours = sum(len(s) for s in streams)
    ref = sum(len(zcomp(d, level, -15)) for d in datas)
    assert [zlib.decompress(s, -15) for s in streams] == datas
    assert ours <= ref * LEVEL_BOUND[level], "level %d: %d bytes vs zlib's %d (%.2f %%)" % (level, ours, ref, 100.0 * ours / ref - 100)
    # monotone where it must be: a higher level never compresses worse than level 1 by more than noise
    if level > 1:
        l1 = sum(len(s) for s in simlib.sim_deflate(datas, seg_bytes=65536, level=1, window_bits=-15)[0])
        assert ours <= l1
