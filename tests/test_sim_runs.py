"""Block-parallel inflate of long streams (inflate_runs.cuh: candidate search, phase A in run mode, 16-bit symbols with
markers, window chain, resolution) with the product's host orchestration (inflate_runs_host.h), on the CPU emulator.
Streams come from zlib (the reference's L0): no flush points, every block type, all three containers. A stream the
parallel path reports as decoded must be bit-exact with all its input accounted for; what it declines (errors, truncation,
small slots) is left to the serial path and must be declined, never mis-decoded."""
import random
import zlib

import pytest

import simlib
from helpers import inspect_deflate, make_data, zcomp


def text(alice, rng, n):
    out = bytearray()
    while len(out) < n:
        o = rng.randrange(0, len(alice) - 2000)
        out += alice[o:o + rng.randrange(200, 2000)]
    return bytes(out[:n])


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_sim_runs_zlib_made_streams(alice, wbits):
    rng = random.Random(31 + wbits)
    datas = [text(alice, rng, 300000), alice + alice[::-1] + alice, make_data(rng, 4, 250000, alice), b"q" * 400000 + text(alice, rng, 50000)]
    streams = [zcomp(d, lvl, wbits) for d, lvl in zip(datas, (6, 9, 1, 6))]
    outs, ok, lens, cons, nruns = simlib.sim_inflate_runs(streams, [len(d) + 7 for d in datas], wbits, chunk_bytes=8192, seed=wbits + 100)
    for i, d in enumerate(datas):
        assert ok[i], "stream %d was declined" % i
        assert outs[i] == d and cons[i] == len(streams[i])
    assert sum(int(x) for x in nruns) > len(datas) + 4, "nothing was cut into runs: %r" % list(nruns)


def test_sim_runs_find_true_block_starts(alice):
    # every run boundary the chain accepts is a real block boundary: the stream decodes bit-exactly from pieces cut there
    rng = random.Random(5)
    d = text(alice, rng, 500000)
    s = zcomp(d, 6, -15)
    blocks, n, end = inspect_deflate(s)
    assert len(blocks) >= 4 and all(b["type"] == 2 for b in blocks)
    outs, ok, lens, cons, nruns = simlib.sim_inflate_runs([s], [len(d)], -15, chunk_bytes=4096, seed=9)
    # every non-final dynamic block starts a run (the final block has BFINAL = 1 and is not searched for)
    assert ok[0] and outs[0] == d and nruns[0] == len(blocks) - 1


def test_sim_runs_mixed_blocks_and_auto_sniff(alice):
    rng = random.Random(8)
    # stored + fixed + dynamic blocks in one stream, sync / full flush points, gzip via auto
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = [text(alice, rng, 90000), bytes(rng.getrandbits(8) for _ in range(80000)), b"ab" * 40000, text(alice, rng, 120000)]
    s = c.compress(parts[0]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(parts[1]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(parts[2]) + \
        c.compress(parts[3]) + c.flush()
    d = b"".join(parts)
    z = zcomp(d, 6, 15)
    for stream, wb in ((s, 47), (s, 31), (z, 47)):
        outs, ok, lens, cons, nruns = simlib.sim_inflate_runs([stream + b"trailing"], [len(d)], wb, chunk_bytes=4096, seed=3)
        assert ok[0] and outs[0] == d and cons[0] == len(stream)


def test_sim_runs_decline_what_they_cannot_prove(alice):
    rng = random.Random(12)
    d = text(alice, rng, 200000)
    s = zcomp(d, 6, 15)
    bad_crc = s[:-1] + bytes([s[-1] ^ 1])
    corrupt = bytearray(s); corrupt[len(s) // 2] ^= 0x10; corrupt = bytes(corrupt)
    truncated = s[:len(s) * 2 // 3]
    cases = [s, bad_crc, corrupt, truncated, s]
    caps = [len(d), len(d), len(d), len(d), len(d) - 1]
    outs, ok, lens, cons, nruns = simlib.sim_inflate_runs(cases, caps, 15, chunk_bytes=8192, seed=21)
    assert ok[0] and outs[0] == d
    assert not ok[1] and not ok[3] and not ok[4]
    assert (not ok[2]) or outs[2] != d  # a flipped bit either breaks the chain / the check ... (it must never pass as the original)
    assert not ok[2]


def test_sim_runs_dense_tokens_are_emitted_again_with_exact_room(alice):
    """A run is decoded once, into a token area bounded by its COMPRESSED size (2 words per byte). Huffman-only output of a
    two-symbol source has ~1 bit per literal = 2.7 token words per compressed byte: those runs outgrow their area, keep
    counting, and are emitted again with exactly the room they need. The result must be the same bytes."""
    rng = random.Random(99)
    d = bytes(98 if rng.random() < 0.03 else 97 for _ in range(400000))
    s = zcomp(d, 6, 15, 2)  # Z_HUFFMAN_ONLY: ~1.03 bits per literal
    assert len(s) < len(d) // 7
    outs, ok, lens, cons, nruns = simlib.sim_inflate_runs([s, zcomp(alice * 2, 6, 15)], [len(d), 2 * len(alice)], 15, chunk_bytes=4096, seed=17)
    assert ok[0] and outs[0] == d and cons[0] == len(s)
    assert ok[1] and outs[1] == alice * 2


@pytest.mark.parametrize("tokw", ["0", "1"])
def test_sim_runs_both_phase_a_kernels(alice, tokw, monkeypatch):
    """Phase A of a run has two kernels — a lane per run (inflate_tok_kernel, also with only every n-th lane taking a run) and a
    warp per run with one decoding lane and look-up tables (inflate_tokw_kernel, launches with few runs): every case above again
    with each of them pinned."""
    monkeypatch.setenv("CUSIM_TOKW", tokw)
    for wbits in (15, 31, -15):
        test_sim_runs_zlib_made_streams(alice, wbits)
    test_sim_runs_find_true_block_starts(alice)
    test_sim_runs_mixed_blocks_and_auto_sniff(alice)
    test_sim_runs_decline_what_they_cannot_prove(alice)
    test_sim_runs_dense_tokens_are_emitted_again_with_exact_room(alice)
