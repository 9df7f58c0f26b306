"""Kernel-logic tests on the CPU emulator (tests/cusim): the SAME inflate_kernel.cuh source nvcc compiles, run one fiber
per CUDA thread, compared with the oracle. CPU only; the real parity tests are test_gpu_*.py."""
import pytest

import simlib
from helpers import assert_inflate_parity, fuzz_cases, oracle_inflate, zcomp


@pytest.mark.parametrize("D", [1, 4, 32, -9, -8, -1, -2, -3, -4, -5, -6])
def test_sim_golden_gzip(golden, D):
    streams = [c for _, c in golden]
    caps = [len(d) for d, _ in golden]
    outs, st, ol, cons, _ = simlib.sim_inflate(streams, caps, 31, D=D)
    assert list(st) == [2, 2]
    assert outs[0] == golden[0][0] and outs[1] == golden[1][0]
    assert list(cons) == [len(s) for s in streams]


@pytest.mark.parametrize("wbits", [15, 31, -15, 47])
def test_sim_fuzz_vs_oracle(alice, wbits):
    datas, streams, caps = fuzz_cases(100 + wbits, wbits, 40, alice, sizes=(0, 1, 2, 5, 100, 1000, 5000, 20000))
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
    for D in (1, 8, -9, -1, -2, -3, -4, -5, -6):
        outs, st, ol, cons, _ = simlib.sim_inflate(streams, caps, wbits, D=D, seed=wbits + D)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "wbits %d D %d" % (wbits, D))


def test_sim_segment_mode_checks(alice):
    import zlib
    # a full-flush segment: raw deflate, no BFINAL, ends with 00 00 ff ff at a byte boundary
    segs, datas = [], []
    for i in range(6):
        d = alice[i * 20000:(i + 1) * 20000 + 37 * i]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        segs.append(c.compress(d) + c.flush(zlib.Z_FULL_FLUSH))
        datas.append(d)
    for D in (4, -9, -1, -2, -3, -4, -5, -6):
        outs, st, ol, cons, ck = simlib.sim_inflate(segs, [len(d) for d in datas], -15, segment_mode=1, check_kind=3, D=D)
        assert list(st) == [2] * 6
        for i, d in enumerate(datas):
            assert outs[i] == d
            assert ck[2 * i] == zlib.adler32(d) and ck[2 * i + 1] == zlib.crc32(d)


def test_sim_cta_tile_cases(alice):
    """Units around the 64 KiB shared-memory tile of inflate_lz_cta_kernel: exact fits, odd slot sizes (misaligned
    output offsets), stored runs, long runs (clipped chunks), one unit beyond the tile (handled by inflate_lz_kernel)."""
    import random
    from helpers import make_data
    rng = random.Random(4242)
    datas, streams, caps, wbs = [], [], [], []
    for kind, n, lvl, wb in [(0, 65536, 6, 15), (1, 65536, 6, 15), (2, 65536, 6, 31), (3, 65535, 9, 31), (4, 60001, 6, 15),
                             (0, 65536, 1, -15), (4, 65536, 0, 15), (0, 70000, 6, 15), (1, 33333, 6, 31), (0, 3, 6, 15)]:
        d = make_data(rng, kind, n, alice)
        datas.append(d)
        streams.append(zcomp(d, lvl, wb))
        caps.append(len(d))
        wbs.append(wb)
    for wbits in (15, 31, -15):
        idx = [i for i in range(len(datas)) if wbs[i] == wbits]
        ss = [streams[i] for i in idx]
        cc = [caps[i] for i in idx]
        ref_outs, ref_st, _ = oracle_inflate(ss, cc, wbits)
        outs, st, ol, cons, _ = simlib.sim_inflate(ss, cc, wbits, D=-4, seed=7 + wbits)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "cta wbits %d" % wbits)
        assert list(st) == [2] * len(idx)


def test_sim_truncated_tail_does_not_see_the_neighbour(alice):
    """Bits past the end of a unit must read as zero whatever follows it in the packed batch: a truncated stream decides
    NEED_INPUT vs data error from its own bytes only (found by a wider fuzz: a 24-byte truncated gzip stream followed by
    another unit reported -3 instead of 0)."""
    base = [zcomp(b"a" * 5000, 6, 31), zcomp(alice[:3000], 6, 31), zcomp(bytes(range(256)) * 8, 9, 31)]
    streams, caps = [], []
    for s, cap in zip(base, (5000, 3000, 2048)):
        for cut in range(11, len(s) - 1, max(1, len(s) // 23)):
            streams.append(s[:cut]); caps.append(cap)
            streams.append(b"\xff" * 7); caps.append(16)     # a neighbour full of one bits
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, 31)
    for D in (1, 8, -1, -2, -3):
        outs, st, ol, cons, _ = simlib.sim_inflate(streams, caps, 31, D=D, seed=3)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "D %d" % D)
