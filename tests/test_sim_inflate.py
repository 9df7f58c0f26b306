"""Kernel-logic tests on the CPU emulator (tests/cusim): the SAME inflate_kernel.cuh source nvcc compiles, run one fiber
per CUDA thread, compared with the oracle. CPU only; the real parity tests are test_gpu_*.py."""
import pytest

import simlib
from helpers import assert_inflate_parity, fuzz_cases, oracle_inflate, zcomp


@pytest.mark.parametrize("D", [1, 4, 32, -9, -8, -1, -2, -3])
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
    for D in (1, 8, -9, -1, -2, -3):
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
    for D in (4, -9, -1, -2, -3):
        outs, st, ol, cons, ck = simlib.sim_inflate(segs, [len(d) for d in datas], -15, segment_mode=1, check_kind=3, D=D)
        assert list(st) == [2] * 6
        for i, d in enumerate(datas):
            assert outs[i] == d
            assert ck[2 * i] == zlib.adler32(d) and ck[2 * i + 1] == zlib.crc32(d)
