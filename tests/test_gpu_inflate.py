"""GPU parity tests for inflate: the CUDA path through the C ABI vs the CPU oracle, on the reference's golden vectors,
seeded synthetic inputs, and the domain's edge cases. Run with -m gpu on a B200."""
import zlib

import numpy as np
import pytest

import protocols
from compu_b200 import _lib, batch
from compu_b200.decoder import DecodeError, DecodeStatus, Interface, ZlibMode
from helpers import assert_inflate_parity, fuzz_cases, oracle_inflate, zcomp

pytestmark = pytest.mark.gpu


def test_device_present():
    assert _lib.require_device() >= 1


@pytest.mark.parametrize("mode", [ZlibMode.Gzip, ZlibMode.Auto])
def test_reference_decoder_protocol_on_golden_gzip(golden, mode):
    # should_decode_zlib_ng_gzip (tests/decoder.rs:141-150) with the CUDA backend in place of zlib_ng
    d = Interface.zlib_cuda(mode)
    assert d is not None, _lib.last_error()
    for data, comp in golden:
        protocols.decoder_test_case(d, data, comp)


def test_batch_golden_all_containers(golden):
    for data, comp in golden:
        outs, st, lens, cons = batch.inflate_batch([comp], [len(data)], 31)
        assert st[0] == 2 and outs[0] == data and cons[0] == len(comp)
        for wb in (15, -15, 31):
            for lvl in (1, 6, 9):
                s = zcomp(data, lvl, wb)
                outs, st, _, cons = batch.inflate_batch([s, s + b"xyz"], [len(data), len(data)], wb)
                assert list(st) == [2, 2] and outs[0] == data and outs[1] == data
                assert list(cons) == [len(s), len(s)]


@pytest.mark.parametrize("wbits", [15, 31, -15, 47])
def test_fuzz_vs_oracle(alice, wbits):
    datas, streams, caps = fuzz_cases(7000 + wbits, wbits, 400, alice)
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
    outs, st, lens, cons = batch.inflate_batch(streams, caps, wbits)
    assert_inflate_parity(outs, st, ref_outs, ref_st, "wbits %d" % wbits)


def test_many_64k_streams_vs_oracle(alice):
    # cfg2 in miniature: independent 64 KiB zlib streams (text), level 6
    rng = np.random.default_rng(5)
    big = (alice * 8)
    chunks = []
    for i in range(512):
        o = int(rng.integers(0, len(big) - 65536))
        chunks.append(big[o:o + 65536])
    streams = [zlib.compress(c, 6) for c in chunks]
    outs, st, lens, cons = batch.inflate_batch(streams, [65536] * len(chunks), 15)
    ref_outs, ref_st, _ = oracle_inflate(streams, [65536] * len(chunks), 15)
    assert_inflate_parity(outs, st, ref_outs, ref_st)
    assert (st == 2).all()


def test_ragged_and_large_streams(alice):
    # stream sizes 4 KiB .. 16 MiB (cfg5 shape), mixed classes
    rng = np.random.default_rng(9)
    datas = []
    for sz in (4096, 5000, 70001, 1 << 20, (1 << 22) + 17, 1 << 24):
        rep = (alice * (sz // len(alice) + 2))
        a = np.frombuffer(rep[:sz], dtype=np.uint8).copy()
        noise = rng.integers(0, 256, sz, dtype=np.uint8)
        mask = rng.random(sz) < 0.02
        a[mask] = noise[mask]
        datas.append(a.tobytes())
    datas.append(rng.integers(0, 256, 300000, dtype=np.uint8).tobytes())  # incompressible -> stored blocks
    datas.append(b"")
    streams = [zlib.compress(d, 6) for d in datas]
    outs, st, lens, cons = batch.inflate_batch(streams, [len(d) for d in datas], 15)
    assert (st == 2).all(), st
    for o, d in zip(outs, datas):
        assert o == d


def test_error_codes_match_oracle(golden):
    data, comp = golden[1]
    d = Interface.zlib_cuda(ZlibMode.Gzip)
    out = bytearray(len(data))
    r = d.decode(comp[:len(comp) // 2], out)
    assert r.status == DecodeStatus.NeedInput and r.input_remain == 0 and r.output_remain > 0
    assert bytes(out[:len(out) - r.output_remain]) == data[:len(out) - r.output_remain]
    r = d.decode(b"", out)  # zero-progress call
    assert r.status == DecodeStatus.NeedOutput
    d.reset()
    bad = bytearray(comp)
    bad[len(bad) // 2] ^= 0x40
    r = d.decode(bytes(bad), out)
    assert r.status == DecodeError(-3)
    assert d.describe_error(r.status) == "data error"
    d.reset()
    bad = bytearray(comp)
    bad[-6] ^= 1
    r = d.decode(bytes(bad), out)
    assert r.status == DecodeError(-3)
    d.reset()
    # trailing bytes after the stream end are handed back as input_remain
    r = d.decode(comp + b"TRAIL", out)
    assert r.status == DecodeStatus.Finished and r.input_remain == 5 and bytes(out) == data
    # FDICT -> need dictionary (zlib code 2)
    dz = Interface.zlib_cuda(ZlibMode.Zlib)
    c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, 0, b"some dictionary")
    s = c.compress(b"hello hello") + c.flush()
    r = dz.decode(s, out)
    assert r.status == DecodeError(2)


def test_inflate_config_sweep_is_consistent(alice):
    # every instantiated (slots-per-warp, warps) configuration must give identical bytes (the variants beyond the defaults
    # only exist in a build with -DCZ_EXPERIMENTS)
    import os
    import subprocess
    import sys
    code = ("import sys,zlib;sys.path.insert(0,'.');from compu_b200 import batch;"
            "a=open('tests/golden/alice29.txt','rb').read();c=[a[i:i+30000] for i in range(0,len(a),30000)];"
            "s=[zlib.compress(x,6) for x in c];o,st,_,_=batch.inflate_batch(s,[len(x) for x in c],15);"
            "assert (st==2).all() and o==c;print('ok')")
    cfgs = ("1,8", "2,8", "4,7", "8,7", "16,3", "32,1") if _lib.lib().cz_has_experiments() else ("1,8", "-2,14")
    for cfg in cfgs:
        env = dict(os.environ, CZ_INFLATE_CFG=cfg)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0 and "ok" in r.stdout, (cfg, r.stdout, r.stderr)


def test_mixed_big_and_small_units_vs_oracle(alice):
    # ragged batch (cfg5 in miniature): megabyte streams go to the warp-per-stream kernel, small ones to the two-phase path
    rng = np.random.default_rng(11)
    big = alice * 20
    datas = []
    for i in range(60):
        n = int(rng.choice([4096, 20000, 65536, 300000, 1 << 21, 3_000_000])) if i % 7 else int(rng.integers(0, 2000))
        o = int(rng.integers(0, len(big) - n))
        datas.append(big[o:o + n])
    for wbits in (15, 31):
        streams = [zcomp(d, 6, wbits) for d in datas]
        caps = [len(d) for d in datas]
        outs, st, lens, cons = batch.inflate_batch(streams, caps, wbits)
        ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "mixed wbits %d" % wbits)
        assert list(cons) == [len(s) for s in streams]


def test_pinned_buffer_and_pointer_array_entry(golden, alice):
    # PinnedBuffer: Buffer<N> over page-locked memory (the pinned analogue of compu_malloc / Buffer<N>)
    from compu_b200 import PinnedBuffer
    d = Interface.zlib_cuda(ZlibMode.Gzip)
    data, comp = golden[1]
    buf = PinnedBuffer(8192)
    out = bytearray()
    inp = comp
    while True:
        consumed, status = buf.decode(d, inp)
        inp = inp[consumed:]
        out += buf.data()
        buf.consume()
        if status == DecodeStatus.Finished:
            break
    assert bytes(out) == data
    buf.close()
    # cz_inflate_batch_ptrs: the pointer-array form of the batched entry point (SURVEY.md §8b proposal)
    import ctypes
    L = _lib.lib()
    chunks = [alice[i * 10000:(i + 1) * 10000] for i in range(12)] + [b""]
    streams = [zlib.compress(c, 6) for c in chunks]
    n = len(streams)
    ins = [np.frombuffer(s, dtype=np.uint8).copy() for s in streams]
    outs = [np.zeros(max(1, len(c)), dtype=np.uint8) for c in chunks]
    in_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in ins])
    out_ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in outs])
    in_lens = (ctypes.c_size_t * n)(*[len(s) for s in streams])
    out_caps = (ctypes.c_size_t * n)(*[len(c) for c in chunks])
    out_lens = (ctypes.c_size_t * n)()
    st = (ctypes.c_int32 * n)()
    rc = L.cz_inflate_batch_ptrs(n, in_ptrs, in_lens, out_ptrs, out_caps, out_lens, st, 15, 0)
    _lib.check(rc, "cz_inflate_batch_ptrs")
    assert list(st) == [2] * n and list(out_lens) == [len(c) for c in chunks]
    for a, c in zip(outs, chunks):
        assert a[:len(c)].tobytes() == c


def test_cta_tile_cases_vs_oracle(alice):
    """Units around the 64 KiB shared-memory tile of inflate_lz_cta_kernel: exact fits, odd slot sizes (misaligned output
    offsets), stored runs, long runs (clipped chunks), units beyond the tile (inflate_lz_kernel), all three containers."""
    import random
    from helpers import make_data
    rng = random.Random(4242)
    cases = [(0, 65536, 6), (1, 65536, 6), (2, 65536, 6), (3, 65535, 9), (4, 60001, 6), (0, 65536, 1), (4, 65536, 0),
             (0, 70000, 6), (1, 33333, 6), (0, 3, 6), (0, 0, 6), (2, 65536, 9), (3, 16385, 6), (4, 65521, 3)] * 8
    L = _lib.lib()
    try:
        for wbits in (15, 31, -15):
            datas = [make_data(rng, kind, n, alice) for kind, n, _ in cases]
            streams = [zcomp(d, lvl, wbits) for d, (_, _, lvl) in zip(datas, cases)]
            caps = [len(d) for d in datas]
            ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
            for mode in ((0, 1, 2) if L.cz_has_experiments() else (0,)):
                assert L.cz_tune_inflate_lz(mode, 100) == 0
                outs, st, lens, cons = batch.inflate_batch(streams, caps, wbits)
                assert_inflate_parity(outs, st, ref_outs, ref_st, "cta mode %d wbits %d" % (mode, wbits))
                assert (st == 2).all()
                assert list(cons) == [len(s) for s in streams]
    finally:
        L.cz_tune_inflate_lz(0, 100)


def test_truncated_tail_does_not_see_the_neighbour(alice):
    """Bits past the end of a unit read as zero whatever follows it in the packed batch (a truncated stream decides NEED_INPUT
    vs data error from its own bytes only) — every kernel family."""
    base = [zcomp(b"a" * 5000, 6, 31), zcomp(alice[:3000], 6, 31), zcomp(bytes(range(256)) * 8, 9, 31), zcomp(alice[:70000], 6, 31)]
    streams, caps = [], []
    for s, cap in zip(base, (5000, 3000, 2048, 70000)):
        for cut in range(11, len(s) - 1, max(1, len(s) // 37)):
            streams.append(s[:cut]); caps.append(cap)
            streams.append(b"\xff" * 7); caps.append(16)
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, 31)
    L = _lib.lib()
    try:
        for cfg in ([(-2, 14), (-1, 14), (1, 8), (4, 7), (-9, 8)] if L.cz_has_experiments() else [(-2, 14), (1, 8)]):
            assert L.cz_tune_inflate(*cfg) == 0
            outs, st, lens, cons = batch.inflate_batch(streams, caps, 31)
            assert_inflate_parity(outs, st, ref_outs, ref_st, "cfg %s" % (cfg,))
    finally:
        L.cz_tune_inflate(-2, 14)


def test_streaming_decoder_slot_growth_byte_by_byte():
    """The streaming decoder re-inflates what is staged into a slot of its own choosing (64 KiB at first) and must grow it
    whenever it fills up — also when the kernels report NeedInput for a full slot (the symbol that did not fit ends in the last
    staged byte, zlib's avail_in == 0 case). Feeding a highly compressible 200 000-byte stream one byte at a time walks through
    every staged-prefix length, including that one."""
    data = (b"a" * 70000) + bytes(range(256)) * 100 + (b"xyz" * 35000)
    for wb, mode in ((15, ZlibMode.Zlib), (-15, ZlibMode.Deflate), (31, ZlibMode.Gzip)):
        stream = zcomp(data, 9, wb)
        dec = Interface.zlib_cuda(mode)
        assert dec is not None
        out = bytearray()
        window = bytearray(1 << 16)
        status = None
        for i in range(len(stream)):
            piece = stream[i:i + 1]
            while True:
                r = dec.decode(piece, window)
                out += window[:len(window) - r.output_remain]
                piece = piece[len(piece) - r.input_remain:]
                status = r.status
                if status != DecodeStatus.NeedOutput:
                    break
            assert not isinstance(status, DecodeError), (wb, i, status)
        for _ in range(8):  # drain (zlib-style callers loop until Finished)
            if status == DecodeStatus.Finished:
                break
            r = dec.decode(b"", window)
            out += window[:len(window) - r.output_remain]
            status = r.status
        assert status == DecodeStatus.Finished and bytes(out) == data, (wb, status, len(out))
        dec.close()
