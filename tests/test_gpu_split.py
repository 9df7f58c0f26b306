"""Long single streams through the ORDINARY entry points (cz_inflate_batch, the streaming Decoder): they are cut into runs of
blocks at block headers found by a candidate search and decoded in parallel (inflate_runs.cuh) — whoever produced them, with
or without flush points; what that path cannot prove (errors, truncation, small slots) falls back to the serial path. Results
must be what the oracle (compu's glue over zlib) gives in every case. Run with -m gpu on a B200."""
import zlib

import numpy as np
import pytest

from compu_b200 import Vec, _lib, batch
from compu_b200 import decoder as dec
from helpers import oracle_inflate

pytestmark = pytest.mark.gpu


def _text(alice, n, seed=1):
    rng = np.random.default_rng(seed)
    parts = []
    total = 0
    while total < n:
        o = int(rng.integers(0, len(alice) - 5000))
        k = int(rng.integers(500, 5000))
        parts.append(alice[o:o + k])
        total += k
    return b"".join(parts)[:n]


def _split_stats():
    import ctypes
    a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.lib().cz_split_stats(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def _check_vs_oracle(streams, caps, wbits):
    outs, st, lens, cons = batch.inflate_batch(streams, caps, wbits)
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, wbits)
    for i in range(len(streams)):
        assert st[i] == ref_st[i], (i, st[i], ref_st[i])
        if st[i] in (1, 2):
            assert outs[i] == ref_outs[i], "stream %d differs" % i
    return outs, st, cons


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_own_long_stream_decodes_through_ordinary_api(alice, wbits):
    data = _text(alice, 24_000_000, 3)
    stream, idx = batch.deflate_segmented(data, level=6, window_bits=wbits, segment_bytes=1 << 20)
    t0, s0 = _split_stats()
    outs, st, cons = _check_vs_oracle([stream, stream + b"trailing garbage"], [len(data), len(data) + 100], wbits)
    t1, s1 = _split_stats()
    assert (t1 - t0, s1 - s0) == (2, 2), "both long streams must have been decoded by the block-parallel path"
    assert list(st) == [2, 2] and outs[0] == data and outs[1] == data
    assert list(cons) == [len(stream), len(stream)]


def test_zlib_full_flush_and_sync_flush_streams(alice):
    data = _text(alice, 12_000_000, 5)
    for flush in (zlib.Z_FULL_FLUSH, zlib.Z_SYNC_FLUSH):
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        parts = []
        for o in range(0, len(data), 700_000):
            parts.append(c.compress(data[o:o + 700_000]))
            parts.append(c.flush(flush))
        parts.append(c.flush())
        s = b"".join(parts)
        t0, s0 = _split_stats()
        outs, st, cons = _check_vs_oracle([s], [len(data)], 31)
        t1, s1 = _split_stats()
        # flush points play no role: both decode on the block-parallel path
        assert (t1 - t0, s1 - s0) == (1, 1)
        assert st[0] == 2 and outs[0] == data and cons[0] == len(s)


def test_plain_long_stream_and_false_markers(alice):
    data = _text(alice, 6_000_000, 7)
    s = zlib.compress(data, 6)
    t0, s0 = _split_stats()
    outs, st, _ = _check_vs_oracle([s], [len(data)], 15)
    t1, s1 = _split_stats()
    assert st[0] == 2 and outs[0] == data
    assert (t1 - t0, s1 - s0) == (1, 1), "a zlib-made stream without any flush point must take the block-parallel path"
    # stored blocks whose DATA is full of 00 00 ff ff: every candidate cut inside them is false
    noisy = (b"\x00\x00\xff\xff" * 5000 + alice[:30000]) * 60
    for lvl in (0, 6):
        c = zlib.compressobj(lvl, zlib.DEFLATED, 15)
        s = c.compress(noisy[:3_000_000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(noisy[3_000_000:]) + c.flush()
        outs, st, _ = _check_vs_oracle([s], [len(noisy)], 15)
        assert st[0] == 2 and outs[0] == noisy


def test_damaged_long_streams_match_the_oracle(alice):
    data = _text(alice, 10_000_000, 9)
    stream, _ = batch.deflate_segmented(data, level=6, window_bits=15, segment_bytes=1 << 20)
    bad_mid = bytearray(stream); bad_mid[len(stream) // 2] ^= 0x10
    bad_trailer = bytearray(stream); bad_trailer[-2] ^= 0x01
    cases = [bytes(bad_mid), bytes(bad_trailer), stream[:len(stream) // 3], stream[:-3]]
    _check_vs_oracle(cases, [len(data)] * len(cases), 15)
    # output slot too small: NeedOutput with the slot filled, as zlib does
    outs, st, _ = _check_vs_oracle([stream], [len(data) - 12345], 15)
    assert st[0] == 1 and outs[0] == data[:len(data) - 12345]


def test_streaming_decoder_on_long_own_stream(alice):
    data = _text(alice, 9_000_000, 11)
    stream, _ = batch.deflate_segmented(data, level=6, window_bits=31, segment_bytes=1 << 20)
    d = dec.Interface.zlib_cuda(dec.ZlibMode.Auto)
    v = Vec()
    r = d.decode_vec_full(stream, v)
    assert r.status == dec.DecodeStatus.Finished and r.input_remain == 0 and v.as_bytes() == data
