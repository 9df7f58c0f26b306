"""Multi-GPU host entry points (devices_mask): units / segments are dealt to the devices in contiguous byte-balanced ranges,
no collective, host-side gather. Results must be byte-identical to the single-device run. Skipped with fewer than 2 GPUs."""
import zlib

import numpy as np
import pytest

from compu_b200 import _lib, batch
from helpers import oracle_inflate

pytestmark = pytest.mark.gpu


def _ndev():
    return _lib.lib().cz_device_count()


@pytest.mark.skipif("_ndev() < 2")
def test_inflate_batch_two_devices(alice):
    rng = np.random.default_rng(21)
    big = alice * 10
    datas = []
    for i in range(300):
        n = int(rng.choice([0, 100, 5000, 65536, 200000]))
        o = int(rng.integers(0, len(big) - n))
        datas.append(big[o:o + n])
    streams = [zlib.compress(d, 6) for d in datas]
    caps = [len(d) for d in datas]
    mask = (1 << min(_ndev(), 8)) - 1
    outs, st, lens, cons = batch.inflate_batch(streams, caps, 15, devices_mask=mask)
    outs1, st1, lens1, cons1 = batch.inflate_batch(streams, caps, 15, devices_mask=1)
    assert outs == outs1 == datas and list(st) == list(st1) and list(cons) == list(cons1)


@pytest.mark.skipif("_ndev() < 2")
def test_deflate_two_devices_identical_bytes(alice):
    mask = (1 << min(_ndev(), 8)) - 1
    data = (alice * 150)[:20_000_000]
    s_multi, idx_multi = batch.deflate_segmented(data, level=6, window_bits=31, segment_bytes=1 << 20, devices_mask=mask)
    s_one, idx_one = batch.deflate_segmented(data, level=6, window_bits=31, segment_bytes=1 << 20, devices_mask=1)
    assert s_multi == s_one and list(idx_multi) == list(idx_one)  # bytes do not depend on the device count
    assert zlib.decompress(s_multi, 31) == data
    back = batch.inflate_segmented(s_multi, len(data), idx_multi, window_bits=31, segment_bytes=1 << 20, devices_mask=mask)
    assert back == data
    bufs = [data[i * 700_000:(i + 1) * 700_000] for i in range(28)]
    a, st = batch.deflate_batch(bufs, level=6, window_bits=15, devices_mask=mask)
    b, st1 = batch.deflate_batch(bufs, level=6, window_bits=15, devices_mask=1)
    assert a == b and (st == 2).all()


@pytest.mark.skipif("_ndev() < 2")
def test_streaming_objects_on_the_second_device(alice):
    """cz_set_stream_device places new Decoder / Encoder objects on another GPU; bytes and statuses do not depend on it."""
    from compu_b200 import Vec
    from compu_b200 import decoder as dec
    from compu_b200 import encoder as enc
    L = _lib.lib()
    data = alice * 3
    opts = lambda: enc.ZlibOptions().mode(enc.ZlibMode.Zlib).compression(6)
    try:
        assert L.cz_set_stream_device(1) == 0
        e = enc.Interface.zlib_cuda(opts())
        d = dec.Interface.zlib_cuda(dec.ZlibMode.Zlib)
    finally:
        assert L.cz_set_stream_device(0) == 0   # (objects stay where they were created)
    cv, cv0, pv = Vec(), Vec(), Vec()
    assert e.encode_vec_full(data, cv, enc.EncodeOp.Finish).status == enc.EncodeStatus.Finished
    enc.Interface.zlib_cuda(opts()).encode_vec_full(data, cv0, enc.EncodeOp.Finish)
    assert cv.as_bytes() == cv0.as_bytes() and zlib.decompress(cv.as_bytes()) == data
    assert d.decode_vec_full(cv.as_bytes(), pv).status == dec.DecodeStatus.Finished
    assert pv.as_bytes() == data
