"""The streaming Decoder on the GPU (cz_decode through the C ABI): zlib's inflate() contract call for call — status,
input_remain, output_remain and bytes of EVERY call equal to the oracle's under random call patterns — and linear device
work: one launch per call that can make progress, a 256 MiB stream in 64 KiB chunks in seconds. Run with -m gpu."""
import ctypes
import random
import time
import zlib

import numpy as np
import pytest

from compu_b200 import _lib
from helpers import gzip_member, make_data, zcomp
from test_sim_stream import drive

pytestmark = pytest.mark.gpu


class GpuStreamDecoder:
    """Raw cz_decode calls (no Python-side policy in between)."""

    def __init__(self, wbits, seed=0):
        self.L = _lib.lib()
        _lib.require_device()
        self.h = self.L.cz_decoder_new(wbits)
        assert self.h, _lib.last_error()
        self.n0 = self.L.cz_launch_count()

    @property
    def launches(self):
        return self.L.cz_launch_count() - self.n0

    def decode(self, data, cap):
        out = ctypes.create_string_buffer(max(cap, 1))
        src = ctypes.create_string_buffer(data, max(len(data), 1))
        r = self.L.cz_decode(self.h, ctypes.cast(src, ctypes.c_void_p), len(data), ctypes.cast(out, ctypes.c_void_p), cap)
        st = r.status
        return st, (None if (st < 0 or st == 3) else r.input_remain), r.output_remain, out.raw[:cap - r.output_remain]

    def __del__(self):
        try:
            self.L.cz_decoder_free(self.h)
        except Exception:
            pass


@pytest.mark.parametrize("wbits", [15, 31, -15, 47])
def test_streaming_calls_equal_the_oracle_call_for_call(alice, wbits):
    rng = random.Random(900 + wbits)
    for case in range(40):
        kind = rng.randrange(5)
        n = rng.choice([0, 1, 7, 300, 5000, 40000, 90000, 300000])
        data = make_data(rng, kind, n, alice)
        lvl = rng.choice([0, 1, 6, 9])
        strat = rng.choice([0, 0, 1, 2, 3, 4])
        wb = wbits if wbits != 47 else rng.choice([15, 31])
        s = zcomp(data, lvl, wb, strat)
        mode = rng.randrange(5)
        if mode == 1 and len(s) > 4:
            s = s[:rng.randrange(1, len(s))]
        elif mode == 2 and len(s) > 8:
            b = bytearray(s)
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            s = bytes(b)
        elif mode == 3:
            s = s + b"trailing bytes that belong to the caller"
        in_sizes = rng.choice([[1, 2, 3], [5, 64, 700], [4096], [len(s) + 100], [1, 5000], [65536]])
        out_sizes = rng.choice([[0, 1, 2], [1, 17, 300], [4096], [len(data) + 10], [0, 1, 70000], [258, 259, 3], [1 << 20]])
        out, launches, calls = drive(s, wbits, rng, in_sizes, out_sizes, case, limit=3000, make=GpuStreamDecoder)
        if mode in (0, 3) and calls < 3000:
            assert out == data
        assert launches <= calls + 2


def test_streaming_header_fields_blocks_of_every_type(alice):
    rng = random.Random(78)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = c.compress(alice[:30000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(bytes(rng.getrandbits(8) for _ in range(70000))) + \
        c.flush(zlib.Z_SYNC_FLUSH) + c.compress(b"z" * 100000) + c.flush()
    plain = zlib.decompress(raw, -15)
    member = gzip_member(raw, plain, fextra=b"AB\x03\x00xyz", fname=b"name", fcomment=b"comment", fhcrc=True)
    for in_sizes, out_sizes in [([1], [1 << 20]), ([997], [1009]), ([3, 50000], [0, 65536]), ([64], [258]), ([len(member)], [len(plain)])]:
        out, launches, calls = drive(member, 31, rng, in_sizes, out_sizes, 11, limit=400000, make=GpuStreamDecoder)
        assert out == plain


def test_256_mib_stream_in_64_kib_chunks_is_linear():
    """the README loop (README.md:33-49) over a long stream. Device work must be O(n): one launch per chunk, and
    the whole thing finishes in seconds, not in the hours a re-decode per call would take."""
    L = _lib.lib()
    corpus = np.frombuffer(open(__file__.replace("test_gpu_stream.py", "golden/alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert L.cz_synth_build_model(p(corpus), len(corpus), p(model)) == 0
    n = 4096
    offs = np.arange(n + 1, dtype=np.uint64) * 65536
    plain = np.empty(n * 65536, dtype=np.uint8)
    assert L.cz_synth_fill_host(0, 4242, n, p(plain), p(offs), p(model)) == 0
    data = plain.tobytes()
    stream = zlib.compress(data, 1)  # ONE zlib-made stream without flush points: nothing for the speculative split to use
    d = GpuStreamDecoder(15)
    out = ctypes.create_string_buffer(1 << 20)
    src = ctypes.create_string_buffer(stream, len(stream))
    base = ctypes.addressof(src)
    got = bytearray()
    t0 = time.perf_counter()
    calls = 0
    pos = 0
    status = 0
    while status != 2:
        k = min(65536, len(stream) - pos)
        give = k
        inp = base + pos
        while True:  # README loop: the same input again while the decoder asks for more output
            r = L.cz_decode(d.h, ctypes.c_void_p(inp), give, ctypes.cast(out, ctypes.c_void_p), 1 << 20)
            calls += 1
            got += out.raw[:(1 << 20) - r.output_remain]
            status = r.status
            assert status in (0, 1, 2), status
            inp += give - r.input_remain
            give = r.input_remain
            if status != 1:
                break
        pos += k
        assert pos < len(stream) or status == 2
    dt = time.perf_counter() - t0
    assert bytes(got) == data
    chunks = (len(stream) + 65535) // 65536
    assert d.launches <= calls + 2 and calls <= 4 * chunks + 4, (d.launches, calls, chunks)
    assert dt < 60, "256 MiB through cz_decode took %.1f s" % dt
    print("256 MiB zlib stream, %d chunks of 64 KiB, %d calls, %d launches, %.2f s (%.1f MB/s)" % (chunks, calls, d.launches, dt, len(data) / dt / 1e6))
