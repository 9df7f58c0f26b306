"""The resumable decode path behind cz_decode, on the CPU emulator: ONE long-lived stream fed and drained in arbitrary
pieces. The host algorithm of compu_b200/csrc/host.cu (cz_decode) is modelled here in Python, call for call, over the SAME
kernel source (inflate_kernel.cuh, resumable mode), and compared with the oracle driven with the same calls: status,
input_remain, output_remain and bytes of EVERY call must be zlib's (/root/reference/src/decoder/mod.rs:459-486).
The GPU version of this fuzz is tests/test_gpu_inflate.py::test_streaming_* and tools/fuzz_stream_gpu.py."""
import ctypes
import random
import zlib

import numpy as np
import pytest

import oracle
import simlib
from helpers import gzip_member, make_data, zcomp


class ResumeState(ctypes.Structure):
    _fields_ = [("bit_pos", ctypes.c_uint64), ("total_out", ctypes.c_uint64), ("phase", ctypes.c_uint32), ("bfinal", ctypes.c_uint32),
                ("nlit", ctypes.c_uint32), ("ndist", ctypes.c_uint32), ("stored_left", ctypes.c_uint32), ("wrap", ctypes.c_uint32),
                ("hist_len", ctypes.c_uint32), ("adler", ctypes.c_uint32), ("crc", ctypes.c_uint32), ("pend_len", ctypes.c_uint32),
                ("pend_dist", ctypes.c_uint32), ("pend_lit", ctypes.c_uint32), ("lens", ctypes.c_uint8 * 320)]


class SimStreamDecoder:
    """cz_decode, modelled: staged tail + 32 KiB history + ResumeState, one emulated launch per call that can make progress."""

    def __init__(self, wbits, seed=1):
        self.L = simlib.lib()
        assert self.L.sim_resume_state_bytes() == ctypes.sizeof(ResumeState)
        self.wbits = wbits
        self.seed = seed
        self.rs = ResumeState()
        self.rs.adler = 1
        self.carry = b""
        self.hist = b""
        self.done = False
        self.error = 0
        self.launches = 0

    def decode(self, data, cap):
        if self.error:
            return self.error, len(data), cap, b""
        if self.done:
            return 2, len(data), cap, b""
        unit = self.carry + data
        hist = self.hist
        buf = np.full(len(hist) + cap + 64, 0xEE, dtype=np.uint8)
        buf[:len(hist)] = np.frombuffer(hist, dtype=np.uint8)
        src = np.frombuffer(unit + b"\xff" * 16, dtype=np.uint8).copy()  # whatever follows the unit must not matter
        out_len = ctypes.c_uint64(0)
        status = ctypes.c_int32(-99)
        self.rs.hist_len = len(hist)
        self.launches += 1
        r = self.L.sim_inflate_resume(src.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(unit)),
                                      ctypes.c_void_p(buf.ctypes.data + len(hist)), ctypes.c_uint64(cap), ctypes.byref(self.rs), self.wbits,
                                      ctypes.byref(out_len), ctypes.byref(status), ctypes.c_uint64(self.seed + self.launches))
        assert r == 0
        assert (buf[len(hist) + cap:] == 0xEE).all(), "wrote past the slot"
        n = out_len.value
        out = buf[len(hist):len(hist) + n].tobytes()
        self.hist = (hist + out)[-32768:]
        st = status.value
        B = self.rs.bit_pos
        if st == 2:
            self.done = True
            used = (B + 7) >> 3
            remain = len(unit) - used
            assert remain <= len(data)
            return 2, remain, cap - n, out
        if st < 0 or st == 3:
            self.error = st
            return st, None, cap - n, out
        if st == 0:  # out of input: everything is taken, the unconsumed tail stays staged
            self.carry = unit[B >> 3:]
            self.rs.bit_pos = B & 7
            remain = 0
        else:        # slot full: zlib has pulled the bytes that hold the bits it used, the rest goes back to the caller
            used = (B + 7) >> 3
            self.carry = unit[B >> 3:used]
            self.rs.bit_pos = B & 7
            remain = len(unit) - used
            assert remain <= len(data)
        made = n
        consumed = len(data) - remain
        code = 0 if remain == 0 else 1
        if consumed == 0 and made == 0:
            code = 1  # no progress: Z_BUF_ERROR -> NeedOutput (src/decoder/mod.rs:481)
        return code, remain, cap - n, out


def oracle_calls(wbits):
    L = oracle.lib()
    st = L.oz_decoder_new(wbits)

    def call(data, cap):
        out = ctypes.create_string_buffer(max(cap, 1))
        r = L.oz_decode(st, data, len(data), out, cap)
        return r.status, r.input_remain, r.output_remain, out.raw[:cap - r.output_remain]
    return call, lambda: L.oz_decoder_free(st)


def drive(stream, wbits, rng, in_sizes, out_sizes, seed, limit=20000, make=None):
    """Feeds `stream` in random pieces with random output sizes to both decoders; every call must agree. `make(wbits, seed)`
    builds the decoder under test (default: the emulated one); it needs .decode(data, cap) and .launches."""
    sim = (make or SimStreamDecoder)(wbits, seed)
    ocall, ofree = oracle_calls(wbits)
    pos = 0
    produced = b""
    calls = 0
    pending = b""  # bytes the decoders handed back (input_remain) are presented again
    try:
        while calls < limit:
            calls += 1
            take = rng.choice(in_sizes)
            chunk = pending + stream[pos:pos + take]
            pos += take
            cap = rng.choice(out_sizes)
            s1 = sim.decode(chunk, cap)
            s2 = ocall(chunk, cap)
            what = "call %d (in %d, out %d, stream pos %d)" % (calls, len(chunk), cap, pos)
            assert s1[0] == s2[0], "%s: status %r, oracle %r" % (what, s1[0], s2[0])
            assert s1[3] == s2[3], "%s: bytes differ" % what
            assert s1[2] == s2[2], "%s: output_remain %r, oracle %r" % (what, s1[2], s2[2])
            if s2[0] >= 0 and s2[0] != 3:
                assert s1[1] == s2[1], "%s: input_remain %r, oracle %r" % (what, s1[1], s2[1])
            produced += s1[3]
            if s2[0] == 2 or s2[0] < 0:
                break
            pending = chunk[len(chunk) - s2[1]:] if s2[1] else b""
            if pos >= len(stream) and not pending and s2[0] == 0 and not chunk:
                break  # truncated stream, nothing more to give
    finally:
        ofree()
    return produced, sim.launches, calls


def test_sim_stream_reference_pattern_on_golden_gzip(golden):
    # README.md:33-49 / tests/decoder.rs: whole input, growing output; and the golden files byte by byte
    for plain, gz in golden:
        rng = random.Random(1)
        out, launches, calls = drive(gz, 31, rng, [len(gz)], [1, 100, 4096], 3)
        assert out == plain
    plain, gz = golden[0]
    out, launches, calls = drive(gz, 47, random.Random(2), [1], [1, 2, 3], 5)
    assert out == plain


@pytest.mark.parametrize("wbits", [15, 31, -15, 47])
def test_sim_stream_fuzz_call_patterns(alice, wbits):
    rng = random.Random(500 + wbits)
    for case in range(9):
        kind = rng.randrange(5)
        n = rng.choice([0, 1, 7, 300, 5000, 40000, 90000])
        data = make_data(rng, kind, n, alice)
        lvl = rng.choice([0, 1, 6, 9])
        strat = rng.choice([0, 0, 2, 3, 4])
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
        in_sizes = rng.choice([[1, 2, 3], [5, 64, 700], [4096], [len(s) + 100], [1, 5000]])
        out_sizes = rng.choice([[0, 1, 2], [1, 17, 300], [4096], [len(data) + 10], [0, 1, 70000], [258, 259, 3]])
        out, launches, calls = drive(s, wbits, rng, in_sizes, out_sizes, case, limit=2500)
        if mode in (0, 3):
            assert out == data if calls < 2500 else data.startswith(out)
        assert launches <= calls  # at most one launch per call: device work is linear in the number of calls


def test_sim_stream_multi_block_and_stored_and_header_fields(alice):
    rng = random.Random(77)
    # many blocks of all three types in one stream (Z_FULL_FLUSH / level changes), gzip member with every optional field
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    raw = c.compress(alice[:8000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(bytes(rng.getrandbits(8) for _ in range(3000))) + \
        c.flush(zlib.Z_SYNC_FLUSH) + c.compress(b"z" * 70000) + c.flush()
    plain = zlib.decompress(raw, -15)
    member = gzip_member(raw, plain, fextra=b"AB\x03\x00xyz", fname=b"name", fcomment=b"comment", fhcrc=True)
    for in_sizes, out_sizes in [([1], [4096]), ([997], [1009]), ([len(member)], [1]), ([3, 50000], [0, 65536]), ([64], [258])]:
        out, launches, calls = drive(member, 31, rng, in_sizes, out_sizes, 11, limit=400000 if out_sizes != [1] else 1500)
        if out_sizes != [1]:
            assert out == plain
        else:
            assert plain.startswith(out) and len(out) >= 1499
