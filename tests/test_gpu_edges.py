"""GPU parity tests for the edges the reference's tests leave open (VERDICT r1 "untested parity edges"), through the C ABI:
gzip optional header fields against the oracle, strategy semantics of the emitted blocks, size per level against the
oracle's codec at the same level, the encoder's no-progress calls against the oracle, and re-entrancy across handles on
several host threads (SURVEY.md 8b "Threading"; the rayon baseline creates one handle per worker). Run with -m gpu."""
import threading
import zlib

import pytest

import oracle_backend
from compu_b200 import batch
from compu_b200 import decoder as dec
from compu_b200 import encoder as enc
from helpers import assert_inflate_parity, gzip_member, oracle_inflate, zcomp
from test_sim_edges import LEVEL_BOUND, check_strategy_semantics, header_cases, strategy_inputs

pytestmark = pytest.mark.gpu


def test_gzip_optional_header_fields(alice):
    data, cases = header_cases(alice)
    caps = [len(data)] * len(cases)
    for wbits in (31, 47):
        ref_outs, ref_st, _ = oracle_inflate(cases, caps, wbits)
        outs, st, _, cons = batch.inflate_batch(cases, caps, wbits)
        assert_inflate_parity(outs, st, ref_outs, ref_st, "wbits %d" % wbits)
        for i in range(len(cases)):
            if st[i] == 2:
                assert cons[i] == len(cases[i])


def test_gzip_header_truncated_at_every_byte(alice):
    data = alice[:600]
    raw = zcomp(data, 6, -15)
    full = gzip_member(raw, data, fextra=b"xy\x03\x00abc", fname=b"name.txt", fcomment=b"comment", fhcrc=True)
    hdr_len = len(full) - len(raw) - 8
    streams = [full[:k] for k in range(0, hdr_len + 6)]
    badm = gzip_member(raw, data, fextra=b"xy\x03\x00abc", fname=b"name.txt", fcomment=b"comment", fhcrc=True, bad_hcrc=True)
    streams += [badm[:k] for k in range(hdr_len - 3, hdr_len + 3)]
    caps = [len(data)] * len(streams)
    ref_outs, ref_st, _ = oracle_inflate(streams, caps, 31)
    outs, st, _, _ = batch.inflate_batch(streams, caps, 31)
    assert_inflate_parity(outs, st, ref_outs, ref_st, "cut")


def test_gzip_header_fields_through_the_streaming_decoder(alice):
    # the same kind of members through Decoder::decode (FHCRC members are refused by the speculative split and must take
    # the serial path: a 3 MiB member with full-flush-free content exercises exactly that)
    big = (alice * 22)[:3 << 20]
    members = [gzip_member(zcomp(alice, 6, -15), alice, fextra=b"AB\x02\x00zz", fname=b"alice", fcomment=b"c", fhcrc=True),
               gzip_member(zcomp(big, 1, -15), big, fname=b"big", fhcrc=True)]
    for m, plain in zip(members, (alice, big)):
        d = dec.Interface.zlib_cuda(dec.ZlibMode.Gzip)
        od = oracle_backend.oracle_decoder(dec.ZlibMode.Gzip)
        out, oout = bytearray(len(plain)), bytearray(len(plain))
        r = d.decode(m, out)
        ro = od.decode(m, oout)
        assert (r.status, r.input_remain, r.output_remain) == (ro.status, ro.input_remain, ro.output_remain)
        assert bytes(out) == plain


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
def test_strategy_semantics(alice, strategy):
    datas = strategy_inputs(alice)
    streams, st = batch.deflate_batch(datas, level=6, window_bits=-15, strategy=strategy, segment_bytes=65536)
    assert (st == 2).all()
    # a stream of the batch API = segments | 03 00: the inspector walks all of it
    check_strategy_semantics(streams, datas, strategy)


@pytest.mark.parametrize("level", [1, 3, 6, 9])
def test_size_per_level_against_oracle_level(alice, level):
    datas = [alice[i * 65536:(i + 1) * 65536] for i in range(2)]
    streams, st = batch.deflate_batch(datas, level=level, window_bits=-15, segment_bytes=65536)
    assert (st == 2).all() and [zlib.decompress(s, -15) for s in streams] == datas
    ours = sum(len(s) for s in streams)
    ref = sum(len(zcomp(d, level, -15)) for d in datas)
    assert ours <= ref * LEVEL_BOUND[level], "level %d: %d bytes vs zlib's %d (%.2f %%)" % (level, ours, ref, 100.0 * ours / ref - 100)


def test_encoder_no_progress_calls_match_the_oracle(alice):
    """Calls zlib's deflate() refuses before touching the stream (no output space; nothing to do: no pending output, no new
    input, flush not stronger than the last; new input after the end), mapped by compu's glue
    (/root/reference/src/encoder/mod.rs:356-369). Statuses and remainders must be the oracle's, call by call."""
    P, F, E = enc.EncodeOp.Process, enc.EncodeOp.Flush, enc.EncodeOp.Finish
    scripts = [
        [(b"", F), (b"", F), (b"", F), (b"", E)],
        [(b"", P), (b"", P), (b"abc", P), (b"", P), (b"", F), (b"", F), (b"", P), (b"", E), (b"", E)],
        [(alice[:5000], P), (b"", F), (b"", F), (alice[5000:6000], F), (b"", P), (b"", E), (b"x", E), (b"", P)],
        [(alice[:100], E), (b"", E), (b"", F)],
    ]
    for script in scripts:
        for zero_out_every_other in (False, True):
            ge = enc.Interface.zlib_cuda(enc.ZlibOptions().mode(enc.ZlibMode.Zlib).compression(6))
            oe = oracle_backend.oracle_encoder(enc.ZlibOptions().mode(enc.ZlibMode.Zlib).compression(6))
            for k, (data, op) in enumerate(script):
                size = 0 if (zero_out_every_other and k % 2 == 1) else 1 << 16
                go, oo = bytearray(size), bytearray(size)
                rg = ge.encode(data, go, op)
                ro = oe.encode(data, oo, op)
                what = "script %r call %d (%r, out %d)" % ([(len(d), o) for d, o in script], k, op, size)
                assert rg.status == ro.status, "%s: %r, oracle %r" % (what, rg.status, ro.status)
                assert rg.input_remain == ro.input_remain, "%s: input_remain %d, oracle %d" % (what, rg.input_remain, ro.input_remain)
                if ro.output_remain == size:
                    assert rg.output_remain == size, "%s: the oracle wrote nothing" % what


def test_eight_threads_with_their_own_handles(alice):
    """8 host threads, each with its own Decoder + Encoder (handles are one-thread-at-a-time, different handles must be
    re-entrant), plus concurrent cz_inflate_batch / cz_deflate_batch calls; every result is checked against zlib."""
    errors = []
    pieces = [alice[(i * 9973) % 90000:(i * 9973) % 90000 + 30000 + 1000 * i] for i in range(8)]

    def worker(i):
        try:
            from compu_b200 import Vec
            data = pieces[i]
            for rep in range(3):
                e = enc.Interface.zlib_cuda(enc.ZlibOptions().mode(enc.ZlibMode.Zlib).compression(6))
                d = dec.Interface.zlib_cuda(dec.ZlibMode.Zlib)
                cv, pv = Vec(), Vec()
                r = e.encode_vec_full(data, cv, enc.EncodeOp.Finish)
                assert r.status == enc.EncodeStatus.Finished
                assert zlib.decompress(cv.as_bytes()) == data
                r = d.decode_vec_full(zlib.compress(data, 6), pv)
                assert r.status == dec.DecodeStatus.Finished and pv.as_bytes() == data
                streams = [zlib.compress(data[k:k + 4096], 6) for k in range(0, len(data), 4096)]
                outs, st, _, _ = batch.inflate_batch(streams, [4096] * len(streams), 15)
                assert (st == 2).all() and b"".join(outs) == data
                comp, st = batch.deflate_batch([data, data[::-1]], level=6, window_bits=31)
                assert (st == 2).all() and zlib.decompress(comp[0], 31) == data and zlib.decompress(comp[1], 31) == data[::-1]
        except BaseException as ex:  # noqa: BLE001 - reported below with the thread index
            errors.append((i, repr(ex)))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_stream_device_selection_rejects_missing_devices():
    from compu_b200 import _lib
    L = _lib.lib()
    assert L.cz_set_stream_device(0) == 0
    assert L.cz_set_stream_device(63) != 0 and L.cz_set_stream_device(-1) != 0   # CZ_E_NO_DEVICE, the selection stays
    d = dec.Interface.zlib_cuda(dec.ZlibMode.Zlib)
    assert d is not None
