"""Pins the CPU oracle (oracle/compu_oracle.c over zlib 1.3) to the reference's golden vectors, known answers and test
protocols. CPU only."""
import ctypes
import json
import os
import zlib

import numpy as np
import pytest

import oracle
import protocols
from compu_b200.decoder import DecodeError, DecodeStatus, Detection
from compu_b200.decoder import ZlibMode as DMode
from compu_b200.encoder import EncodeOp, EncodeStatus, ZlibMode, ZlibOptions
from conftest import GOLDEN, read_golden
from oracle_backend import oracle_decoder, oracle_encoder


@pytest.mark.parametrize("mode", [DMode.Gzip, DMode.Auto])
def test_reference_decoder_protocol_on_golden_gzip(golden, mode):
    # should_decode_zlib_ng_gzip, tests/decoder.rs:141-150
    d = oracle_decoder(mode)
    for data, comp in golden:
        protocols.decoder_test_case(d, data, comp)


@pytest.mark.parametrize("mode,det", [(ZlibMode.Gzip, Detection.Gzip), (ZlibMode.Zlib, Detection.Zlib),
                                      (ZlibMode.Deflate, Detection.Unknown)])
def test_reference_encoder_protocols(golden, mode, det):
    # tests/encoder.rs:205-302 (default level 9)
    e = oracle_encoder(ZlibOptions().mode(mode))
    d = oracle_decoder(DMode(int(mode)))
    for data, _ in golden:
        protocols.encoder_test_case(e, d, data, det)
        protocols.encoder_empty_final(e, d, data)
    protocols.doc_chunked_roundtrip(e, d, golden[1][0][:4000])


def test_known_answers(golden):
    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for fn in ("10x10y", "alice29.txt"):
        data = read_golden(fn)
        info = ka["files"][fn]
        L = oracle.lib()
        buf = np.frombuffer(data, dtype=np.uint8)
        p = buf.ctypes.data_as(ctypes.c_void_p)
        assert "%08x" % L.oz_crc32(0, p, len(data)) == info["crc32"]
        assert "%08x" % L.oz_adler32(1, p, len(data)) == info["adler32"]
        for lvl, sizes in info["levels"].items():
            for name, mode in (("raw", ZlibMode.Deflate), ("zlib", ZlibMode.Zlib), ("gzip", ZlibMode.Gzip)):
                e = oracle_encoder(ZlibOptions().mode(mode).compression(int(lvl)))
                out = bytearray(len(data) + 200)
                r = e.encode(data, out, EncodeOp.Finish)
                assert r.status == EncodeStatus.Finished
                assert len(out) - r.output_remain == sizes[name]
    # L9 gzip of 10x10y into 20 bytes -> NeedOutput, in_rem 0, out_rem 0; +100 -> Finished, 26 bytes
    e = oracle_encoder(ZlibOptions())
    out = bytearray(20)
    r = e.encode(read_golden("10x10y"), out, EncodeOp.Finish)
    assert (r.status, r.input_remain, r.output_remain) == (EncodeStatus.NeedOutput, 0, 0)
    out2 = bytearray(100)
    r = e.encode(b"", out2, EncodeOp.Finish)
    assert r.status == EncodeStatus.Finished
    assert (bytes(out) + bytes(out2[:100 - r.output_remain])).hex() == ka["gzip_l9_10x10y_hex"]


def test_oracle_error_and_edge_statuses(golden):
    data, comp = golden[1]
    d = oracle_decoder(DMode.Gzip)
    out = bytearray(len(data))
    # half the input -> NeedInput, all consumed (SURVEY.md §8c probe)
    r = d.decode(comp[:len(comp) // 2], out)
    assert r.status == DecodeStatus.NeedInput and r.input_remain == 0 and r.output_remain > 0
    # zero-progress call: raw Z_BUF_ERROR is mapped to NeedOutput (src/decoder/mod.rs:481)
    r = d.decode(b"", out)
    assert r.status == DecodeStatus.NeedOutput
    d.reset()
    bad = bytearray(comp)
    bad[len(bad) // 2] ^= 0x40
    r = d.decode(bytes(bad), out)
    assert r.status == DecodeError(-3)
    assert d.describe_error(r.status) == "data error"
    d.reset()
    bad = bytearray(comp)
    bad[-6] ^= 1  # CRC
    r = d.decode(bytes(bad), out)
    assert r.status == DecodeError(-3)


def test_oracle_batch_matches_python_zlib(alice):
    L = oracle.lib()
    chunks = [alice[i:i + 65536] for i in range(0, len(alice), 65536)]
    streams = [zlib.compress(c, 6) for c in chunks]
    from simlib import pack
    inbuf, in_off = pack(streams)
    out_off = np.zeros(len(chunks) + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum([len(c) for c in chunks])
    out = np.zeros(len(alice) + 16, dtype=np.uint8)
    lens = np.zeros(len(chunks), dtype=np.uint64)
    st = np.zeros(len(chunks), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    bad = L.oz_inflate_batch(len(chunks), p(inbuf), p(in_off), p(out), p(out_off), p(lens), p(st), 15, 2)
    assert bad == 0 and (st == 2).all()
    assert out[:len(alice)].tobytes() == alice


def test_combine_helpers_match_zlib():
    import simlib
    L = oracle.lib()
    S = simlib.lib()
    rng = np.random.default_rng(3)
    for _ in range(50):
        a = rng.integers(0, 256, int(rng.integers(0, 5000)), dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, int(rng.integers(0, 200000)), dtype=np.uint8).tobytes()
        assert S.sim_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert S.sim_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
        assert L.oz_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
