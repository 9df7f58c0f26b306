"""GPU parity tests for the encoder: the CUDA path through the C ABI, checked with the CPU oracle (compu's zlib glue — the
reference DEcoder must inflate GPU-encoded streams bit-exactly), with the sequential host model of the same decisions
(byte-for-byte), with the reference's own encoder test protocols, and for ratio against the oracle at level 6.
Run with -m gpu on a B200."""
import ctypes
import random
import zlib

import numpy as np
import pytest

import oracle_backend
import protocols
from compu_b200 import _lib, batch
from compu_b200 import decoder as dec
from compu_b200 import encoder as enc
from compu_b200.decoder import Detection
from helpers import make_data, oracle_inflate, zcomp

pytestmark = pytest.mark.gpu

MODES = [(enc.ZlibMode.Gzip, dec.ZlibMode.Gzip, Detection.Gzip), (enc.ZlibMode.Zlib, dec.ZlibMode.Zlib, Detection.Zlib),
         (enc.ZlibMode.Deflate, dec.ZlibMode.Deflate, Detection.Unknown)]


def cuda_encoder(mode, level=None):
    o = enc.ZlibOptions().mode(mode)
    if level is not None:
        o = o.compression(level)
    e = enc.Interface.zlib_cuda(o)
    assert e is not None, _lib.last_error()
    return e


@pytest.mark.parametrize("emode,dmode,detect", MODES)
def test_reference_encoder_protocol(golden, emode, dmode, detect):
    # should_encode_zlib_ng_{gzip,zlib,deflate} (tests/encoder.rs:205-302) with the CUDA backend on both sides
    e = cuda_encoder(emode)
    d = dec.Interface.zlib_cuda(dmode)
    for data, _ in golden:
        comp = protocols.encoder_test_case(e, d, data, detect)
        # and the reference DEcoder (oracle = compu's glue over zlib) inflates the GPU-encoded stream bit-exactly
        od = oracle_backend.oracle_decoder(dmode)
        out = bytearray(len(data))
        r = od.decode(comp, out)
        assert r.status == dec.DecodeStatus.Finished and bytes(out) == data and r.input_remain == 0


@pytest.mark.parametrize("emode,dmode,detect", MODES)
def test_reference_empty_final_and_doc_chunked(golden, emode, dmode, detect):
    e = cuda_encoder(emode)
    d = dec.Interface.zlib_cuda(dmode)
    for data, _ in golden:
        protocols.encoder_empty_final(e, d, data)
    protocols.doc_chunked_roundtrip(e, d, golden[1][0][:2000], chunk=400)


def test_mixed_backends_roundtrip(alice):
    # GPU encoder -> oracle decoder, oracle encoder -> GPU decoder, through the streaming API
    for emode, dmode, _ in MODES:
        e = cuda_encoder(emode, 6)
        od = oracle_backend.oracle_decoder(dmode)
        from compu_b200 import Vec
        cv = Vec()
        r = e.encode_vec_full(alice, cv, enc.EncodeOp.Finish)
        assert r.status == enc.EncodeStatus.Finished
        out = Vec()
        r = od.decode_vec_full(cv.as_bytes(), out)
        assert r.status == dec.DecodeStatus.Finished and out.as_bytes() == alice
        oe = oracle_backend.oracle_encoder(enc.ZlibOptions().mode(emode).compression(6))
        cv2 = Vec()
        oe.encode_vec_full(alice, cv2, enc.EncodeOp.Finish)
        gd = dec.Interface.zlib_cuda(dmode)
        out2 = Vec()
        r = gd.decode_vec_full(cv2.as_bytes(), out2)
        assert r.status == dec.DecodeStatus.Finished and out2.as_bytes() == alice


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_batch_deflate_decodes_with_oracle(alice, wbits):
    rng = random.Random(900 + wbits)
    bufs = [make_data(rng, rng.randrange(5), rng.choice([0, 1, 2, 7, 100, 1000, 5000, 20000, 70000, 300000]), alice)
            for _ in range(200)]
    streams, st = batch.deflate_batch(bufs, level=6, window_bits=wbits, segment_bytes=65536)
    assert (st == 2).all()
    outs, ost, _ = oracle_inflate(streams, [len(b) for b in bufs], wbits)
    assert list(ost) == [2] * len(bufs)
    assert outs == bufs
    # and our own decoder agrees
    outs2, st2, _, cons = batch.inflate_batch(streams, [len(b) for b in bufs], wbits)
    assert (st2 == 2).all() and outs2 == bufs and list(cons) == [len(s) for s in streams]


def test_batch_deflate_matches_sequential_model(alice):
    from test_sim_deflate import model_lib, model_segment
    L = model_lib()
    rng = random.Random(3)
    bufs = [make_data(rng, k % 5, n, alice) for k, n in enumerate([100000, 3000, 50000, 65536, 17, 200000])]
    streams, st = batch.deflate_batch(bufs, level=6, window_bits=-15, segment_bytes=1 << 20)
    assert (st == 2).all()
    for b, s in zip(bufs, streams):
        assert s == model_segment(L, b) + b"\x03\x00"


@pytest.mark.parametrize("level,strategy", [(0, 0), (1, 0), (3, 0), (9, 0), (6, 1), (6, 2), (6, 3), (6, 4)])
def test_levels_and_strategies(alice, level, strategy):
    rng = random.Random(level * 10 + strategy)
    bufs = [alice, make_data(rng, 1, 50000, alice), make_data(rng, 2, 30000, alice), make_data(rng, 4, 80000, alice)]
    streams, st = batch.deflate_batch(bufs, level=level, window_bits=15, strategy=strategy)
    assert (st == 2).all()
    for b, s in zip(bufs, streams):
        assert zlib.decompress(s) == b


def test_ratio_within_3pct_of_level6(alice):
    # north star: compression ratio within 3 % of the reference codec at the same level (zlib 1.3 stands in for zlib-ng)
    from compu_b200 import _lib as lib_
    L = lib_.lib()
    n = 256
    offs = np.arange(n + 1, dtype=np.uint64) * 65536
    corpus = np.frombuffer(alice, dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert L.cz_synth_build_model(p(corpus), len(corpus), p(model)) == 0
    for kind in (0, 1, 2, 3):
        buf = np.empty(n * 65536, dtype=np.uint8)
        assert L.cz_synth_fill_host(kind, 77, n, p(buf), p(offs), p(model)) == 0
        data = buf.tobytes()
        # one stream, 1 MiB segments (cfg3 shape) against whole-stream zlib level 6
        stream, idx = batch.deflate_segmented(data, level=6, window_bits=15, segment_bytes=1 << 20)
        ref = len(zlib.compress(data, 6))
        assert zlib.decompress(stream) == data
        assert len(stream) <= ref * 1.03, "kind %d: %d vs zlib L6 %d (%.2f%%)" % (kind, len(stream), ref, 100.0 * len(stream) / ref - 100)
        # 64 KiB independent streams (cfg2 shape) against per-stream zlib level 6
        bufs = [data[i * 65536:(i + 1) * 65536] for i in range(n)]
        streams, st = batch.deflate_batch(bufs, level=6, window_bits=15)
        ours = sum(len(s) for s in streams)
        ref = sum(len(zlib.compress(b, 6)) for b in bufs)
        assert ours <= ref * 1.03, "kind %d (64 KiB streams): %d vs %d" % (kind, ours, ref)


@pytest.mark.parametrize("wbits", [15, 31, -15])
def test_segmented_stream_and_parallel_inflate(alice, wbits):
    data = (alice * 30)[:4_000_000 + 12345]
    stream, idx = batch.deflate_segmented(data, level=6, window_bits=wbits, segment_bytes=262144)
    assert len(idx) == (len(data) + 262143) // 262144 + 1
    # ONE valid stream for the reference decoder
    outs, ost, _ = oracle_inflate([stream], [len(data)], wbits)
    assert ost[0] == 2 and outs[0] == data
    # every segment ends with the full-flush marker at its index boundary
    for i in range(len(idx) - 1):
        assert stream[int(idx[i + 1]) - 4:int(idx[i + 1])] == b"\x00\x00\xff\xff"
    # segment-parallel inflate with the side index; checksum of checksums verified against the trailer
    back = batch.inflate_segmented(stream, len(data), idx, window_bits=wbits, segment_bytes=262144)
    assert back == data
    # a corrupted trailer is caught by the parallel combine
    if wbits != -15:
        bad = bytearray(stream)
        bad[-1 if wbits == 15 else -5] ^= 0x40
        with pytest.raises(RuntimeError):
            batch.inflate_segmented(bytes(bad), len(data), idx, window_bits=wbits, segment_bytes=262144)


def test_gzip_crc_combine_matches_whole(alice):
    # cfg4 in miniature: gzip trailer CRC-32 / ISIZE produced by the parallel combine equal zlib.crc32 of the whole input
    data = (alice * 12)[:1_500_000]
    stream, idx = batch.deflate_segmented(data, level=6, window_bits=31, segment_bytes=65536)
    crc = int.from_bytes(stream[-8:-4], "little")
    isz = int.from_bytes(stream[-4:], "little")
    assert crc == zlib.crc32(data) and isz == len(data)
    L = _lib.lib()
    a = zlib.crc32(data[:700000]); b = zlib.crc32(data[700000:])
    assert L.cz_crc32_combine(a, b, len(data) - 700000) == zlib.crc32(data)
    a = zlib.adler32(data[:700000]); b = zlib.adler32(data[700000:])
    assert L.cz_adler32_combine(a, b, len(data) - 700000) == zlib.adler32(data)


def test_capacity_too_small_reports_need_output(alice):
    L = _lib.lib()
    src = np.frombuffer(alice + b"\0" * 16, dtype=np.uint8)
    in_off = np.array([0, len(alice)], dtype=np.uint64)
    out = np.full(1000 + 16, 0xEE, dtype=np.uint8)
    out_off = np.array([0, 1000], dtype=np.uint64)
    lens = np.zeros(1, dtype=np.uint64)
    st = np.zeros(1, dtype=np.int32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = L.cz_deflate_batch(1, p(src), p(in_off), p(out), p(out_off), p(lens), p(st), 6, 15, 0, 0, 0)
    assert rc == 0 and st[0] == 1 and lens[0] > 1000
    assert (out[1000:] == 0xEE).all()


def test_device_api_segments(alice):
    torch = pytest.importorskip("torch")
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    segs = [alice[i * 50000:(i + 1) * 50000] for i in range(3)] + [b"", alice[:7]]
    n = len(segs)
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in segs])
    total = int(in_off[-1])
    d_in = torch.from_numpy(np.frombuffer(b"".join(segs) + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    caps = [int(L.cz_deflate_segment_bound(len(s))) for s in segs]
    out_off = np.zeros(n + 1, dtype=np.int64)
    out_off[1:] = np.cumsum(caps)
    d_out = torch.zeros(int(out_off[-1]) + 16, dtype=torch.uint8, device=dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out_off = torch.from_numpy(out_off).to(dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    d_chk = torch.zeros(2 * n, dtype=torch.int32, device=dev)
    ws = int(L.cz_deflate_workspace_bytes(n, total))
    d_ws = torch.zeros(ws, dtype=torch.uint8, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.cz_deflate_segments_device(sp, n, d_in.data_ptr(), d_in_off.data_ptr(), total, d_out.data_ptr(), d_out_off.data_ptr(),
                                      d_lens.data_ptr(), d_st.data_ptr(), d_chk.data_ptr(), 6, 0, d_ws.data_ptr(), ws)
    _lib.check(rc, "cz_deflate_segments_device")
    torch.cuda.synchronize()
    assert d_st.tolist() == [2] * n
    out = d_out.cpu().numpy()
    chk = d_chk.cpu().numpy().view(np.uint32)
    for i, s in enumerate(segs):
        piece = out[out_off[i]:out_off[i] + int(d_lens[i])].tobytes()
        assert piece[-4:] == b"\x00\x00\xff\xff"
        assert zlib.decompressobj(-15).decompress(piece) == s
        assert chk[2 * i] == zlib.adler32(s) and chk[2 * i + 1] == zlib.crc32(s)
    # and back through the segment-mode inflate on the device
    d_back = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    d_lens2 = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st2 = torch.zeros(n, dtype=torch.int32, device=dev)
    d_chk2 = torch.zeros(2 * n, dtype=torch.int32, device=dev)
    comp_off = np.zeros(n + 1, dtype=np.int64)
    # segments sit in their slots, so inflate them slot by slot: in_off = slot start .. slot start + len
    ws2 = int(L.cz_inflate_workspace_bytes(1, max(len(x) for x in segs)))
    d_ws2 = torch.zeros(ws2, dtype=torch.uint8, device=dev)
    for i in range(n):
        io = torch.tensor([out_off[i], out_off[i] + int(d_lens[i])], dtype=torch.int64, device=dev)
        oo = torch.tensor([in_off[i], in_off[i + 1]], dtype=torch.int64, device=dev)
        rc = L.cz_inflate_segments_device(sp, 1, d_out.data_ptr(), io.data_ptr(), d_back.data_ptr(), oo.data_ptr(), len(segs[i]),
                                          d_lens2[i:].data_ptr(), d_st2[i:].data_ptr(), d_chk2[2 * i:].data_ptr(), d_ws2.data_ptr(), ws2)
        _lib.check(rc, "cz_inflate_segments_device")
    torch.cuda.synchronize()
    assert d_st2.tolist() == [2] * n
    assert d_back[:total].cpu().numpy().tobytes() == b"".join(segs)
    assert (d_chk2.cpu().numpy().view(np.uint32) == chk).all()


def _synth_host(kind, nbytes, seed=4242):
    L = _lib.lib()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    import conftest
    corpus = np.frombuffer(conftest.read_golden("alice29.txt"), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    assert L.cz_synth_build_model(p(corpus), len(corpus), p(model)) == 0
    n = nbytes // 65536
    offs = np.arange(n + 1, dtype=np.uint64) * 65536
    buf = np.empty(n * 65536 + 16, dtype=np.uint8)
    assert L.cz_synth_fill_host(kind, seed, n, p(buf), p(offs), p(model)) == 0
    return buf, n * 65536


def test_full_size_gzip_1gib_roundtrip_with_parallel_crc_combine():
    """BASELINE.json configs[3] at full size: gzip encode + decode of 1 GiB mixed-entropy data; the CRC-32 in the trailer comes
    from the per-segment parallel combine and must equal zlib.crc32 of the whole input (checksum of checksums)."""
    L = _lib.lib()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    src, n = _synth_host(3, 1 << 30)
    cap = int(L.cz_deflate_bound(n, 31, 1 << 20))
    out = np.empty(cap + 16, dtype=np.uint8)
    out_len = ctypes.c_uint64(0)
    nseg = ctypes.c_uint64(0)
    idx = np.zeros(n // (1 << 20) + 8, dtype=np.uint64)
    rc = L.cz_deflate_segmented(p(src), n, p(out), cap, ctypes.byref(out_len), 6, 31, 0, 1 << 20, 0, p(idx), len(idx), ctypes.byref(nseg))
    _lib.check(rc, "cz_deflate_segmented")
    clen = out_len.value
    assert int.from_bytes(out[clen - 8:clen - 4].tobytes(), "little") == zlib.crc32(src[:n])
    assert int.from_bytes(out[clen - 4:clen].tobytes(), "little") == n & 0xffffffff
    # decode: segment-parallel with the side index (checks the combined CRC against the trailer on the way)
    back = np.empty(n + 16, dtype=np.uint8)
    got = ctypes.c_uint64(0)
    rc = L.cz_inflate_segmented(p(out), clen, p(back), n, ctypes.byref(got), 31, 1 << 20, p(idx), nseg.value, 0)
    _lib.check(rc, "cz_inflate_segmented")
    assert got.value == n and np.array_equal(back[:n], src[:n])
    # and the reference decoder's L0 accepts it as ONE gzip member (sampled: the first 64 MiB of output)
    d = zlib.decompressobj(31)
    head = d.decompress(out[:clen].tobytes(), 64 << 20)
    assert head == src[:64 << 20].tobytes()


def test_full_size_deflate_4gib_single_stream():
    """BASELINE.json configs[2] at full size: one 4 GiB buffer -> ONE valid zlib stream of full-flush segments. Properties:
    Adler-32 in the trailer equals zlib.adler32 of the whole input; segment-parallel inflate round-trips bit-exactly; >4 GiB-safe
    offsets (the compressed stream and the index are addressed with 64-bit offsets)."""
    L = _lib.lib()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    src, n = _synth_host(0, 1 << 32)
    assert n == 1 << 32
    cap = int(L.cz_deflate_bound(n, 15, 1 << 20))
    out = np.empty(cap + 16, dtype=np.uint8)
    out_len = ctypes.c_uint64(0)
    nseg = ctypes.c_uint64(0)
    idx = np.zeros(n // (1 << 20) + 8, dtype=np.uint64)
    rc = L.cz_deflate_segmented(p(src), n, p(out), cap, ctypes.byref(out_len), 6, 15, 0, 1 << 20, 0, p(idx), len(idx), ctypes.byref(nseg))
    _lib.check(rc, "cz_deflate_segmented")
    clen = out_len.value
    assert nseg.value == 4096
    adler = 1
    for o in range(0, n, 1 << 30):
        adler = zlib.adler32(src[o:o + (1 << 30)], adler)
    assert int.from_bytes(out[clen - 4:clen].tobytes(), "big") == adler
    ratio = n / clen
    assert ratio > 2.0, ratio
    back = np.empty(n + 16, dtype=np.uint8)
    got = ctypes.c_uint64(0)
    rc = L.cz_inflate_segmented(p(out), clen, p(back), n, ctypes.byref(got), 15, 1 << 20, p(idx), nseg.value, 0)
    _lib.check(rc, "cz_inflate_segmented")
    assert got.value == n
    for o in range(0, n, 1 << 28):
        assert np.array_equal(back[o:o + (1 << 28)], src[o:o + (1 << 28)])


def test_streaming_encoder_slices_match_one_shot(alice):
    """A long Process() sequence is compressed slice by slice (64 MiB) as the input arrives; the bytes must be exactly those of
    a single Finish call with the whole input (tests/encoder.rs:56-57: output independent of the caller's chunking).
    Driven through the raw C ABI with preallocated buffers (the Vec mirror reallocates per call, which is quadratic here)."""
    L = _lib.lib()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    data = np.frombuffer((alice * 1000)[:150_000_000], dtype=np.uint8)
    cap = int(L.cz_deflate_bound(len(data), 15, 0))

    def run(chunks):
        st = L.cz_encoder_new(6, 15, 8, 0)
        assert st, _lib.last_error()
        out = np.empty(cap, dtype=np.uint8)
        pos = 0
        for c in chunks:
            r = L.cz_encode(st, p(c) if len(c) else p(out), len(c), ctypes.c_void_p(out.ctypes.data + pos), cap - pos, 0)
            assert r.status == 0 and r.input_remain == 0
            pos += cap - pos - r.output_remain
        r = L.cz_encode(st, p(out), 0, ctypes.c_void_p(out.ctypes.data + pos), cap - pos, 2)
        assert r.status == 2
        pos += cap - pos - r.output_remain
        L.cz_encoder_free(st)
        return out[:pos]

    one = run([data])
    step = 7_000_003
    chunked = run([data[o:o + step] for o in range(0, len(data), step)])
    assert len(one) == len(chunked) and np.array_equal(one, chunked)
    assert zlib.decompress(one.tobytes()) == data.tobytes()


def test_cfg5_shaped_batch_roundtrip_mixed_sizes_and_classes():
    """BASELINE.json configs[4] in shape, 1 GiB on the devices present: stream sizes log-uniform in [4 KiB, 16 MiB], the three
    synthetic classes round robin, every unit deflated (zlib, level 6) and inflated again by the batched entry points with all
    devices. Size-independent properties: every unit finishes, the round trip is bit-exact, compressed bytes do not depend on
    the device count (checked when there are >= 2 devices), and a sample of the GPU-encoded streams decodes with zlib (the
    reference decoder's L0) while a sample of zlib-encoded units decodes on the GPU."""
    L = _lib.lib()
    rng = np.random.default_rng(5)
    total, sizes = 0, []
    while total < (1 << 30):
        s = int(np.exp(rng.uniform(np.log(4096), np.log(16 << 20))))
        sizes.append(s)
        total += s
    srcs = {k: _synth_host(k, (max(sizes) + 65535) // 65536 * 65536 * 2, seed=77 + k)[0] for k in (0, 1, 2)}
    units = []
    for i, s in enumerate(sizes):
        src = srcs[i % 3]
        o = int(rng.integers(0, len(src) - 16 - s))
        units.append(src[o:o + s].tobytes())
    mask = (1 << min(L.cz_device_count(), 8)) - 1
    comp, st = batch.deflate_batch(units, level=6, window_bits=15, devices_mask=mask)
    assert (st == 2).all()
    if mask != 1:
        comp1, st1 = batch.deflate_batch(units, level=6, window_bits=15, devices_mask=1)
        assert comp1 == comp
    outs, ist, lens, cons = batch.inflate_batch(comp, [len(u) for u in units], 15, devices_mask=mask)
    assert (ist == 2).all()
    assert outs == units
    assert list(cons) == [len(c) for c in comp]
    pick = sorted(set([0, len(units) - 1, int(np.argmax(sizes)), int(np.argmin(sizes))] + [int(x) for x in rng.integers(0, len(units), 8)]))
    for i in pick:
        assert zlib.decompress(comp[i]) == units[i]
    zs = [zlib.compress(units[i], 6) for i in pick]
    outs2, st2, _, _ = batch.inflate_batch(zs, [len(units[i]) for i in pick], 15, devices_mask=mask)
    assert (st2 == 2).all() and outs2 == [units[i] for i in pick]
    ratio = sum(len(u) for u in units) / sum(len(c) for c in comp)
    assert 1.2 < ratio < 4.0, ratio
