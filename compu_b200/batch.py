"""Batched many-stream entry points (NEW relative to the reference; BASELINE.json north_star) over host numpy buffers."""
import ctypes

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pack(chunks):
    """list of bytes -> (uint8 buffer with 16 B slack, uint64 offsets[n+1])"""
    offs = np.zeros(len(chunks) + 1, dtype=np.uint64)
    if chunks:
        offs[1:] = np.cumsum([len(c) for c in chunks], dtype=np.uint64)
    buf = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy()
    return buf, offs


def inflate_batch(streams, caps, window_bits=47, devices_mask=0):
    """Inflates independent streams. Returns (outputs: list[bytes], statuses: int32[n], out_lens, in_consumed)."""
    _lib.require_device()
    n = len(streams)
    inbuf, in_off = pack(streams)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        out_off[1:] = np.cumsum(np.asarray(caps, dtype=np.uint64))
    out = np.empty(int(out_off[-1]) + 16, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint64)
    statuses = np.full(n, -99, dtype=np.int32)
    consumed = np.zeros(n, dtype=np.uint64)
    rc = _lib.lib().cz_inflate_batch(n, _p(inbuf), _p(in_off), _p(out), _p(out_off), _p(out_lens), _p(statuses), _p(consumed),
                                     int(window_bits), devices_mask)
    _lib.check(rc, "cz_inflate_batch")
    outs = [out[int(out_off[i]):int(out_off[i]) + int(out_lens[i])].tobytes() for i in range(n)]
    return outs, statuses, out_lens, consumed


def deflate_batch(buffers, level=6, window_bits=15, strategy=0, segment_bytes=0, devices_mask=0):
    """Deflates independent buffers, each into its own complete stream. Returns (streams: list[bytes], statuses)."""
    _lib.require_device()
    L = _lib.lib()
    n = len(buffers)
    inbuf, in_off = pack(buffers)
    caps = [int(L.cz_deflate_bound(len(b), window_bits, segment_bytes)) for b in buffers]
    out_off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        out_off[1:] = np.cumsum(np.asarray(caps, dtype=np.uint64))
    out = np.empty(int(out_off[-1]) + 16, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint64)
    statuses = np.full(n, -99, dtype=np.int32)
    rc = L.cz_deflate_batch(n, _p(inbuf), _p(in_off), _p(out), _p(out_off), _p(out_lens), _p(statuses), level, window_bits,
                            strategy, segment_bytes, devices_mask)
    _lib.check(rc, "cz_deflate_batch")
    outs = [out[int(out_off[i]):int(out_off[i]) + int(out_lens[i])].tobytes() for i in range(n)]
    return outs, statuses


def deflate_segmented(data, level=6, window_bits=15, strategy=0, segment_bytes=0, devices_mask=0):
    """One buffer -> one valid stream of full-flush segments. Returns (stream: bytes, seg_index: uint64[n_seg+1])."""
    _lib.require_device()
    L = _lib.lib()
    src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
    n = len(data)
    cap = int(L.cz_deflate_bound(n, window_bits, segment_bytes))
    out = np.empty(cap + 16, dtype=np.uint8)
    out_len = ctypes.c_uint64(0)
    nseg = ctypes.c_uint64(0)
    seg = int(segment_bytes) if segment_bytes else (1 << 20)
    seg = min(max(seg, 4096), int(L.cz_deflate_max_segment()))  # the library clamps the same way
    idx_cap = (n + seg - 1) // seg + 2
    idx = np.zeros(idx_cap, dtype=np.uint64)
    rc = L.cz_deflate_segmented(_p(src), n, _p(out), cap, ctypes.byref(out_len), level, window_bits, strategy, segment_bytes,
                                devices_mask, _p(idx), idx_cap, ctypes.byref(nseg))
    _lib.check(rc, "cz_deflate_segmented")
    return out[:out_len.value].tobytes(), idx[:nseg.value + 1].copy()


def inflate_segmented(stream, out_len, seg_index, window_bits=15, segment_bytes=0, devices_mask=0):
    """Segment-parallel inflate of a stream produced by deflate_segmented, using its side index."""
    _lib.require_device()
    L = _lib.lib()
    src = np.frombuffer(bytes(stream) + b"\0" * 16, dtype=np.uint8)
    out = np.empty(out_len + 16, dtype=np.uint8)
    got = ctypes.c_uint64(0)
    idx = np.ascontiguousarray(seg_index, dtype=np.uint64)
    rc = L.cz_inflate_segmented(_p(src), len(stream), _p(out), out_len, ctypes.byref(got), window_bits, segment_bytes, _p(idx),
                                len(idx) - 1, devices_mask)
    _lib.check(rc, "cz_inflate_segmented")
    return out[:got.value].tobytes()


def partition_by_bytes(offsets, parts):
    """Contiguous, byte-balanced shard cuts (the host-side partitioner of the batched entry points). No device needed."""
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    cuts = np.zeros(parts + 1, dtype=np.uint64)
    rc = _lib.lib().cz_partition_by_bytes(len(offs) - 1, _p(offs), parts, _p(cuts))
    _lib.check(rc, "cz_partition_by_bytes")
    return [int(c) for c in cuts]
