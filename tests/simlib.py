"""ctypes access to the CPU-emulated kernels (tests/cusim) — test infrastructure only."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "cusim")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s"])
        _lib = ctypes.CDLL(os.path.join(_DIR, "libcusim_kernels.so"))
        _lib.sim_crc32_combine.restype = ctypes.c_uint32
        _lib.sim_crc32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
        _lib.sim_adler32_combine.restype = ctypes.c_uint32
        _lib.sim_adler32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    return _lib


def pack(chunks):
    offs = np.zeros(len(chunks) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in chunks])
    buf = np.frombuffer(b"".join(chunks) + b"\0" * 8, dtype=np.uint8).copy()
    return buf, offs


def sim_inflate(streams, caps, window_bits, segment_mode=0, check_kind=0, D=4, grid=2, seed=1):
    """Returns (outputs, statuses, out_lens, in_consumed, checks)."""
    L = lib()
    n = len(streams)
    inbuf, in_off = pack(streams)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)
    out = np.full(int(out_off[-1]) + 64, 0xEE, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint64)
    statuses = np.full(n, -99, dtype=np.int32)
    consumed = np.zeros(n, dtype=np.uint64)
    checks = np.zeros(2 * n, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    r = L.sim_inflate(ctypes.c_size_t(n), p(inbuf), p(in_off), p(out), p(out_off), p(out_lens), p(statuses), p(consumed),
                      p(checks), window_bits, segment_mode, check_kind, D, grid, ctypes.c_uint64(seed))
    assert r == 0
    outs = [bytes(out[int(out_off[i]):int(out_off[i]) + int(out_lens[i])]) for i in range(n)]
    # nothing may be written past a slot's capacity
    assert (out[int(out_off[-1]):] == 0xEE).all()
    return outs, statuses, out_lens, consumed, checks


def sim_deflate(units, seg_bytes=65536, level=6, strategy=0, window_bits=15, piece_mode=0, packed=0, caps=None, seed=1):
    """Runs the encoder kernel chain on the emulator. Returns (streams, statuses, out_lens, checks, seg_sizes)."""
    L = lib()
    n = len(units)
    inbuf, unit_off = pack(units)
    seg_off, unit_seg = [], [0]
    for u in range(n):
        a, b = int(unit_off[u]), int(unit_off[u + 1])
        while True:
            seg_off.append(a)
            a += min(seg_bytes, b - a)
            if a >= b:
                break
        unit_seg.append(len(seg_off))
    seg_off.append(int(unit_off[n]))
    nseg = len(seg_off) - 1
    seg_off = np.asarray(seg_off, dtype=np.uint64)
    unit_seg = np.asarray(unit_seg, dtype=np.uint32)
    if caps is None:
        caps = [len(u) + len(u) // 2048 + 64 * (len(u) // seg_bytes + 1) + 32 for u in units]
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(np.asarray(caps, dtype=np.uint64))
    out = np.full(int(out_off[-1]) + 64, 0xEE, dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.full(n, -99, dtype=np.int32)
    checks = np.zeros(2 * n, dtype=np.uint32)
    seg_sizes = np.zeros(nseg, dtype=np.uint64)
    pos = np.zeros(n, dtype=np.uint64)
    total = np.zeros(1, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    r = L.sim_deflate(ctypes.c_size_t(nseg), ctypes.c_size_t(n), p(inbuf), p(seg_off), p(unit_seg), p(out), p(out_off), p(out_len),
                      p(status), p(checks), p(seg_sizes), level, strategy, window_bits, piece_mode, packed, p(pos), p(total),
                      ctypes.c_uint64(seed))
    assert r == 0
    assert (out[int(out_off[-1]):] == 0xEE).all()
    streams = []
    for i in range(n):
        o = int(pos[i]) if packed else int(out_off[i])
        streams.append(bytes(out[o:o + int(out_len[i])]) if status[i] == 2 else b"")
    return streams, status, out_len, checks, seg_sizes


def sim_inflate_runs(streams, caps, window_bits, chunk_bytes=4096, grid=3, seed=1):
    """The block-parallel path for long streams (inflate_runs.cuh + inflate_runs_host.h) on the emulator.
    Returns (outputs, ok[], out_lens, consumed, n_runs)."""
    L = lib()
    n = len(streams)
    inbuf, in_off = pack(streams)
    inbuf = np.concatenate([inbuf, np.full(64, 0xA5, dtype=np.uint8)])
    out_off = np.zeros(n + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)
    out = np.full(int(out_off[-1]) + 64, 0xEE, dtype=np.uint8)
    out_lens = np.zeros(n, dtype=np.uint64)
    consumed = np.zeros(n, dtype=np.uint64)
    ok = np.zeros(n, dtype=np.uint8)
    nruns = np.zeros(n, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    r = L.sim_inflate_runs(ctypes.c_size_t(n), p(inbuf), p(in_off), p(out), p(out_off), p(out_lens), p(consumed), p(ok), p(nruns),
                           window_bits, ctypes.c_uint64(chunk_bytes), grid, ctypes.c_uint64(seed))
    assert r == 0
    assert (out[int(out_off[-1]):] == 0xEE).all()
    outs = [bytes(out[int(out_off[i]):int(out_off[i]) + int(out_lens[i])]) if ok[i] else None for i in range(n)]
    return outs, ok, out_lens, consumed, nruns
