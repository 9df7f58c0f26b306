#!/usr/bin/env python
"""LONG zlib-made streams (no flush points) through the ordinary host entry point cz_inflate_batch — the block-parallel path of
compu_b200/csrc/inflate_runs.cuh — timed end to end (pinned host buffers, H2D + D2H inside), one stream alone and a batch of them,
with CZ_NO_RUNS=1 (serial warp-per-stream kernel) beside it. Development tool.
usage: big_stream_probe.py [n_streams] [MiB per stream] [level]"""
import ctypes
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from compu_b200 import _lib
    n, mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 8
    level = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    L = _lib.lib()
    _lib.require_device()
    SB = mib << 20
    plain = bench.oracle_synth(0, n * SB, 99)
    with ThreadPoolExecutor(max_workers=16) as ex:
        streams = list(ex.map(lambda i: zlib.compress(plain[i * SB:(i + 1) * SB].tobytes(), level), range(n)))
    t0 = time.perf_counter()
    zlib.decompress(streams[0])
    t_cpu = time.perf_counter() - t0
    p = lambda a: ctypes.c_void_p(a.ctypes.data)

    def run(k):
        C = sum(len(s) for s in streams[:k])
        hi = L.cz_host_alloc(C + 16)
        ho = L.cz_host_alloc(k * SB + 16)
        h_in = np.ctypeslib.as_array(ctypes.cast(hi, ctypes.POINTER(ctypes.c_uint8)), shape=(C + 16,))
        h_out = np.ctypeslib.as_array(ctypes.cast(ho, ctypes.POINTER(ctypes.c_uint8)), shape=(k * SB + 16,))
        in_off = np.zeros(k + 1, dtype=np.uint64)
        in_off[1:] = np.cumsum([len(s) for s in streams[:k]])
        for i in range(k):
            h_in[int(in_off[i]):int(in_off[i + 1])] = np.frombuffer(streams[i], dtype=np.uint8)
        out_off = np.arange(k + 1, dtype=np.uint64) * SB
        lens, st = np.zeros(k, dtype=np.uint64), np.zeros(k, dtype=np.int32)
        best = None
        for it in range(4):
            h_out[:] = 0
            t0 = time.perf_counter()
            _lib.check(L.cz_inflate_batch(k, ctypes.c_void_p(hi), p(in_off), ctypes.c_void_p(ho), p(out_off), p(lens), p(st), None, 15, 1), "inflate")
            dt = time.perf_counter() - t0
            if it:
                best = dt if best is None else min(best, dt)
        ok = bool((st == 2).all()) and (h_out[:k * SB] == plain[:k * SB]).all()
        L.cz_host_free(ctypes.c_void_p(hi)); L.cz_host_free(ctypes.c_void_p(ho))
        return ok, best

    a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
    ok1, t1 = run(1)
    okn, tn = run(n)
    L.cz_split_stats(ctypes.byref(a), ctypes.byref(b))
    print("zlib level %d streams of %d MiB (no flush points), cz_inflate_batch end to end, mode %s: ONE stream ok=%s %.2f ms = %.2f GB/s per stream; "
          "%d streams ok=%s %.1f ms = %.2f GB/s aggregate; long units seen / decoded in parallel: %d / %d; python zlib on one core: %.2f GB/s"
          % (level, mib, "serial (CZ_NO_RUNS)" if os.environ.get("CZ_NO_RUNS") else "block-parallel", ok1, t1 * 1e3, SB / t1 / 1e9, n, okn, tn * 1e3,
             n * SB / tn / 1e9, a.value, b.value, SB / t_cpu / 1e9))


if __name__ == "__main__":
    main()
