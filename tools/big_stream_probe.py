#!/usr/bin/env python
"""Times the warp-per-stream inflate kernel (cz_tune_inflate(1, 8)) on a few LONG zlib-made streams, device-resident —
development tool. usage: big_stream_probe.py [n_streams] [MiB per stream]"""
import ctypes
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from compu_b200 import _lib
    n, mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 8
    L = _lib.lib()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    SB = mib << 20
    plain = bench.host_synth_bytes(n * SB, 99)
    t0 = time.perf_counter()
    streams = [zlib.compress(plain[i * SB:(i + 1) * SB].tobytes(), 6) for i in range(n)]
    t0 = time.perf_counter()
    for s in streams:
        zlib.decompress(s)
    t_cpu = time.perf_counter() - t0
    U = n * SB
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    d_in = torch.from_numpy(np.frombuffer(b"".join(streams) + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * SB
    d_out = torch.empty(U + 16, dtype=torch.uint8, device=dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    ws_bytes = int(L.cz_inflate_workspace_bytes(n, U))
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.cz_tune_inflate(1, 8)

    def step():
        _lib.check(L.cz_inflate_batch_device(sp, n, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(), U,
                                             d_lens.data_ptr(), d_stat.data_ptr(), None, 15, d_ws.data_ptr(), ws_bytes), "inflate")
    step()
    torch.cuda.synchronize()
    ok = bool((d_stat == 2).all()) and d_out[:U].cpu().numpy().tobytes() == plain[:U].tobytes()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%d streams x %d MiB, warp per stream: ok=%s  %.1f ms  -> %.1f MB/s per stream, %.2f GB/s aggregate (python zlib on one core: "
          "%.1f MB/s per stream)" % (n, mib, ok, ms, SB / ms / 1e3, U / ms / 1e6, U / t_cpu / 1e6))


if __name__ == "__main__":
    main()
