#!/usr/bin/env python
"""Device-resident timing of the deflate kernel chain (development tool; bench.py carries the headline numbers).

usage: bench_deflate.py [--mib 1024] [--seg-kib 1024] [--kind 0] [--level 6] [--reps 3] [--check]
Generates synthetic data on the device, compresses it as independent full-flush segments (cz_deflate_segments_device),
reports uncompressed GB/s, ratio, and (with --check) verifies a sample of segments against zlib on the host.
"""
import argparse
import ctypes
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--seg-kib", type=int, default=1024)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import torch
    from compu_b200 import _lib
    L = _lib.lib()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    seg = a.seg_kib * 1024
    total = a.mib << 20
    n = total // seg
    p = lambda x: ctypes.c_void_p(x.ctypes.data)
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    _lib.check(L.cz_synth_build_model(p(corpus), len(corpus), p(model)), "model")
    d_model = torch.from_numpy(model).to(dev)
    d_in = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    d_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * seg
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # generator units of 64 KiB so that the data do not depend on the segment size
    gu = total // 65536
    d_goff = torch.arange(gu + 1, dtype=torch.int64, device=dev) * 65536
    _lib.check(L.cz_synth_fill_device(sp, a.kind, 1234, gu, d_in.data_ptr(), d_goff.data_ptr(), d_model.data_ptr()), "synth")
    bound = int(L.cz_deflate_segment_bound(seg))
    d_out = torch.empty(n * bound + 16, dtype=torch.uint8, device=dev)
    d_out_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * bound
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    d_chk = torch.zeros(2 * n, dtype=torch.int32, device=dev)
    ws = int(L.cz_deflate_workspace_bytes(n, total))
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)

    def step():
        rc = L.cz_deflate_segments_device(sp, n, d_in.data_ptr(), d_off.data_ptr(), total, d_out.data_ptr(), d_out_off.data_ptr(),
                                          d_lens.data_ptr(), d_st.data_ptr(), d_chk.data_ptr(), a.level, 0, d_ws.data_ptr(), ws)
        _lib.check(rc, "cz_deflate_segments_device")

    step()
    torch.cuda.synchronize()
    assert bool((d_st == 2).all())
    C = int(d_lens.sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    res = {"mib": a.mib, "seg_kib": a.seg_kib, "kind": a.kind, "level": a.level, "segments": n, "ms": ms,
           "GBps_uncompressed": total / ms / 1e6, "ratio": total / C, "workspace_gib": ws / 2**30,
           "hbm_frac": (total + C) / ms / 1e6 / 6554.2}
    if a.check:
        lens = d_lens.cpu().numpy()
        out = d_out.cpu().numpy()
        src = d_in.cpu().numpy()
        zsum = 0
        for i in list(range(0, n, max(1, n // 16)))[:16]:
            piece = out[i * bound:i * bound + int(lens[i])].tobytes()
            plain = src[i * seg:(i + 1) * seg].tobytes()
            assert zlib.decompressobj(-15).decompress(piece) == plain
            zsum += len(zlib.compress(plain, a.level))
        res["checked"] = True
        ours = sum(int(lens[i]) for i in list(range(0, n, max(1, n // 16)))[:16])
        res["size_vs_zlib_same_level"] = ours / zsum
    print(json.dumps(res))


if __name__ == "__main__":
    main()
