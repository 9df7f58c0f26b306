#!/usr/bin/env python
"""Times the device-resident batched inflate for several (slots-per-warp, warps-per-CTA) kernel configurations on one
data set (development tool; prints one JSON line per configuration)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--cfgs", default="1,8;2,8;4,7;8,7;16,3;32,1")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--lz", default="", help="phase B variants to time with the default kernel: 'cta_mode,spin_ns;...'")
    args = ap.parse_args()
    import torch
    from compu_b200 import _lib
    L = _lib.lib()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    n, SB = args.streams, bench.STREAM_BYTES
    U = n * SB
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    L.cz_synth_build_model(bench._p(corpus), len(corpus), bench._p(model))
    d_model = torch.from_numpy(model).to(dev)
    d_plain = torch.empty(U, dtype=torch.uint8, device=dev)
    d_out_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * SB
    L.cz_synth_fill_device(sp, args.kind, 777, n, d_plain.data_ptr(), d_out_off.data_ptr(), d_model.data_ptr())
    torch.cuda.synchronize()
    plain = d_plain.cpu().numpy()
    t0 = time.perf_counter()
    streams = bench.compress_streams(plain, n, os.cpu_count() or 1)
    t_comp = time.perf_counter() - t0
    C = sum(len(s) for s in streams)
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    d_in = torch.from_numpy(np.frombuffer(b"".join(streams) + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out = torch.empty(U + 16, dtype=torch.uint8, device=dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    ws_bytes = int(L.cz_inflate_workspace_bytes(n, U))
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    print(json.dumps({"streams": n, "kind": args.kind, "ratio": U / C, "host_compress_s": t_comp, "cores": os.cpu_count()}))
    runs = [(cfg, None) for cfg in args.cfgs.split(";") if cfg] + [("-2,14", lz) for lz in args.lz.split(";") if lz]
    for cfg, lz in runs:
        D, W = [int(x) for x in cfg.split(",")]
        L.cz_tune_inflate(D, W)
        if lz:
            L.cz_tune_inflate_lz(*[int(x) for x in lz.split(",")])
            cfg = cfg + " lz " + lz

        def step():
            rc = L.cz_inflate_batch_device(sp, n, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                           U, d_lens.data_ptr(), d_stat.data_ptr(), None, 15, d_ws.data_ptr(), ws_bytes)
            _lib.check(rc, "inflate")
        d_out.zero_()
        step()
        torch.cuda.synchronize()
        ok = bool((d_stat == 2).all()) and torch.equal(d_out[:U], d_plain)
        for _ in range(2):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"cfg": cfg, "ok": ok, "ms": ms, "GBps_uncompressed": U / ms / 1e6,
                          "hbm_frac": (U + C) / ms / 1e6 / 6554.2}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
