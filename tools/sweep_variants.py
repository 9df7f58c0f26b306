#!/usr/bin/env python
"""Times the device-resident cfg2 inflate (tok + lz kernel times from the library's own CUDA events) for several BUILDS of the
library in one process: the data set is made once, every library given on the command line (default: the product library and
variants/lib_*.so, see tools/build_variant.sh) is loaded side by side. Development tool; one JSON line per library."""
import argparse
import ctypes
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def load(path):
    from compu_b200 import _lib
    L = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--env", default="", help="NAME=v1,v2,...: additionally run every library once per value (read at first use, so one process per value is NOT needed only for knobs read per call)")
    args = ap.parse_args()
    import torch
    from compu_b200 import _lib
    L0 = _lib.lib()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    n, SB = args.streams, bench.STREAM_BYTES
    U = n * SB
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    corpus = np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read(), dtype=np.uint8)
    model = np.zeros(int(L0.cz_synth_model_bytes()), dtype=np.uint8)
    L0.cz_synth_build_model(bench._p(corpus), len(corpus), bench._p(model))
    d_model = torch.from_numpy(model).to(dev)
    d_plain = torch.empty(U, dtype=torch.uint8, device=dev)
    d_out_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * SB
    L0.cz_synth_fill_device(sp, args.kind, 777, n, d_plain.data_ptr(), d_out_off.data_ptr(), d_model.data_ptr())
    torch.cuda.synchronize()
    plain = d_plain.cpu().numpy()
    streams = bench.compress_streams(plain, n, os.cpu_count() or 1)
    C = sum(len(s) for s in streams)
    in_off = np.zeros(n + 1, dtype=np.int64)
    in_off[1:] = np.cumsum([len(s) for s in streams])
    d_in = torch.from_numpy(np.frombuffer(b"".join(streams) + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d_in_off = torch.from_numpy(in_off).to(dev)
    d_out = torch.empty(U + 16, dtype=torch.uint8, device=dev)
    d_lens = torch.zeros(n, dtype=torch.int64, device=dev)
    d_stat = torch.zeros(n, dtype=torch.int32, device=dev)
    ws_bytes = int(L0.cz_inflate_workspace_bytes(n, U))
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    print(json.dumps({"streams": n, "kind": args.kind, "ratio": U / C}))
    libs = args.libs or [_lib.SO_PATH] + sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))
    for path in libs:
        L = load(path)
        if L.cz_device_count() < 1:
            print(json.dumps({"lib": path, "error": "no device"}))
            continue

        def step():
            rc = L.cz_inflate_batch_device(sp, n, d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(), d_out_off.data_ptr(),
                                           U, d_lens.data_ptr(), d_stat.data_ptr(), None, 15, d_ws.data_ptr(), ws_bytes)
            if rc != 0:
                raise RuntimeError("%s: inflate failed %d: %s" % (path, rc, L.cz_last_error().decode()))
        d_out.zero_()
        step()
        torch.cuda.synchronize()
        ok = bool((d_stat == 2).all()) and torch.equal(d_out[:U], d_plain)
        for _ in range(2):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        L.cz_profile_enable(1)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        L.cz_profile_enable(0)
        ka, kb = ctypes.c_double(0), ctypes.c_double(0)
        nprof = L.cz_profile_read(ctypes.byref(ka), ctypes.byref(kb))
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"lib": os.path.basename(path), "ok": ok, "ms": round(ms, 3), "tok_ms": round(ka.value / max(1, nprof), 3),
                          "lz_ms": round(kb.value / max(1, nprof), 3), "GBps": round(U / ms / 1e6, 1),
                          "hbm_frac": round((U + C) / ms / 1e6 / 6554.2, 4)}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
