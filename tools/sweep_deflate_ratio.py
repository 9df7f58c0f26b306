#!/usr/bin/env python
"""Ratio / work sweep of the level-6 match search on the CPU (no GPU needed): compiles the encoder's shared decision code
(tests/model, the SAME deflate_core.cuh the kernels use) with (max_chain, nice_len) overridden and compresses 1 MiB of each
synthetic class plus alice29.txt as ONE segment; prints the size relative to zlib 1.3 level 6 on the same bytes and the chain
steps walked per input byte (the match search's work). Picks feed deflate_tuning() (compu_b200/csrc/deflate_core.cuh)."""
import ctypes
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "model")


def build(chain, nice, lazy=None):
    so = "/tmp/czk_model_%d_%d.so" % (chain, nice)
    flags = ["-DCZK_SWEEP_CHAIN=%d" % chain, "-DCZK_SWEEP_NICE=%d" % nice]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DCZK_MODEL"] + flags + ["-o", so, os.path.join(MODEL, "deflate_model.cpp")])
    L = ctypes.CDLL(so)
    L.model_deflate_segment.restype = ctypes.c_long
    return L


def run(L, data, level=6):
    src = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    out = np.zeros(len(data) + len(data) // 8 + 1024, dtype=np.uint8)
    stats = np.zeros(8, dtype=np.uint64)
    r = L.model_deflate_segment(src.ctypes.data_as(ctypes.c_void_p), len(data), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(out)),
                                level, 0, None, None, stats.ctypes.data_as(ctypes.c_void_p))
    assert r > 0
    assert zlib.decompress(out[:r].tobytes() + b"\x03\x00", -15) == data
    return r, int(stats[7])


def main():
    alice = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read()
    datas = {"markov": bench.oracle_synth(0, 1 << 20, 5)[:1 << 20].tobytes(), "repeat": bench.oracle_synth(1, 1 << 20, 5)[:1 << 20].tobytes(),
             "random": bench.oracle_synth(2, 1 << 20, 5)[:1 << 20].tobytes(), "alice29": alice}
    ref = {k: len(zlib.compress(v, 6)) - 6 for k, v in datas.items()}
    print("| chain | nice | " + " | ".join("%s size/zlib, steps/B" % k for k in datas) + " |")
    print("|---|---|" + "---|" * len(datas))
    for chain, nice in [(16, 64), (12, 64), (8, 64), (8, 32), (6, 32), (6, 16), (4, 32), (4, 16), (3, 16), (2, 8)]:
        L = build(chain, nice)
        cells = []
        for k, v in datas.items():
            size, steps = run(L, v)
            cells.append("%.4f, %.2f" % (size / ref[k], steps / len(v)))
        print("| %d | %d | " % (chain, nice) + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
