"""The C-ABI library loads and exports every symbol include/compu_b200.h declares; without a GPU the product fails
loudly (no CPU path). No compute calls here."""
import os
import re

import pytest

from compu_b200 import _lib
from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "compu_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cz_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 30
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), "libcompu_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "%s has no ctypes signature" % s
    assert set(_lib.SIGNATURES) == set(syms)


def test_describe_error_table():
    L = _lib.lib()
    assert L.cz_describe_error(0) is not None          # tests/decoder.rs:74-75: no_error() must describe
    assert L.cz_describe_error(-3) == b"data error"
    assert L.cz_describe_error(-5) == b"buffer error"
    assert L.cz_describe_error(2) == b"need dictionary"


def test_fails_loudly_without_device():
    L = _lib.lib()
    if L.cz_device_count() > 0:
        pytest.skip("a GPU is present")
    from compu_b200 import batch, decoder, encoder
    assert decoder.Interface.zlib_cuda() is None       # ctor -> None, like zlib_ng on init failure
    assert encoder.Interface.zlib_cuda() is None
    with pytest.raises(RuntimeError):
        batch.inflate_batch([b"x"], [10], 15)
    with pytest.raises(RuntimeError):
        batch.deflate_batch([b"x"])


def test_product_does_not_touch_the_oracle():
    # the product package must never import, link or call anything under oracle/
    pkg = os.path.join(ROOT, "compu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "compu_oracle" not in txt and "oz_" not in txt, f
    import subprocess
    out = subprocess.run(["ldd", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "libz" not in out and "oracle" not in out
