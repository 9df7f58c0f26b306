"""CPU oracle loader — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. It wraps oracle/libcompu_oracle.so (compu's zlib glue restated in C over madler zlib 1.3, see
compu_oracle.c for the reference file:line map). The product package compu_b200 never imports it.
"""
import ctypes, os, subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcompu_oracle.so")


class Result(ctypes.Structure):
    _fields_ = [("input_remain", ctypes.c_size_t), ("output_remain", ctypes.c_size_t), ("status", ctypes.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "compu_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcompu_oracle.so"])
    return _SO


_synth = None


def synth():
    """oracle/libcompu_synth.so: the synthetic-data generator alone (host build of compu_b200/csrc/synth.cuh, no codec)."""
    global _synth
    if _synth is None:
        so = os.path.join(_HERE, "libcompu_synth.so")
        src = os.path.join(_HERE, "synth_host.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcompu_synth.so"])
        S = ctypes.CDLL(so)
        S.oz_synth_model_bytes.restype = ctypes.c_uint64
        S.oz_synth_build_model.restype = ctypes.c_int
        S.oz_synth_build_model.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        S.oz_synth_fill.restype = ctypes.c_int
        S.oz_synth_fill.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _synth = S
    return _synth


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = ctypes.CDLL(_SO)
    vp, sz, i32, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_void_p
    L.oz_decoder_new.restype = vp; L.oz_decoder_new.argtypes = [ctypes.c_int]
    L.oz_decode.restype = Result; L.oz_decode.argtypes = [vp, u8p, sz, u8p, sz]
    L.oz_decoder_reset.restype = vp; L.oz_decoder_reset.argtypes = [vp]
    L.oz_decoder_free.restype = None; L.oz_decoder_free.argtypes = [vp]
    L.oz_describe_error.restype = ctypes.c_char_p; L.oz_describe_error.argtypes = [i32]
    L.oz_encoder_new.restype = vp; L.oz_encoder_new.argtypes = [ctypes.c_int] * 4
    L.oz_encode.restype = Result; L.oz_encode.argtypes = [vp, u8p, sz, u8p, sz, ctypes.c_int]
    L.oz_encoder_reset.restype = vp; L.oz_encoder_reset.argtypes = [vp]
    L.oz_encoder_free.restype = None; L.oz_encoder_free.argtypes = [vp]
    L.oz_crc32.restype = ctypes.c_uint32; L.oz_crc32.argtypes = [ctypes.c_uint32, u8p, sz]
    L.oz_adler32.restype = ctypes.c_uint32; L.oz_adler32.argtypes = [ctypes.c_uint32, u8p, sz]
    L.oz_crc32_combine.restype = ctypes.c_uint32; L.oz_crc32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    L.oz_adler32_combine.restype = ctypes.c_uint32; L.oz_adler32_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    L.oz_max_threads.restype = ctypes.c_int
    L.oz_inflate_batch.restype = ctypes.c_long
    L.oz_inflate_batch.argtypes = [sz, u8p, vp, u8p, vp, vp, vp, ctypes.c_int, ctypes.c_int]
    L.oz_deflate_batch.restype = ctypes.c_long
    L.oz_deflate_batch.argtypes = [sz, u8p, vp, u8p, vp, vp, vp] + [ctypes.c_int] * 5
    L.oz_deflate_bound.restype = ctypes.c_uint64; L.oz_deflate_bound.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
    L.oz_zlib_version.restype = ctypes.c_char_p
    _lib = L
    return L
