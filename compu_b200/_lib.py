"""ctypes binding of compu_b200/libcompu_b200.so (the C ABI declared in include/compu_b200.h).

There is no CPU path: if the CUDA library is missing this module raises at load, and if no sm_100 device is usable
every constructor / batch call fails loudly (RuntimeError), never silently.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcompu_b200.so")


class CzResult(ctypes.Structure):
    _fields_ = [("input_remain", ctypes.c_size_t), ("output_remain", ctypes.c_size_t), ("status", ctypes.c_int32)]


_lib = None

vp, sz, i32, u32, u64, ci = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int)

# name -> (restype, argtypes); every symbol include/compu_b200.h declares
SIGNATURES = {
    "cz_device_count": (ci, []),
    "cz_version": (ctypes.c_char_p, []),
    "cz_last_error": (ctypes.c_char_p, []),
    "cz_launch_count": (u64, []),
    "cz_host_alloc": (vp, [sz]),
    "cz_host_free": (None, [vp]),
    "cz_set_stream_device": (ci, [ci]),
    "cz_decoder_new": (vp, [ci]),
    "cz_decode": (CzResult, [vp, vp, sz, vp, sz]),
    "cz_decoder_reset": (vp, [vp]),
    "cz_decoder_free": (None, [vp]),
    "cz_describe_error": (ctypes.c_char_p, [i32]),
    "cz_encoder_new": (vp, [ci, ci, ci, ci]),
    "cz_encode": (CzResult, [vp, vp, sz, vp, sz, ci]),
    "cz_encoder_reset": (vp, [vp]),
    "cz_encoder_free": (None, [vp]),
    "cz_inflate_batch": (ci, [sz, vp, vp, vp, vp, vp, vp, vp, ci, u32]),
    "cz_inflate_batch_ptrs": (ci, [sz, vp, vp, vp, vp, vp, vp, ci, u32]),
    "cz_deflate_bound": (u64, [u64, ci, u64]),
    "cz_deflate_batch": (ci, [sz, vp, vp, vp, vp, vp, vp, ci, ci, ci, u64, u32]),
    "cz_deflate_segmented": (ci, [vp, u64, vp, u64, vp, ci, ci, ci, u64, u32, vp, u64, vp]),
    "cz_inflate_segmented": (ci, [vp, u64, vp, u64, vp, ci, u64, vp, u64, u32]),
    "cz_has_experiments": (ci, []),
    "cz_tune_inflate": (ci, [ci, ci]),
    "cz_tune_inflate_lz": (ci, [ci, ci]),
    "cz_inflate_workspace_bytes": (u64, [sz, u64]),
    "cz_inflate_batch_device": (ci, [vp, sz, vp, vp, vp, vp, u64, vp, vp, vp, ci, vp, u64]),
    "cz_inflate_segments_device": (ci, [vp, sz, vp, vp, vp, vp, u64, vp, vp, vp, vp, u64]),
    "cz_deflate_max_segment": (u64, []),
    "cz_deflate_segment_bound": (u64, [u64]),
    "cz_deflate_workspace_bytes": (u64, [sz, u64]),
    "cz_deflate_segments_device": (ci, [vp, sz, vp, vp, u64, vp, vp, vp, vp, vp, ci, ci, vp, u64]),
    "cz_partition_by_bytes": (ci, [sz, vp, ci, vp]),
    "cz_split_stats": (None, [vp, vp]),
    "cz_profile_enable": (None, [ci]),
    "cz_profile_read": (ci, [vp, vp]),
    "cz_profile_read_deflate": (ci, [vp, vp]),
    "cz_adler32_combine": (u32, [u32, u32, u64]),
    "cz_crc32_combine": (u32, [u32, u32, u64]),
    "cz_synth_model_bytes": (u64, []),
    "cz_synth_build_model": (ci, [vp, u64, vp]),
    "cz_synth_fill_device": (ci, [vp, ci, u64, sz, vp, vp, vp]),
    "cz_synth_fill_host": (ci, [ci, u64, sz, vp, vp, vp]),
}


def lib():
    """Loads the CUDA library. Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "compu_b200: %s is missing — build the CUDA extension first (make -C compu_b200/csrc). "
                "There is no CPU fallback." % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().cz_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("compu_b200: %s failed with code %d: %s" % (what, rc, last_error()))


def require_device():
    n = lib().cz_device_count()
    if n < 1:
        raise RuntimeError("compu_b200: no usable sm_100 CUDA device (%s); there is no CPU fallback" % last_error())
    return n
