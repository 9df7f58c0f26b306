"""`Buffer<N>` and a minimal `Vec<u8>` — host-side mirrors of the reference's staging types.

Buffer: /root/reference/src/buffer.rs:4-49 (fixed array + cursor; split_buffer/data/consume/spare_capacity_mut),
        Buffer::decode src/decoder/mod.rs:507-531, Buffer::encode src/encoder/mod.rs:395-412.
Vec:    the subset of alloc::vec::Vec<u8> the reference's helpers use (len/capacity/spare capacity/set_len/reserve_exact),
        so decode_vec / decode_vec_full / encode_vec / encode_vec_full keep their exact semantics.
"""
import ctypes

import numpy as np

_DUMMY = np.zeros(16, dtype=np.uint8)


def ptr_len(obj):
    """(address, length, keepalive) of any bytes-like object; zero-length objects get a valid non-null address."""
    if isinstance(obj, np.ndarray):
        a = obj.reshape(-1).view(np.uint8)
    else:
        a = np.frombuffer(obj, dtype=np.uint8)
    if a.size == 0:
        return _DUMMY.ctypes.data, 0, _DUMMY
    return a.ctypes.data, a.size, a


class Vec:
    """Growable byte vector with Rust's len/capacity split."""

    def __init__(self, data=b""):
        self._buf = bytearray(data)
        self._len = len(data)

    @classmethod
    def with_capacity(cls, cap):
        v = cls()
        v._buf = bytearray(cap)
        return v

    def len(self):
        return self._len

    def __len__(self):
        return self._len

    def capacity(self):
        return len(self._buf)

    def reserve_exact(self, additional):
        need = self._len + additional
        if need > len(self._buf):
            nb = bytearray(need)  # a fresh allocation: live memoryviews of the old one stay valid, like a realloc'd Vec
            nb[:self._len] = self._buf[:self._len]
            self._buf = nb

    reserve = reserve_exact
    try_reserve_exact = reserve_exact

    def spare_capacity_mut(self):
        return memoryview(self._buf)[self._len:]

    def set_len(self, n):
        assert n <= len(self._buf)
        self._len = n

    def clear(self):
        self._len = 0

    def truncate(self, n):
        if n < self._len:
            self._len = n

    def extend_from_slice(self, data):
        self.reserve_exact(len(data))
        self._buf[self._len:self._len + len(data)] = bytes(data)
        self._len += len(data)

    def as_bytes(self):
        return bytes(self._buf[:self._len])

    def __bytes__(self):
        return self.as_bytes()

    def __eq__(self, other):
        return self.as_bytes() == bytes(other)


class Buffer:
    """Fixed-size staging buffer with a cursor (`Buffer<N>`, src/buffer.rs:4-49)."""

    def __init__(self, n=4096):
        assert n >= 128, "Buffer less than 128 bytes makes no sense"  # src/buffer.rs:12
        self._buf = bytearray(n)
        self.cursor = 0

    def split_buffer(self):
        mv = memoryview(self._buf)
        return mv[:self.cursor], mv[self.cursor:]

    def data(self):
        return bytes(self._buf[:self.cursor])

    def consume(self):
        self.cursor = 0

    def spare_capacity_mut(self):
        return memoryview(self._buf)[self.cursor:]

    def decode(self, decoder, input):
        """src/decoder/mod.rs:517-530 -> (consumed, status) or raises DecodeFailure carrying the DecodeError."""
        from .decoder import DecodeError, DecodeFailure
        spare = self.spare_capacity_mut()
        spare_len = len(spare)
        result = decoder.decode_uninit(input, spare)
        if isinstance(result.status, DecodeError):
            raise DecodeFailure(result.status)
        self.cursor = self.cursor + spare_len - result.output_remain
        return len(input) - result.input_remain, result.status

    def encode(self, encoder, input, op):
        """src/encoder/mod.rs:403-411 -> (consumed, status)."""
        spare = self.spare_capacity_mut()
        spare_len = len(spare)
        result = encoder.encode_uninit(input, spare, op)
        self.cursor = self.cursor + spare_len - result.output_remain
        return len(input) - result.input_remain, result.status


class PinnedBuffer(Buffer):
    """`Buffer<N>` over page-locked host memory (cz_host_alloc — the pinned analogue of compu_malloc, src/mem.rs:27-49):
    the batched entry points DMA straight from / to it, no driver staging. Same cursor API as Buffer."""

    def __init__(self, n=1 << 20):
        from . import _lib
        assert n >= 128, "Buffer less than 128 bytes makes no sense"
        self._lib = _lib.lib()
        self._ptr = self._lib.cz_host_alloc(n)
        if not self._ptr:
            raise RuntimeError("compu_b200: cz_host_alloc(%d) failed: %s" % (n, _lib.last_error()))
        self._buf = np.ctypeslib.as_array(ctypes.cast(self._ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
        self.cursor = 0

    def data(self):
        return self._buf[:self.cursor].tobytes()

    def view(self):
        """numpy view of the whole pinned array (zero copy)."""
        return self._buf

    def close(self):
        if getattr(self, "_ptr", None):
            self._buf = None
            self._lib.cz_host_free(ctypes.c_void_p(self._ptr))
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
