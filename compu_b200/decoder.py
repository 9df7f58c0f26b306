"""Decoder side of compu's API, mirrored for the CUDA backend.

Mirrors /root/reference/src/decoder/mod.rs: `Interface` (:160-200), `Decoder` (:269-455), `Decode` (:150-157),
`DecodeStatus` (:139-146), `DecodeError` (:120-135), `Detection::detect` (:9-115), and
src/decoder/zlib_common.rs `ZlibMode` (:4-28). `Interface.zlib_cuda(mode)` is the new backend constructor; it has the
shape of `Interface::zlib_ng(mode)` (src/decoder/zlib_ng.rs:61-90) and returns None when no sm_100 device is usable.
"""
import enum
from dataclasses import dataclass

from . import _lib
from .buffer import Vec, ptr_len


class DecodeStatus(enum.Enum):
    NeedInput = 0
    NeedOutput = 1
    Finished = 2


class DecodeError:
    """`DecodeError(i32)`: raw zlib-numbered code (src/decoder/mod.rs:120-135)."""

    def __init__(self, code):
        self.code = int(code)

    @classmethod
    def no_error(cls):
        return cls(0)

    def as_raw(self):
        return self.code

    def __eq__(self, other):
        return isinstance(other, DecodeError) and other.code == self.code

    def __hash__(self):
        return hash(("DecodeError", self.code))

    def __repr__(self):
        return "DecodeError(%d)" % self.code


class DecodeFailure(Exception):
    """Raised by Buffer.decode where the Rust API returns Err(DecodeError)."""

    def __init__(self, error):
        super().__init__(repr(error))
        self.error = error


@dataclass
class Decode:
    input_remain: int
    output_remain: int
    status: object  # DecodeStatus (Ok) or DecodeError (Err)

    def is_ok(self):
        return isinstance(self.status, DecodeStatus)


class ZlibMode(enum.IntEnum):
    Deflate = -15
    Zlib = 15
    Gzip = 15 + 16
    Auto = 15 + 32

    def max_bits(self):
        return int(self)


class Detection(enum.Enum):
    Zstd = 0
    Gzip = 1
    Zlib = 2
    Unknown = 3

    @staticmethod
    def detect(data):
        """Magic-byte sniffing, same table as src/decoder/mod.rs:28-114 — including its quirk that CINFO=6 (`68 xx`)
        headers are never reported (the arm lacks `return`, :80-82)."""
        b = bytes(data[:4])
        if len(b) < 2:
            return None
        if b[0] == 0x1F and b[1] == 0x8B:
            return Detection.Gzip
        if ((b[0] << 8) | b[1]) % 31 == 0:
            table = {0x78: (0x01, 0x5E, 0x9C, 0xDA), 0x08: (0x1D, 0x5B, 0x99, 0xD7), 0x18: (0x19, 0x57, 0x95, 0xD3),
                     0x28: (0x15, 0x53, 0x91, 0xCF), 0x38: (0x11, 0x4F, 0x8D, 0xCB), 0x48: (0x0D, 0x4B, 0x89, 0xC7),
                     0x58: (0x09, 0x47, 0x85, 0xC3)}
            if b[0] in table and b[1] in table[b[0]]:
                return Detection.Zlib
        if len(b) < 4:
            return None
        if b == b"\x28\xb5\x2f\xfd":
            return Detection.Zstd
        return Detection.Unknown


def _status_from_raw(code):
    if code in (0, 1, 2):
        return DecodeStatus(code)
    # 3 = need dictionary (zlib's Z_NEED_DICT = 2 would collide with Finished in the C result), else zlib code
    return DecodeError(2 if code == 3 else code)


class Interface:
    """Decoder vtable (src/decoder/mod.rs:160-200): decode_fn(state, in_ptr, in_len, out_ptr, out_len) -> Decode,
    reset_fn(state) -> new state or None, drop_fn(state), describe_error_fn(code) -> str or None."""

    def __init__(self, decode_fn, reset_fn, drop_fn, describe_error_fn):
        self.decode_fn = decode_fn
        self.reset_fn = reset_fn
        self.drop_fn = drop_fn
        self.describe_error_fn = describe_error_fn

    def decoder(self, state):
        return Decoder(state, self)

    @staticmethod
    def zlib_cuda(mode=ZlibMode.Auto):
        """B200 backend; None if it cannot be initialised (no usable device), like zlib_ng's None on init failure."""
        L = _lib.lib()
        state = L.cz_decoder_new(int(mode))
        if not state:
            return None
        return ZLIB_CUDA.decoder(state)


def _cuda_decode_fn(state, in_ptr, in_len, out_ptr, out_len):
    r = _lib.lib().cz_decode(state, in_ptr, in_len, out_ptr, out_len)
    return Decode(r.input_remain, r.output_remain, _status_from_raw(r.status))


def _cuda_reset_fn(state):
    return _lib.lib().cz_decoder_reset(state) or None


def _cuda_drop_fn(state):
    _lib.lib().cz_decoder_free(state)


def _cuda_describe_error_fn(code):
    s = _lib.lib().cz_describe_error(code)
    return None if s is None else s.decode()


ZLIB_CUDA = Interface(_cuda_decode_fn, _cuda_reset_fn, _cuda_drop_fn, _cuda_describe_error_fn)


class Decoder:
    """`Decoder` (src/decoder/mod.rs:269-455): an instance pointer plus its vtable."""

    def __init__(self, instance, interface):
        self.instance = instance
        self.interface = interface

    def raw_decode(self, in_ptr, in_len, out_ptr, out_len):
        return self.interface.decode_fn(self.instance, in_ptr, in_len, out_ptr, out_len)

    def decode_uninit(self, input, output):
        ip, il, k1 = ptr_len(input)
        op, ol, k2 = ptr_len(output)
        return self.raw_decode(ip, il, op, ol)

    decode = decode_uninit

    def decode_vec(self, input, output: Vec):
        """src/decoder/mod.rs:323-335: writes into spare capacity; len advances only when status is Ok."""
        spare = output.spare_capacity_mut()
        spare_len = len(spare)
        result = self.decode_uninit(input, spare)
        if result.is_ok():
            output.set_len(output.len() + spare_len - result.output_remain)
        return result

    def decode_vec_full(self, input, output: Vec):
        """src/decoder/mod.rs:360-385 including its reserve policy."""
        RESERVE_DEFAULT = 1024
        input = memoryview(bytes(input))
        n = len(input)
        if n < RESERVE_DEFAULT:
            output.try_reserve_exact(n)
            reserve = n // 3
        elif n < RESERVE_DEFAULT * 16:
            output.try_reserve_exact(n + n // 3)
            reserve = RESERVE_DEFAULT
        else:
            output.try_reserve_exact(n * 2)
            reserve = RESERVE_DEFAULT * 8
        while True:
            result = self.decode_vec(input, output)
            if result.status == DecodeStatus.NeedOutput:
                input = input[len(input) - result.input_remain:]
                output.try_reserve_exact(max(reserve, 1))
                continue
            return result

    def reset(self):
        p = self.interface.reset_fn(self.instance)
        if p:
            self.instance = p  # the returned pointer MUST replace the old one (src/decoder/mod.rs:435-437)
            return True
        return False

    def describe_error(self, error):
        return self.interface.describe_error_fn(error.as_raw())

    def close(self):
        if self.instance:
            self.interface.drop_fn(self.instance)
            self.instance = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
