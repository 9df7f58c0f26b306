"""Encoder side of compu's API, mirrored for the CUDA backend.

Mirrors /root/reference/src/encoder/mod.rs: `Interface` (:52-92), `Encoder` (:148-330), `Encode` (:42-49),
`EncodeStatus` (:27-38), `EncodeOp` (:12-23) and src/encoder/zlib_common.rs `ZlibOptions`/`ZlibMode`/`ZlibStrategy`
(:5-103). `Interface.zlib_cuda(opts)` has the shape of `Interface::zlib_ng(opts)` (src/encoder/zlib_ng.rs:50-87).
"""
import enum
from dataclasses import dataclass

from . import _lib
from .buffer import Vec, ptr_len


class EncodeOp(enum.IntEnum):
    Process = 0
    Flush = 1   # a sync-flush point, as in the reference (src/encoder/mod.rs:338)
    Finish = 2


class EncodeStatus(enum.Enum):
    Continue = 0
    NeedOutput = 1
    Finished = 2
    Error = 3


@dataclass
class Encode:
    input_remain: int
    output_remain: int
    status: EncodeStatus


class ZlibStrategy(enum.IntEnum):
    Default = 0
    Filtered = 1
    HuffmanOnly = 2
    Rle = 3
    Fixed = 4


class ZlibMode(enum.IntEnum):
    Deflate = -15
    Zlib = 15
    Gzip = 15 + 16


MAX_MEM_LEVEL = 8


class ZlibOptions:
    """Builder with the reference defaults: Gzip, Default strategy, memLevel 8, level 9 (zlib_common.rs:59-66)."""

    def __init__(self):
        self._mode = ZlibMode.Gzip
        self._strategy = ZlibStrategy.Default
        self._mem_level = MAX_MEM_LEVEL
        self._compression = 9

    @classmethod
    def new(cls):
        return cls()

    def mode(self, new_mode):
        self._mode = ZlibMode(new_mode)
        return self

    def strategy(self, new_strategy):
        self._strategy = ZlibStrategy(new_strategy)
        return self

    def mem_level(self, mem_level):
        # The reference's setter asserts `mem_level > MAX_MEM_LEVEL` (zlib_common.rs:88), i.e. it panics for every
        # valid value, so memLevel is effectively always 8. Valid values are accepted here and otherwise ignored.
        assert 0 < mem_level <= 9
        self._mem_level = mem_level
        return self

    def compression(self, compression):
        self._compression = int(compression)
        return self


class Interface:
    """Encoder vtable (src/encoder/mod.rs:52-92): reset_fn(state, opts) -> state or None,
    encode_fn(state, in_ptr, in_len, out_ptr, out_len, op) -> Encode, drop_fn(state)."""

    def __init__(self, reset_fn, encode_fn, drop_fn):
        self.reset_fn = reset_fn
        self.encode_fn = encode_fn
        self.drop_fn = drop_fn

    def encoder(self, state, opts=(0, 0)):
        return Encoder(state, self, opts)

    @staticmethod
    def zlib_cuda(opts=None):
        opts = opts or ZlibOptions()
        L = _lib.lib()
        state = L.cz_encoder_new(opts._compression, int(opts._mode), opts._mem_level, int(opts._strategy))
        if not state:
            return None
        return ZLIB_CUDA.encoder(state, (0, 0))


def _cuda_encode_fn(state, in_ptr, in_len, out_ptr, out_len, op):
    r = _lib.lib().cz_encode(state, in_ptr, in_len, out_ptr, out_len, int(op))
    return Encode(r.input_remain, r.output_remain, EncodeStatus(r.status))


def _cuda_reset_fn(state, opts):
    return _lib.lib().cz_encoder_reset(state) or None


def _cuda_drop_fn(state):
    _lib.lib().cz_encoder_free(state)


ZLIB_CUDA = Interface(_cuda_reset_fn, _cuda_encode_fn, _cuda_drop_fn)


class Encoder:
    """`Encoder` (src/encoder/mod.rs:148-330)."""

    def __init__(self, instance, interface, opts=(0, 0)):
        self.instance = instance
        self.interface = interface
        self.opts = opts

    def raw_encode(self, in_ptr, in_len, out_ptr, out_len, op):
        return self.interface.encode_fn(self.instance, in_ptr, in_len, out_ptr, out_len, op)

    def encode_uninit(self, input, output, op):
        ip, il, k1 = ptr_len(input)
        op_, ol, k2 = ptr_len(output)
        return self.raw_encode(ip, il, op_, ol, op)

    encode = encode_uninit

    def encode_vec(self, input, output: Vec, op):
        """src/encoder/mod.rs:203-213: len always advances by what was written."""
        spare = output.spare_capacity_mut()
        spare_len = len(spare)
        result = self.encode_uninit(input, spare, op)
        output.set_len(output.len() + spare_len - result.output_remain)
        return result

    def encode_vec_full(self, input, output: Vec, op):
        """src/encoder/mod.rs:238-267 including its reserve policy."""
        RESERVE_DEFAULT = 1024
        input = memoryview(bytes(input))
        n = len(input)
        if n < RESERVE_DEFAULT:
            output.try_reserve_exact(n)
            reserve = n // 3
        elif n < RESERVE_DEFAULT * 16:
            output.try_reserve_exact(n // 2)
            reserve = RESERVE_DEFAULT
        else:
            output.try_reserve_exact(n // 3)
            reserve = RESERVE_DEFAULT * 8
        while True:
            result = self.encode_vec(input, output, op)
            if result.status == EncodeStatus.NeedOutput:
                input = input[len(input) - result.input_remain:]
                output.try_reserve_exact(max(reserve, 1))
                continue
            if result.status == EncodeStatus.Continue and op == EncodeOp.Finish:
                input = input[len(input) - result.input_remain:]
                continue
            return result

    def reset(self):
        p = self.interface.reset_fn(self.instance, self.opts)
        if p:
            self.instance = p
            return True
        return False

    def close(self):
        if self.instance:
            self.interface.drop_fn(self.instance)
            self.instance = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
