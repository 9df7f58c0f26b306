"""compu_b200 — a B200-native (sm_100a) DEFLATE-family backend behind compu's `Interface` vtables.

Host-side mirror of the reference's API for this path (same names, argument meaning and error behaviour):
  decoder.Interface.zlib_cuda(mode) -> Decoder      encoder.Interface.zlib_cuda(opts) -> Encoder
  Decoder.decode / decode_vec / decode_vec_full / reset / describe_error
  Encoder.encode / encode_vec / encode_vec_full / reset
  Buffer (Buffer<N>), plus the new batched entry points in compu_b200.batch.
All compute happens in compu_b200/libcompu_b200.so (hand-written CUDA, C ABI in include/compu_b200.h).
There is no CPU fallback: without the built library or without an sm_100 device, calls fail loudly.
"""
from . import decoder, encoder  # noqa: F401
from .buffer import Buffer, PinnedBuffer, Vec  # noqa: F401
from .decoder import Decode, DecodeError, DecodeStatus, Decoder, Detection  # noqa: F401
from .encoder import Encode, EncodeOp, EncodeStatus, Encoder  # noqa: F401
