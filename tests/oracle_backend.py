"""Plugs the CPU oracle (oracle/libcompu_oracle.so) into the SAME host-side Interface/Decoder/Encoder mirror the CUDA
backend uses, so the reference's test protocols (tests/protocols.py) run unchanged against either. Tests only."""
import oracle
from compu_b200 import decoder as dec
from compu_b200 import encoder as enc


def _decode_fn(state, ip, il, op, ol):
    r = oracle.lib().oz_decode(state, ip, il, op, ol)
    st = dec.DecodeStatus(r.status) if r.status in (0, 1, 2) else dec.DecodeError(r.status)
    return dec.Decode(r.input_remain, r.output_remain, st)


def _describe(code):
    s = oracle.lib().oz_describe_error(code)
    return None if s is None else s.decode()


ORACLE_DEC = dec.Interface(_decode_fn, lambda s: oracle.lib().oz_decoder_reset(s) or None,
                           lambda s: oracle.lib().oz_decoder_free(s), _describe)


def oracle_decoder(mode=dec.ZlibMode.Auto):
    st = oracle.lib().oz_decoder_new(int(mode))
    return ORACLE_DEC.decoder(st) if st else None


def _encode_fn(state, ip, il, op, ol, eop):
    r = oracle.lib().oz_encode(state, ip, il, op, ol, int(eop))
    return enc.Encode(r.input_remain, r.output_remain, enc.EncodeStatus(r.status))


ORACLE_ENC = enc.Interface(lambda s, o: oracle.lib().oz_encoder_reset(s) or None, _encode_fn,
                           lambda s: oracle.lib().oz_encoder_free(s))


def oracle_encoder(opts=None):
    opts = opts or enc.ZlibOptions()
    st = oracle.lib().oz_encoder_new(opts._compression, int(opts._mode), opts._mem_level, int(opts._strategy))
    return ORACLE_ENC.encoder(st) if st else None
