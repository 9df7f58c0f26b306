"""The reference's own test protocols, transliterated statement by statement so they read like the originals.

  decoder_test_case      <- /root/reference/tests/decoder.rs:21-77   (test_case)
  encoder_test_case      <- /root/reference/tests/encoder.rs:10-78   (test_case)
  encoder_empty_final    <- /root/reference/tests/encoder.rs:115-173 (test_case_empty_final)
  doc_chunked_roundtrip  <- doctests src/encoder/mod.rs:104-147 / src/decoder/mod.rs:212-268 (chunked encode with Flush,
                            decode fed 4 bytes at a time)

They run against any backend that provides compu_b200.decoder.Decoder / encoder.Encoder objects (the CUDA backend on
the GPU box; the CPU oracle in the CPU-only suite, which pins the oracle to the reference's golden vectors).
"""
from compu_b200 import Buffer, Vec
from compu_b200.decoder import DecodeError, DecodeStatus, Detection
from compu_b200.encoder import EncodeOp, EncodeStatus

N_DATA = 2  # `DATA.len()` in tests/decoder.rs:34 is the length of the 2-element fixture array, so "half" is ONE byte


def decoder_test_case(decoder, data, compressed):
    # Full
    output = bytearray(len(data))
    result = decoder.decode(compressed, output)
    assert result.status == DecodeStatus.Finished
    assert result.input_remain == 0
    assert result.output_remain == 0
    assert data == bytes(output)
    decoder.reset()

    # Partial buffer (1 byte)
    mv = memoryview(output)
    for i in range(len(output)):
        output[i] = 0
    result = decoder.decode(compressed, mv[:N_DATA // 2])
    assert result.status == DecodeStatus.NeedOutput
    assert result.output_remain == 0
    remaining = compressed[len(compressed) - result.input_remain:]
    result = decoder.decode(remaining, mv[N_DATA // 2:])
    assert result.status == DecodeStatus.Finished
    assert data == bytes(output)
    decoder.reset()

    # Buffered decoder
    buffer = Buffer(4096)
    buffer_input = compressed
    out = bytearray()
    while True:
        consumed, status = buffer.decode(decoder, buffer_input)
        buffer_input = buffer_input[consumed:]
        out += buffer.data()
        buffer.consume()
        if status == DecodeStatus.Finished:
            break
    assert data == bytes(out)
    decoder.reset()

    # Full vec
    vec = Vec()
    result = decoder.decode_vec_full(compressed, vec)
    assert result.status == DecodeStatus.Finished
    assert result.input_remain == 0
    assert data == vec.as_bytes()
    decoder.reset()

    error = decoder.describe_error(DecodeError.no_error())
    assert error is not None


def encoder_test_case(encoder, decoder, data, expected_detection):
    compressed = Vec(bytes(len(data)))
    compressed_full = Vec()
    decompressed = bytearray(len(data))
    decompressed_full = Vec()
    result = encoder.encode(data, memoryview(compressed._buf)[:len(data)], EncodeOp.Finish)
    assert result.input_remain == 0

    if result.status == EncodeStatus.NeedOutput:
        # header overhead on tiny data: allocate more space and finalise
        compressed.reserve(100)
        spare = compressed.spare_capacity_mut()
        spare_len = len(spare)
        result = encoder.encode_uninit(b"", spare, EncodeOp.Finish)
        assert result.status == EncodeStatus.Finished
        compressed.set_len(compressed.len() + spare_len - result.output_remain)
    else:
        assert result.status == EncodeStatus.Finished
        compressed.truncate(compressed.len() - result.output_remain)

    assert Detection.detect(compressed.as_bytes()) == expected_detection
    result = decoder.decode(compressed.as_bytes(), decompressed)
    assert result.status == DecodeStatus.Finished
    assert data == bytes(decompressed)

    # Buffered encoder
    encoder.reset()
    buffer = Buffer(4096)
    buffer_input = data
    while True:
        consumed, status = buffer.encode(encoder, buffer_input, EncodeOp.Finish)
        buffer_input = buffer_input[consumed:]
        compressed_full.extend_from_slice(buffer.data())
        buffer.consume()
        assert status != EncodeStatus.Error
        if status == EncodeStatus.Finished:
            break
    assert compressed.len() == compressed_full.len(), "compressed != compressed_full"
    assert compressed.as_bytes() == compressed_full.as_bytes()
    compressed_full.clear()

    # Full vec encoding
    encoder.reset()
    result = encoder.encode_vec_full(data, compressed_full, EncodeOp.Finish)
    assert result.status == EncodeStatus.Finished
    assert result.input_remain == 0
    assert compressed.as_bytes() == compressed_full.as_bytes()

    decoder.reset()
    result = decoder.decode_vec_full(compressed_full.as_bytes(), decompressed_full)
    assert result.status == DecodeStatus.Finished, decoder.describe_error(result.status) if isinstance(result.status, DecodeError) else result
    assert data == decompressed_full.as_bytes()

    encoder.reset()
    decoder.reset()
    return compressed.as_bytes()


def encoder_empty_final(encoder, decoder, data):
    compressed = Vec.with_capacity(len(data))
    decompressed = Vec.with_capacity(len(data) + 100)

    output = compressed.spare_capacity_mut()
    output_len = len(output)
    result = encoder.encode_uninit(data, output, EncodeOp.Process)
    assert result.status != EncodeStatus.Error
    compressed.set_len(output_len - result.output_remain)

    output = compressed.spare_capacity_mut()
    output_len = len(output)
    result = encoder.encode_uninit(data[len(data) - result.input_remain:], output, EncodeOp.Flush)
    assert result.input_remain == 0
    assert result.status == EncodeStatus.Continue
    compressed.set_len(compressed.len() + output_len - result.output_remain)

    compressed.reserve(100)
    output = compressed.spare_capacity_mut()
    output_len = len(output)
    result = encoder.encode_uninit(b"", output, EncodeOp.Finish)
    assert result.status == EncodeStatus.Finished
    compressed.set_len(compressed.len() + output_len - result.output_remain)

    cbytes = compressed.as_bytes()
    step = len(cbytes) // 4
    for i in range(0, len(cbytes), step):
        chunk = cbytes[i:i + step]
        current_len = decompressed.len()
        output = decompressed.spare_capacity_mut()
        output_len = len(output)
        result = decoder.decode_uninit(chunk, output)
        assert result.input_remain == 0
        assert result.output_remain > 0
        decompressed.set_len(current_len + output_len - result.output_remain)
        assert isinstance(result.status, DecodeStatus), result.status
        if result.status == DecodeStatus.Finished:
            break
        assert result.status == DecodeStatus.NeedInput
    assert data == decompressed.as_bytes()

    encoder.reset()
    decoder.reset()


def doc_chunked_roundtrip(encoder, decoder, data, chunk=40):
    """Chunked encode (Process per chunk, Flush at the end of each, Finish last) then decode 4 bytes at a time."""
    out = Vec.with_capacity(len(data) * 2 + 256)
    chunks = [data[i:i + chunk] for i in range(0, len(data), chunk)] or [b""]
    for idx, c in enumerate(chunks):
        op = EncodeOp.Finish if idx == len(chunks) - 1 else EncodeOp.Flush
        result = encoder.encode_vec(c, out, op)
        assert result.input_remain == 0
        assert result.status == (EncodeStatus.Finished if op == EncodeOp.Finish else EncodeStatus.Continue)
    comp = out.as_bytes()
    dst = Vec.with_capacity(len(data) + 64)
    status = None
    for i in range(0, len(comp), 4):
        result = decoder.decode_vec(comp[i:i + 4], dst)
        assert result.input_remain == 0
        status = result.status
        if status == DecodeStatus.Finished:
            break
        assert status == DecodeStatus.NeedInput, status
    assert status == DecodeStatus.Finished
    assert dst.as_bytes() == data
    encoder.reset()
    decoder.reset()
