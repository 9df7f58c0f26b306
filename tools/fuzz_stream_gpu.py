#!/usr/bin/env python
"""One-off fuzz of the STREAMING Decoder / Encoder (cz_decode / cz_encode through the Python mirror) with random call
patterns, against the CPU oracle driven with the SAME calls (development tool). Checked: the bytes produced, the final
status, and the per-call invariants of the reference contract (NeedInput => all input consumed; NeedOutput => output full
or no progress; statuses of corrupted / truncated streams equal to the oracle's at the end)."""
import os
import random
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from compu_b200.decoder import DecodeStatus, Interface, ZlibMode  # noqa: E402
from helpers import make_data, zcomp  # noqa: E402
from oracle_backend import oracle_decoder  # noqa: E402

MODES = {15: ZlibMode.Zlib, 31: ZlibMode.Gzip, -15: ZlibMode.Deflate, 47: ZlibMode.Auto}


def drive(dec, stream, in_chunks, out_chunks, total_cap):
    """Feeds `stream` in pieces of in_chunks (cycled), output windows of out_chunks (cycled). Returns (bytes, last status, log)."""
    out = bytearray()
    pos, ic, oc, log = 0, 0, 0, []
    pending = b""
    for _ in range(20000):
        if not pending and pos < len(stream):
            k = in_chunks[ic % len(in_chunks)]; ic += 1
            pending = stream[pos:pos + k]; pos += k
        window = bytearray(out_chunks[oc % len(out_chunks)]); oc += 1
        r = dec.decode(pending, window)
        produced = len(window) - r.output_remain
        out += window[:produced]
        consumed = len(pending) - r.input_remain
        log.append((len(pending), len(window), consumed, produced, r.status))
        pending = pending[consumed:]
        if not isinstance(r.status, DecodeStatus) or r.status == DecodeStatus.Finished:
            return bytes(out), r.status, log
        if r.status == DecodeStatus.NeedInput:
            assert r.input_remain == 0, log[-1]
            # (zlib reports Z_OK with avail_in == 0 — NeedInput in compu's glue — also while it still holds output that did
            # not fit the window: with the input exhausted, only a call that makes no progress means "truncated")
            if pos >= len(stream) and not pending and produced == 0 and consumed == 0:
                return bytes(out), r.status, log
        if len(out) > total_cap + 16:
            break
    return bytes(out), None, log


def main():
    s0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    alice = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read()
    n = 0
    for seed in range(s0, s0 + ns):
        rng = random.Random(seed)
        wb = rng.choice([15, 31, -15, 47])
        d = make_data(rng, rng.randrange(5), rng.choice([0, 1, 5, 100, 3000, 20000, 70000]), alice)
        s = zcomp(d, rng.choice([0, 1, 6, 9]), wb if wb != 47 else rng.choice([15, 31]), rng.choice([0, 0, 1, 2, 3, 4]))
        mode = rng.randrange(4)
        if mode == 1 and len(s) > 4:
            s = s[:rng.randrange(1, len(s))]
        elif mode == 2 and len(s) > 8:
            b = bytearray(s); k = rng.randrange(len(b)); b[k] ^= 1 << rng.randrange(8); s = bytes(b)
        elif mode == 3:
            s = s + b"trailing"
        in_chunks = [rng.choice([1, 2, 3, 7, 64, 1000, 1 << 20]) for _ in range(5)]
        out_chunks = [rng.choice([1, 2, 5, 100, 4096, 1 << 17]) for _ in range(5)]
        gpu = Interface.zlib_cuda(MODES[wb])
        ref = oracle_decoder(MODES[wb])
        assert gpu is not None and ref is not None
        o1, st1, log1 = drive(gpu, s, in_chunks, out_chunks, len(d))
        o2, st2, log2 = drive(ref, s, in_chunks, out_chunks, len(d))
        same_status = (st1 == st2) or (not isinstance(st1, DecodeStatus) and not isinstance(st2, DecodeStatus) and st1.as_raw() == st2.as_raw())
        assert same_status, "seed %d wb %d mode %d: final status %r vs oracle %r (%d calls / %d calls)" % (seed, wb, mode, st1, st2, len(log1), len(log2))
        if isinstance(st1, DecodeStatus):
            assert o1 == o2, "seed %d: bytes differ (%d vs %d)" % (seed, len(o1), len(o2))
        else:
            k = min(len(o1), len(o2))
            assert o1[:k] == o2[:k], "seed %d: partial output is not a prefix" % seed
        gpu.close(); ref.close()
        n += 1
    print("stream fuzz ok: %d streams with random call patterns, final status and bytes equal to the oracle's" % n)


if __name__ == "__main__":
    main()
