"""Small fixtures through every default kernel, for compute-sanitizer (tools/sanitize.sh). Kept tiny: the sanitizer slows
kernels down by one to two orders of magnitude. Every result is still checked (against zlib), so a run that passes here
passed functionally too."""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compu_b200 import batch  # noqa: E402
from compu_b200 import decoder as dec  # noqa: E402
from compu_b200 import encoder as enc  # noqa: E402
from compu_b200 import Vec  # noqa: E402

alice = open(os.path.join(ROOT, "tests", "golden", "alice29.txt"), "rb").read()
gz = open(os.path.join(ROOT, "tests", "golden", "alice29.txt.compressed.gz"), "rb").read()

# two-phase inflate (lane-per-stream decode + warp-per-stream LZ77): many small zlib streams, stored / fixed / dynamic blocks
chunks = [alice[i:i + 4096] for i in range(0, 65536, 4096)] + [b"", b"a", bytes(range(256)) * 4, b"ab" * 3000]
streams = [zlib.compress(c, lvl) for c in chunks for lvl in (0, 1, 6)]
outs, st, _, _ = batch.inflate_batch(streams, [len(c) for c in chunks for _ in range(3)], 15)
assert (st == 2).all() and outs == [c for c in chunks for _ in range(3)]
# truncated / corrupt / short-slot units
bad = [streams[2][:40], streams[5][:-3], bytes([streams[8][0], streams[8][1] ^ 0xff]) + streams[8][2:], streams[11]]
batch.inflate_batch(bad, [4096, 4096, 4096, 100], 15)
# warp-per-stream kernel (big unit) and the golden gzip member
outs, st, _, _ = batch.inflate_batch([gz, zlib.compress(alice * 3, 6)], [len(alice), 3 * len(alice)], 47)
assert (st == 2).all() and outs[0] == alice and outs[1] == alice * 3
# deflate chain: three containers, small segments, strategies
for wb in (15, 31, -15):
    comp, st = batch.deflate_batch([alice[:70000], b"", b"x" * 5000], level=6, window_bits=wb, segment_bytes=32768)
    assert (st == 2).all() and zlib.decompress(comp[0], wb) == alice[:70000]
for strategy in (1, 2, 3, 4):
    comp, st = batch.deflate_batch([alice[:30000]], level=6, window_bits=15, strategy=strategy)
    assert zlib.decompress(comp[0]) == alice[:30000]
comp, st = batch.deflate_batch([alice[:30000]], level=0, window_bits=15)
assert zlib.decompress(comp[0]) == alice[:30000]
# one stream of full-flush segments and its segment-parallel decode
s, idx = batch.deflate_segmented(alice, level=6, window_bits=31, segment_bytes=32768)
assert zlib.decompress(s, 31) == alice
assert batch.inflate_segmented(s, len(alice), idx, window_bits=31, segment_bytes=32768) == alice
# streaming objects
e = enc.Interface.zlib_cuda(enc.ZlibOptions().mode(enc.ZlibMode.Gzip).compression(6))
cv = Vec()
for k in range(0, 40000, 10000):
    e.encode_vec_full(alice[k:k + 10000], cv, enc.EncodeOp.Process if k < 30000 else enc.EncodeOp.Finish)
d = dec.Interface.zlib_cuda(dec.ZlibMode.Auto)
pv = Vec()
comp = cv.as_bytes()
r = None
for k in range(0, len(comp), 3000):
    r = d.decode_vec_full(comp[k:k + 3000], pv)
assert r.status == dec.DecodeStatus.Finished and pv.as_bytes() == alice[:40000]
print("sanitize_case ok")
