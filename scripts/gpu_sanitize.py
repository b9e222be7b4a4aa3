"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel on a few frames of every kind.
usage: compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from zstd_jni_b200 import corpus
from zstd_jni_b200.zstd import ZstdBatchContext, Zstd, ZstdCompressCtx

rng = np.random.default_rng(1)
chunks = [corpus.chunk(i).tobytes() for i in range(8)] + [corpus.chunk(i)[: int(rng.integers(1, 131072))].tobytes() for i in range(8)] + [b"", b"a", bytes(7), bytes(131072)]
with ZstdBatchContext(0) as b:
    for level in (3, 1, -2, 5, 9, 12):
        todo = [c for c in chunks if level < 11 or len(c) > 16384]
        frames = b.compressBatch(todo, level)
        back = b.decompressBatch(frames, [len(c) for c in todo])
        assert back == todo, level
        print("level", level, "ok", sum(map(len, frames)), flush=True)
    # corrupted frames: error paths of the staged and the fused decoder
    frames = b.compressBatch(chunks[:8], 3)
    bad = []
    for f in frames:
        a = bytearray(f); a[int(rng.integers(9, len(a)))] ^= 0x5A; bad.append(bytes(a))
    res = b.decompressBatch(bad, [131072] * len(bad), raise_on_error=False)
    print("corrupted:", [r if isinstance(r, int) else len(r) for r in res], flush=True)
# libzstd-compatible one-shot layer, flags, multi-frame extension, streaming-style multi-block decode
with ZstdCompressCtx() as c:
    c.setLevel(3).setChecksum(True)
    z = c.compress(chunks[0]); assert Zstd.decompress(z, len(chunks[0])) == chunks[0]
    c.setMultiFrame(True)
    big = b"".join(chunks[:3]); z = c.compress(big); assert Zstd.decompress(z, len(big)) == big
# r4 additions: sequence export, magicless frames, explicit parameters
with ZstdBatchContext(0) as b:
    seqs = b.generateSequences(chunks[:6] + [chunks[9]], 3)
    print("sequences:", [s.shape[0] for s in seqs], flush=True)
    b.setOption("magicless", 1)
    frames = b.compressBatch(chunks[:4] + [b""], 3)
    assert b.decompressBatch(frames, [len(c) for c in chunks[:4]] + [0]) == chunks[:4] + [b""]
    b.setOption("magicless", 0)
    for k, v in (("c_hashLog", 10), ("c_chainLog", 9), ("c_minMatch", 6), ("c_strategy", 4), ("c_searchLog", 6)):
        b.setOption(k, v)
    frames = b.compressBatch(chunks[:10], 3)
    assert b.decompressBatch(frames, [len(c) for c in chunks[:10]]) == chunks[:10]
    print("cparams ok", sum(map(len, frames)), flush=True)
# round 2 additions: streams through the batch layer (direct buffers), the asynchronous begin/end pair, packed (unaligned) outputs
import io, ctypes as C
from zstd_jni_b200.zstd import ByteBuffer, ZstdDirectBufferCompressingStream, ZstdDirectBufferDecompressingStream, ZstdOutputStream, ZstdInputStream
data = b"".join(chunks[:5]) + chunks[9]
src = ByteBuffer.allocateDirect(len(data)); src.array[:] = np.frombuffer(data, dtype=np.uint8)
tgt = ByteBuffer.allocateDirect(len(data) + 4096)
with ZstdDirectBufferCompressingStream(tgt, 3) as zc:
    zc.compress(src)
tgt.flip(); back = ByteBuffer.allocateDirect(len(data) + 1)
zd = ZstdDirectBufferDecompressingStream(tgt)
while zd.hasRemaining():
    if zd.read(back) == 0 and not back.hasRemaining(): break
zd.close()
assert back.array[: len(data)].tobytes() == data
sink = io.BytesIO()
with ZstdOutputStream(sink, 1) as zo:
    zo.write(data)
with ZstdInputStream(io.BytesIO(sink.getvalue())) as zi:
    assert zi.read() == data
with ZstdBatchContext(0) as b:
    a = np.frombuffer(b"".join(chunks[:8]), dtype=np.uint8).copy()
    out = np.empty(a.size + 65536, dtype=np.uint8); sz = (C.c_size_t * 8)()
    b.compressChunksBegin(0, a, 131072, 3); b.compressChunksBegin(1, a, 131072, 1)
    t0 = b.compressChunksEnd(0, out, sz); b.compressChunksEnd(1, np.empty(a.size + 65536, dtype=np.uint8))
    bk = np.empty(a.size, dtype=np.uint8); cap = (C.c_size_t * 8)(*([131072] * 8)); res = (C.c_size_t * 8)()
    b.decompressFramesBegin(2, out[:t0], sz, bk, cap); b.decompressFramesEnd(2, res)
    assert (bk == a).all()
    ragged = [chunks[8], chunks[3][:1001], chunks[10], chunks[5][:77777], chunks[0][:3]]          # packed outputs at odd addresses
    fr = b.compressBatch(ragged, 3)
    assert b.decompressBatch(fr, [len(c) for c in ragged]) == ragged
print("round 2 paths ok", flush=True)
print("sanitize script done")
