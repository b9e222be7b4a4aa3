"""First-light check on a GPU box: parity of a small batch + rough kernel timings."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext, Zstd
from tests.oracle_util import oracle_compress, ref_compress, ref

L = _native.lib()
print("devices", L.zstdb200_device_count(), torch.cuda.get_device_name(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t = time.time(); data = corpus.corpus(n); print("corpus", n, "chunks in %.1fs" % (time.time() - t))
ctx = ZstdBatchContext(0)
t = time.time(); stream, sizes = ctx.compressChunks(data.reshape(-1), 131072, 3); print("compress e2e %.3fs" % (time.time() - t), "ratio %.3f" % (data.size / stream.size))
# parity vs oracle on the first 64 + every 8th
bad = 0; off = 0
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
cmp_fn = ref_compress if ref() is not None else oracle_compress
for i in range(n):
    if i < 64 or i % 8 == 0:
        exp = cmp_fn(data[i].tobytes(), 3)
        got = stream[offs[i]:offs[i + 1]].tobytes()
        if exp != got:
            bad += 1
            k = next((j for j in range(min(len(exp), len(got))) if exp[j] != got[j]), -1)
            print("MISMATCH chunk", i, "class", i % 8, len(exp), len(got), "first diff", k)
print("compress parity mismatches:", bad)
t = time.time(); out, osz = ctx.decompressFrames(stream, sizes, [131072] * n); print("decompress e2e %.3fs" % (time.time() - t))
print("roundtrip ok:", bool((out.reshape(n, -1) == data).all()), "sizes ok:", bool((osz == 131072).all()))

# device-resident timings
dev = torch.device("cuda:0")
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
d_back = torch.empty(n * 131072, dtype=torch.uint8, device=dev)
d_res = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
st = stream.cuda_stream
def comp():
    L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
    L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st)
def decomp():
    L.zstdb200_decompress_device(ctx.handle, n, d_out.data_ptr(), d_ooff.data_ptr(), d_back.data_ptr(), d_off.data_ptr(), d_res.data_ptr(), st)
for name, opt, vals in (("enc", "enc_warps_per_sm", [0, 12, 8, 4, 2, 1]), ("dec", "dec_warps_per_sm", [0, 8, 4, 2])):
    for v in vals:
        ctx.setOption(opt, v)
        fn = comp if name == "enc" else decomp
        if name == "dec": ctx.setOption("enc_warps_per_sm", 0); comp()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{name} warps/SM={v}: {ms:.2f} ms  -> {n*131072/ms/1e6:.2f} GB/s uncompressed")
torch.cuda.synchronize()
print("device roundtrip ok:", bool(torch.equal(d_back, d_src)), "res ok", bool((d_res == 131072).all()))
