"""Stress of the overlapped compression stages (k_entropy as programmatic dependent of k_parse + completion queue) and of the kernel FIFO:
odd batch sizes, repeated calls, several levels, two work sets in flight, inputs that fail in the parse.  Every result is compared with a
run that has the overlap switched off."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0); dev = torch.device("cuda:0")
CH = 131072; stride = (L.ZSTD_compressBound(CH) + 32 + 63) // 64 * 64
stream = torch.cuda.Stream(); st = stream.cuda_stream
def compress(n, level, start, overlap):
    ctx.setOption("entropy_overlap", overlap)
    data = corpus.corpus(n, start=start)
    d_src = torch.from_numpy(data.reshape(-1)).to(dev)
    d_off = torch.arange(0, (n + 1) * CH, CH, dtype=torch.int64, device=dev)
    d_slots = torch.zeros(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
    d_out = torch.zeros(n * stride, dtype=torch.uint8, device=dev); d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    r = L.zstdb200_compress_device(ctx.handle, level, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st); assert r == 0, r
    r = L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st); assert r == 0, r
    torch.cuda.synchronize()
    tot = int(d_ooff[-1]); return d_out[:tot].cpu().numpy().tobytes(), d_sizes.cpu().numpy()
bad = 0; cases = 0
for n, level, start in ((1, 3, 0), (2, 3, 5), (3, 1, 9), (31, 3, 2), (257, 3, 100), (1000, 3, 7), (4737, 3, 11), (5000, 1, 3), (9001, 3, 0), (700, 9, 40), (900, 5, 77), (16385, 3, 0), (64, 12, 8)):
    a, sa = compress(n, level, start, 1)
    for rep in range(2):
        a2, _ = compress(n, level, start, 1); bad += a2 != a
    b, sb = compress(n, level, start, 0)
    ok = a == b and (sa == sb).all(); cases += 1; bad += not ok
    print(f"n={n} level={level}: {'ok' if ok else 'MISMATCH'} ({len(a)} bytes)", flush=True)
# two operations in flight on two work sets through the host API, several rounds, alternating sizes
for overlap in (1, 0):
    ctx.setOption("entropy_overlap", overlap)
    outs = []
    for rnd in range(3):
        ns = (3000 + 17 * rnd, 1200 + rnd)
        hs = [torch.from_numpy(corpus.corpus(n, start=rnd * 50 + k).reshape(-1)).pin_memory() for k, n in enumerate(ns)]
        ho = [torch.empty(n * stride, dtype=torch.uint8).pin_memory() for n in ns]
        fs = [(C.c_size_t * n)() for n in ns]; tot = [C.c_size_t(0), C.c_size_t(0)]
        for k, n in enumerate(ns): assert L.zstdb200_compress_chunks_begin(ctx.handle, k, 3, hs[k].data_ptr(), n * CH, CH) == 0
        for k, n in enumerate(ns): assert L.zstdb200_compress_chunks_end(ctx.handle, k, ho[k].data_ptr(), ho[k].numel(), fs[k], C.byref(tot[k])) == 0
        outs.append([ho[k][: tot[k].value].numpy().tobytes() for k in range(2)])
    if overlap: ref_outs = outs
    else:
        same = outs == ref_outs; cases += 1; bad += not same; print("two work sets in flight:", "ok" if same else "MISMATCH", flush=True)
# an input whose parse fails (chunk > 128 KB through the device API): statuses, not a hang
ctx.setOption("entropy_overlap", 1)
n = 40; big = 200000
d_src = torch.zeros(n * big, dtype=torch.uint8, device=dev); d_off = torch.arange(0, (n + 1) * big, big, dtype=torch.int64, device=dev)
bstride = 262144; d_slots = torch.zeros(n * bstride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
r = L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), bstride, d_sizes.data_ptr(), st); torch.cuda.synchronize()
sz = d_sizes.cpu().numpy().astype(np.uint64)
errs = int((sz > np.uint64(2**63)).sum()); cases += 1; bad += errs != n
print("oversized chunks: rc", r, "error statuses", errs, "of", n, flush=True)
print("cases", cases, "bad", bad)
