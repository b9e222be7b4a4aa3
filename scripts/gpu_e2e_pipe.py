"""The asynchronous begin/end pipeline of bench.py's e2e leg with host timestamps per call (where does a step wait?).  usage: [steps]"""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n = 8192; CH = 131072; U = n * CH
data = corpus.corpus(n)
h_src = torch.from_numpy(data.reshape(-1)).pin_memory()
stride = (L.ZSTD_compressBound(CH) + 32 + 63) // 64 * 64
h_stream = [torch.empty(n * stride, dtype=torch.uint8).pin_memory() for _ in range(2)]
h_back = [torch.empty(n * CH, dtype=torch.uint8).pin_memory() for _ in range(2)]
fsz = [(C.c_size_t * n)() for _ in range(2)]; dsz = [(C.c_size_t * n)() for _ in range(2)]
dsz_in = (C.c_size_t * n)(*([CH] * n)); tot = C.c_size_t(0)
def chk(r): assert r == 0, (r, L.zstdb200_last_error())
def cb(k): chk(L.zstdb200_compress_chunks_begin(ctx.handle, k % 2, 3, h_src.data_ptr(), U, CH))
def ce(k): chk(L.zstdb200_compress_chunks_end(ctx.handle, k % 2, h_stream[k % 2].data_ptr(), h_stream[k % 2].numel(), fsz[k % 2], C.byref(tot)))
def db(k): chk(L.zstdb200_decompress_frames_begin(ctx.handle, 2 + k % 2, h_stream[k % 2].data_ptr(), fsz[k % 2], n, h_back[k % 2].data_ptr(), h_back[k % 2].numel(), dsz_in))
def de(k): chk(L.zstdb200_decompress_frames_end(ctx.handle, 2 + k % 2, dsz[k % 2]))
def run(K, log=None):
    T = time.perf_counter
    t0 = T(); cb(0)
    for k in range(K):
        a = T()
        if k + 1 < K: cb(k + 1)
        b = T(); ce(k); c = T(); db(k); d = T()
        if k >= 1: de(k - 1)
        e = T()
        if log is not None: log.append((a - t0, b - a, c - b, d - c, e - d))
    de(K - 1); torch.cuda.synchronize()
    return T() - t0
run(2)
for mode in (1, 0, 2):
    ctx.setOption("entropy_overlap", 1 if mode else 0); ctx.setOption("kernel_fifo", 0 if mode == 2 else 1)
    log = []; tt = run(K, log)
    print(f"entropy_overlap={mode}: {K} steps in {tt*1e3:.1f} ms = {tt/K*1e3:.1f} ms/step -> {U/(tt/K)/1e9:.2f} GB/s")
    for k, (at, tcb, tce, tdb, tde) in enumerate(log):
        print(f"  step {k:2d} @ {at*1e3:7.1f} ms: cb {tcb*1e3:6.1f}  ce {tce*1e3:6.1f}  db {tdb*1e3:6.1f}  de(k-1) {tde*1e3:6.1f}")
assert torch.equal(h_back[0], h_src) and torch.equal(h_back[1], h_src)
