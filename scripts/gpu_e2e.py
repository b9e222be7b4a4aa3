"""e2e timing of the host-memory API for several slice counts."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
CH = 131072
data = corpus.corpus(n)
h_src = torch.from_numpy(data.reshape(-1)).pin_memory()
stride = (L.ZSTD_compressBound(CH) + 32 + 63) // 64 * 64
h_stream = torch.empty(n * stride, dtype=torch.uint8).pin_memory(); h_back = torch.empty(n * CH, dtype=torch.uint8).pin_memory()
fsz = (C.c_size_t * n)(); tot = C.c_size_t(0); dsz = (C.c_size_t * n)()
for slices, dslices in ((2, 2), (1, 1), (2, 1), (1, 2)):
    ctx.setOption("host_slices", slices); ctx.setOption("host_slices_dec", dslices)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        r = L.zstdb200_compress_chunks(ctx.handle, 3, h_src.data_ptr(), n * CH, CH, h_stream.data_ptr(), h_stream.numel(), fsz, C.byref(tot)); assert r == 0, r
        t1 = time.perf_counter()
        for i in range(n): dsz[i] = CH
        t1b = time.perf_counter()
        r = L.zstdb200_decompress_frames(ctx.handle, h_stream.data_ptr(), fsz, n, h_back.data_ptr(), h_back.numel(), dsz); assert r == 0, r
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1b))
    assert torch.equal(h_back, h_src)
    c = min(t[0] for t in ts[1:]); d = min(t[1] for t in ts[1:])
    print(f"slices={slices}/{dslices}: compress {c*1e3:7.1f} ms ({n*CH/c/1e9:5.2f} GB/s)  decompress {d*1e3:7.1f} ms ({n*CH/d/1e9:5.2f} GB/s)  round trip {n*CH/(c+d)/1e9:5.2f} GB/s", flush=True)
