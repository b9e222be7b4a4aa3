"""Encoder check on a GPU box: N corpus chunks through k_parse / k_entropy, per-kernel times, byte parity of a sample against the
compiled reference.  usage: gpu_enc.py [N] [reps] [level]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import ctypes as C
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
from tests.oracle_util import oracle_compress, ref_compress, ref

L = _native.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
level = int(sys.argv[3]) if len(sys.argv) > 3 else 3
data = corpus.corpus(n)
ctx = ZstdBatchContext(0)
dev = torch.device("cuda:0")
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(); st = stream.cuda_stream
def comp():
    L.zstdb200_compress_device(ctx.handle, level, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
    L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st)
comp(); torch.cuda.synchronize()
offs = d_ooff.cpu().numpy(); out = d_out[: int(offs[-1])].cpu().numpy()
cmp_fn = ref_compress if ref() is not None else oracle_compress
bad = 0
for i in list(range(min(n, 64))) + list(range(64, n, 61)):
    if out[offs[i]:offs[i + 1]].tobytes() != cmp_fn(data[i].tobytes(), level):
        bad += 1; print("MISMATCH chunk", i, "class", i % 8)
print("compressed", int(offs[-1]), "parity mismatches:", bad, flush=True)
ctx.setOption("timing", 1)
buf = C.create_string_buffer(4096)
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); comp(); e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    L.zstdb200_kernel_times(ctx.handle, buf, 4096)
    print(f"rep {rep}: {ms:.3f} ms -> {n*131072/ms/1e6:.2f} GB/s | {buf.value.decode()}", flush=True)
# the same without per-kernel timers: the entropy stage then runs beside the parse (entropy_overlap, the default)
ctx.setOption("timing", 0)
for ov in (1, 0):
    ctx.setOption("entropy_overlap", ov)
    for rep in range(max(2, reps)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); comp(); e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"untimed rep {rep} entropy_overlap={ov}: {ms:.3f} ms -> {n*131072/ms/1e6:.2f} GB/s", flush=True)
    offs2 = d_ooff.cpu().numpy(); out2 = d_out[: int(offs2[-1])].cpu().numpy()
    print("entropy_overlap", ov, "output identical to the first pass:", bool(int(offs2[-1]) == int(offs[-1]) and (out2 == out).all()), flush=True)
