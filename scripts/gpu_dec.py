"""Decoder check on a GPU box: round trip of N corpus chunks through the staged decoder, per-kernel times (CUDA events).
usage: gpu_dec.py [N] [reps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import ctypes as C
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext

L = _native.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
level = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t = time.time(); data = corpus.corpus(n); print("corpus", n, "chunks in %.1fs" % (time.time() - t), flush=True)
ctx = ZstdBatchContext(0)
dev = torch.device("cuda:0")
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
d_back = torch.zeros(n * 131072, dtype=torch.uint8, device=dev)
d_res = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(); st = stream.cuda_stream
def comp():
    L.zstdb200_compress_device(ctx.handle, level, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
    L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st)
def decomp():
    return L.zstdb200_decompress_device(ctx.handle, n, d_out.data_ptr(), d_ooff.data_ptr(), d_back.data_ptr(), d_off.data_ptr(), d_res.data_ptr(), st)
comp(); torch.cuda.synchronize()
csize = int(d_ooff[-1].item()); print("compressed", csize, "ratio %.3f" % (n * 131072 / csize), flush=True)
r = decomp(); torch.cuda.synchronize()
print("first decompress rc", r, "roundtrip ok:", bool(torch.equal(d_back, d_src)), "res ok", bool((d_res == 131072).all()), flush=True)
if not torch.equal(d_back, d_src):
    bad = (d_back.view(n, -1) != d_src.view(n, -1)).any(dim=1).nonzero().flatten().tolist()
    print("bad frames", len(bad), bad[:32], "classes", sorted(set(b % 8 for b in bad)))
    res = d_res.cpu().numpy(); print("res of bad", [int(res[b]) for b in bad[:16]])
ctx.setOption("timing", 1)
buf = C.create_string_buffer(4096)
for rep in range(reps):
    d_back.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); decomp(); e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    L.zstdb200_kernel_times(ctx.handle, buf, 4096)
    print(f"rep {rep}: {ms:.3f} ms -> {n*131072/ms/1e6:.1f} GB/s out | {buf.value.decode()}", flush=True)
print("final roundtrip ok:", bool(torch.equal(d_back, d_src)))
for roles in (1, 2):      # which chains bound k_dec_chains: sequence chains alone, Huffman chains alone (outputs are wrong here)
    ctx.setOption("chain_roles", roles)
    for rep in range(2):
        decomp(); torch.cuda.synchronize()
        L.zstdb200_kernel_times(ctx.handle, buf, 4096)
    print(f"roles={roles}: {buf.value.decode()}", flush=True)
ctx.setOption("chain_roles", 3)
