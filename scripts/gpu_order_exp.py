"""Experiment: how much would a perfect parse-cost estimator buy?  parse_order = 2 orders by the times measured in the previous call."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0"); data = corpus.corpus(n)
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
ctx.setOption("skip_entropy", 1)
for mode in (0, 1, 2, 2, 2):
    ctx.setOption("parse_order", mode)
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
    print(f"parse_order={mode}: parse phase {e0.elapsed_time(e1):7.2f} ms", flush=True)
