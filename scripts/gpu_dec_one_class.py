"""Decode n frames of ONE corpus class twice (for ncu: profile the second k_dec_exec / k_dec_chains launch).  usage: cls n"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
cls = int(sys.argv[1]); n = int(sys.argv[2])
dev = torch.device("cuda:0")
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
stream = torch.cuda.Stream(); st = stream.cuda_stream
data = np.stack([corpus.chunk(cls + 8 * (i % 64)) for i in range(n)])
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
d_back = torch.zeros(n * 131072, dtype=torch.uint8, device=dev); d_res = torch.zeros(n, dtype=torch.int64, device=dev)
L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st)
torch.cuda.synchronize()
for rep in range(2):
    L.zstdb200_decompress_device(ctx.handle, n, d_out.data_ptr(), d_ooff.data_ptr(), d_back.data_ptr(), d_off.data_ptr(), d_res.data_ptr(), st)
    torch.cuda.synchronize()
print("ok", bool(torch.equal(d_back, d_src)))
