import sys; sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
data = corpus.corpus(n); ctx = ZstdBatchContext(0); dev = torch.device("cuda:0")
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(); st = stream.cuda_stream
L.zstdb200_last_error.restype = C.c_char_p
for k in range(3):
    d_sizes.zero_(); torch.cuda.synchronize()
    r = L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
    torch.cuda.synchronize()
    sz = d_sizes.cpu().numpy()
    print("call", k, "rc", r, "err", L.zstdb200_last_error(), "total", int(sz.sum()), "zero sizes", int((sz == 0).sum()), "huge", int((sz > 200000).sum()), "min", int(sz.min()), flush=True)
