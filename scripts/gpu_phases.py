"""Per-phase cycle shares of k_entropy (profiling build: nvcc -DZB_PHASE_TIMERS -> lib/libzstdb200_phases.so)."""
import os, sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ["ZSTDB200_LIBRARY"] = str(ROOT / "zstd_jni_b200" / "lib" / "libzstdb200_phases.so")
sys.path.insert(0, str(ROOT))
import torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
L.zstdb200_debug_phases.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0"); data = corpus.corpus(n)
d_src = torch.from_numpy(data.reshape(-1)).to(dev)
d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(); out = (C.c_ulonglong * 16)()
for rep in range(2):
    L.zstdb200_debug_phases(out, 1)
    r = L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), stream.cuda_stream)
    assert r == 0, r
    torch.cuda.synchronize()
L.zstdb200_debug_phases(out, 0)
names = {1: "literal gather", 7: "literal histogram", 8: "Huffman tree + table (lane 0)", 9: "Huffman streams", 2: "seqToCodes", 3: "sequence statistics + tables", 4: "FSE state chains", 5: "sequence bit packing"}
tot = sum(out[k] for k in names)
for k, nm in names.items():
    print(f"{nm:32s} {out[k] / tot * 100:5.1f} %   {out[k] / 1e9:8.2f} Gcycles")
