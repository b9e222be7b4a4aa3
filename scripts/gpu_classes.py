"""Per-class kernel timings (which chunk class has the longest serial chain?)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from zstd_jni_b200 import corpus, _native
from zstd_jni_b200.zstd import ZstdBatchContext
L = _native.lib(); ctx = ZstdBatchContext(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
stride = (L.ZSTD_compressBound(131072) + 32 + 63) // 64 * 64
stream = torch.cuda.Stream(); st = stream.cuda_stream
mode = sys.argv[2] if len(sys.argv) > 2 else "classes"
if mode == "sweep":
    n = int(sys.argv[1]); data = corpus.corpus(n)
    d_src = torch.from_numpy(data.reshape(-1)).to(dev)
    d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
    d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
    for lanes, wps, pb, gran in ((32, 0, 8, 64), (32, 0, 10, 64), (32, 0, 8, 32), (32, 0, 10, 32)):
        for _rep in (0,):
            ctx.setOption("l2_fetch_granularity", gran)
            ctx.setOption("parse_lanes", lanes); ctx.setOption("enc_warps_per_sm", wps); ctx.setOption("skip_entropy", 1); ctx.setOption("parse_blocks_per_sm", pb)
            with torch.cuda.stream(stream):
                fn = lambda: L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
            print(f"parse only: lanes={lanes:2d} blocks/SM={pb:2d} l2fetch={gran:3d}: {e0.elapsed_time(e1):8.2f} ms for {n} frames", flush=True)
    sys.exit(0)
for cls in list(range(8)) + [-1]:
    idx = [cls + 8 * k for k in range(n)] if cls >= 0 else list(range(n))
    data = np.stack([corpus.chunk(i) for i in idx])
    d_src = torch.from_numpy(data.reshape(-1)).to(dev)
    d_off = torch.arange(0, (n + 1) * 131072, 131072, dtype=torch.int64, device=dev)
    d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev); d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_back = torch.empty(n * 131072, dtype=torch.uint8, device=dev); d_res = torch.zeros(n, dtype=torch.int64, device=dev)
    def comp():
        L.zstdb200_compress_device(ctx.handle, 3, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st)
        L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st)
    def dec():
        L.zstdb200_decompress_device(ctx.handle, n, d_out.data_ptr(), d_ooff.data_ptr(), d_back.data_ptr(), d_off.data_ptr(), d_res.data_ptr(), st)
    res = {}
    with torch.cuda.stream(stream):
        for name, fn in (("comp", comp), ("dec", dec)):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1)
    ratio = data.size / float(d_sizes.sum().item())
    print(f"class {cls:2d}: n={n} ratio {ratio:6.2f}  compress {res['comp']:8.2f} ms  decompress {res['dec']:7.2f} ms", flush=True)
