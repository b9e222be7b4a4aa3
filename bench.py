#!/usr/bin/env python
"""bench.py -- the headline benchmark (BASELINE.json: uncompressed GB/s, level-3 compress + decompress).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--chunks C] [--level L]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch: every rank compresses its shard of C 128 KB chunks
(level 3, one frame per chunk, sizes scanned and frames concatenated on the device) and decompresses the
resulting stream back.  Default workload = BASELINE.json configs[1]: a 1 GiB synthetic Silesia-mix corpus
(8192 x 131072 B, seed 20240901) on one B200.  With N GPUs every rank gets its own 8192-chunk shard (weak
scaling); frames are independent so there is no data-path exchange, only an all_gather of the per-frame sizes
(the global stream index) and the timing reduction.

Printed (rank 0, one JSON line): value = uncompressed bytes taken through compress+decompress per second with
inputs resident in HBM (CUDA events on the launching stream, max over ranks); e2e = the same through the
host-memory C-ABI calls (pinned host buffers, H2D + kernels + D2H timed); roofline = dominant kernel vs the
measured HBM peak; cpu_baseline = the reference's own libzstd (oracle/_ref) on this box's host cores.
`--impl reference` times only that CPU path with all host threads and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
CHUNK = 131072
METRIC = "uncompressed GB/s, level-3 compress + decompress round trip (zstd, 128 KB frames)"


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_lib():
    """The reference's own C sources compiled in place (kind 'reference'), else the plain-C port."""
    ref = ROOT / "oracle" / "_ref" / "libzstd-oracle.so"
    if ref.exists():
        L = C.CDLL(str(ref))
        for n, a in (("ZSTD_compress2", [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
                     ("ZSTD_decompressDCtx", [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
                     ("ZSTD_CCtx_setParameter", [C.c_void_p, C.c_int, C.c_int]), ("ZSTD_freeCCtx", [C.c_void_p]), ("ZSTD_freeDCtx", [C.c_void_p])):
            getattr(L, n).argtypes = a
            getattr(L, n).restype = C.c_size_t
        L.ZSTD_createCCtx.restype = C.c_void_p
        L.ZSTD_createDCtx.restype = C.c_void_p
        return L, "reference"
    port = ROOT / "oracle" / "libzso.so"
    if not port.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "libzso.so"], check=True)
    L = C.CDLL(str(port))
    L.zso_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    L.zso_compress.restype = C.c_size_t
    L.zso_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.zso_decompress.restype = C.c_size_t
    return L, "port"


class CpuRoundTrip:
    """Per-call path of N/jni_fast_zstd.c:633-635,825-826 minus JNI: one ctx per thread, static partition of the
    chunks, ZSTD_compress2 then ZSTD_decompressDCtx.  ctypes releases the GIL so the threads run in parallel.
    Buffers are allocated and touched once so that page faults stay out of the timed region."""

    def __init__(self, data: np.ndarray, level: int, threads: int):
        self.L, self.kind = _cpu_lib()
        self.data, self.level, self.threads = data, level, threads
        n = data.shape[0]
        self.bound = CHUNK + (CHUNK >> 8) + 64
        self.comp = np.zeros((n, self.bound), dtype=np.uint8)
        self.sizes = np.zeros(n, dtype=np.int64)
        self.back = np.zeros_like(data)
        self.parts = [range(t, n, threads) for t in range(threads)]
        self.src_ptr = [int(data[i].ctypes.data) for i in range(n)]
        self.comp_ptr = [int(self.comp[i].ctypes.data) for i in range(n)]
        self.back_ptr = [int(self.back[i].ctypes.data) for i in range(n)]
        self.pool = ThreadPoolExecutor(threads)

    def _c(self, idx):
        L, size, bound, level = self.L, self.data.shape[1], self.bound, self.level
        if self.kind == "reference":
            cctx = L.ZSTD_createCCtx()
            L.ZSTD_CCtx_setParameter(cctx, 100, level)
            for i in idx:
                self.sizes[i] = L.ZSTD_compress2(cctx, self.comp_ptr[i], bound, self.src_ptr[i], size)
            L.ZSTD_freeCCtx(cctx)
        else:
            for i in idx:
                self.sizes[i] = L.zso_compress(self.comp_ptr[i], bound, self.src_ptr[i], size, level)

    def _d(self, idx):
        L, size = self.L, self.data.shape[1]
        if self.kind == "reference":
            dctx = L.ZSTD_createDCtx()
            for i in idx:
                L.ZSTD_decompressDCtx(dctx, self.back_ptr[i], size, self.comp_ptr[i], int(self.sizes[i]))
            L.ZSTD_freeDCtx(dctx)
        else:
            for i in idx:
                L.zso_decompress(self.back_ptr[i], size, self.comp_ptr[i], int(self.sizes[i]))

    def run(self):
        """-> (t_compress, t_decompress, compressed_bytes)"""
        t0 = time.perf_counter(); list(self.pool.map(self._c, self.parts)); t1 = time.perf_counter()
        list(self.pool.map(self._d, self.parts)); t2 = time.perf_counter()
        return t1 - t0, t2 - t1, int(self.sizes.sum())

    def check(self):
        assert np.array_equal(self.back, self.data), "CPU reference round trip failed"


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples if len(s) >= 7 for k in range(4) if s[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=8192, help="128 KB chunks per GPU (8192 = 1 GiB, BASELINE.json configs[1])")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--cpu-chunks", type=int, default=0, help="chunks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from zstd_jni_b200 import corpus

    config = {"workload": f"{args.chunks} x {CHUNK} B synthetic Silesia-mix chunks per GPU (seed {corpus.SEED}), level {args.level}, one frame per chunk",
              "chunks_per_gpu": args.chunks, "chunk_bytes": CHUNK, "level": args.level, "cache": "inputs_larger_than_L2 (1 GiB working set per pass vs 126 MB L2)",
              "parallelism": f"frames sharded over {args.gpus} GPU(s), no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = host_threads()
        # bounded sample of the same workload: the first n chunks of the corpus, sized for a few seconds per step
        n = args.cpu_chunks or min(args.chunks, max(512, 128 * threads))
        data = corpus.corpus(n)
        cpu = CpuRoundTrip(data, args.level, threads)
        for _ in range(max(1, args.warmup)):
            cpu.run()
        cpu.check()
        times = []
        for _ in range(args.steps):
            tc, td, csize = cpu.run()
            times.append((tc, td))
        kind = cpu.kind
        tc = float(np.median([t[0] for t in times])); td = float(np.median([t[1] for t in times]))
        U = data.size
        val = U / (tc + td) / 1e9
        print(json.dumps({"metric": METRIC, "value": val, "unit": "GB/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": (tc + td) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": config, "compress_gbs": U / tc / 1e9, "decompress_gbs": U / td / 1e9, "ratio": U / csize,
                          "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind, "cpu": cpu_model(),
                                           "sample": f"first {n} chunks ({U / 2**20:.0f} MiB) of the corpus, {threads} threads, ctx per thread"},
                          "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import torch.distributed as dist
    from zstd_jni_b200 import _native
    from zstd_jni_b200.zstd import ZstdBatchContext
    from zstd_jni_b200 import sharding
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _native.lib()
    ctx = ZstdBatchContext(local)
    n = args.chunks
    dev = torch.device("cuda", local)
    data = corpus.corpus(n, start=rank * n)                                   # this rank's shard
    h_src = torch.from_numpy(data.reshape(-1)).pin_memory()
    d_src = h_src.to(dev)
    d_off = torch.arange(0, (n + 1) * CHUNK, CHUNK, dtype=torch.int64, device=dev)
    stride = (L.ZSTD_compressBound(CHUNK) + 32 + 63) // 64 * 64
    d_slots = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(n, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_back = torch.empty(n * CHUNK, dtype=torch.uint8, device=dev)
    d_res = torch.zeros(n, dtype=torch.int64, device=dev)
    index = {}
    stream = torch.cuda.Stream(device=dev)
    st = stream.cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def check(r):
        if _native.is_error(r):
            raise RuntimeError(f"C ABI error {L.ZSTD_getErrorName(r).decode()} / {L.zstdb200_last_error().decode()}")

    def step(events=None):
        """compress (+scan+concat) then decompress, all on `stream`; optional per-kernel events"""
        if events: events[0].record(stream)
        check(L.zstdb200_compress_device(ctx.handle, args.level, n, d_src.data_ptr(), d_off.data_ptr(), d_slots.data_ptr(), stride, d_sizes.data_ptr(), st))
        if events: events[1].record(stream)
        check(L.zstdb200_compact_device(ctx.handle, n, d_slots.data_ptr(), stride, d_sizes.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), st))
        if world > 1:
            # global stream index (8 B x frames over NCCL); the only exchange on this path
            index["offsets"] = sharding.global_offsets(sharding.gather_sizes(d_sizes, world * n))
        if events: events[2].record(stream)
        check(L.zstdb200_decompress_device(ctx.handle, n, d_out.data_ptr(), d_ooff.data_ptr(), d_back.data_ptr(), d_off.data_ptr(), d_res.data_ptr(), st))
        if events: events[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        assert torch.equal(d_back, d_src) and bool((d_res == CHUNK).all()), "device round trip mismatch"
        csize = int(d_sizes.sum().item())
        sampler = ClockSampler(local); sampler.start()
        launches0 = ctx.kernelLaunches()
        ctx.setOption("timing", 1)          # the library brackets every kernel with CUDA events on its launching stream
        ctx.kernelTimes()
        evs = [[ev() for _ in range(4)] for _ in range(args.steps)]
        barrier()
        t_begin, t_end = ev(), ev()
        t_begin.record(stream)
        for k in range(args.steps):
            step(evs[k])
        t_end.record(stream)
        barrier()
        launches = ctx.kernelLaunches() - launches0
        ktimes = ctx.kernelTimes()
        ctx.setOption("timing", 0)
        clocks = sampler.stop()
    total_ms = t_begin.elapsed_time(t_end)
    k_comp = float(np.mean([e[0].elapsed_time(e[1]) for e in evs])); k_pack = float(np.mean([e[1].elapsed_time(e[2]) for e in evs])); k_dec = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    tms = torch.tensor([total_ms, k_comp, k_pack, k_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, k_comp, k_pack, k_dec = [float(x) for x in tms.tolist()]
    U = n * CHUNK
    ms_per_step = total_ms / args.steps
    value = world * U / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-memory C ABI (pinned buffers; H2D + kernels + D2H inside the timed region)
    h_stream = torch.empty(n * stride, dtype=torch.uint8).pin_memory()
    h_back = torch.empty(n * CHUNK, dtype=torch.uint8).pin_memory()
    fsz = (C.c_size_t * n)(); tot = C.c_size_t(0); dsz = (C.c_size_t * n)()
    dsz_in = (C.c_size_t * n)(*([CHUNK] * n))          # expected sizes (in) -> regenerated sizes (out): refreshed per step
    def e2e_step():
        check(L.zstdb200_compress_chunks(ctx.handle, args.level, h_src.data_ptr(), U, CHUNK, h_stream.data_ptr(), h_stream.numel(), fsz, C.byref(tot)))
        C.memmove(dsz, dsz_in, C.sizeof(dsz))
        check(L.zstdb200_decompress_frames(ctx.handle, h_stream.data_ptr(), fsz, n, h_back.data_ptr(), h_back.numel(), dsz))
    e2e_step(); barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert torch.equal(h_back, h_src), "e2e round trip mismatch"
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * U / float(te.item()) / 1e9

    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak = float(json.loads(peaks_path.read_text())["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    algo_bytes = U + csize                                  # uncompressed + compressed bytes of every frame in the launch
    # per-kernel averages over the timed region (rank 0), from the events the library records around every launch
    kernels = {k: v[0] for k, v in ktimes.items()}
    phases = {"compress": k_comp, "scan+compact": k_pack, "decompress": k_dec}          # API-call brackets, max over ranks
    dom = max(kernels, key=kernels.get)
    traffic = {}
    tpath = ROOT / "profiles" / "dram_traffic.json"          # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)
    if tpath.exists():
        tj = json.loads(tpath.read_text())
        if tj.get("chunks_per_gpu") == n and tj.get("level") == args.level:
            traffic = tj.get("kernels", {})
    def roof(name):
        a = algo_bytes / (kernels[name] * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": traffic.get(name)}
    out = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
           "compress_gbs": world * U / ((k_comp + k_pack) * 1e-3) / 1e9, "decompress_gbs": world * U / (k_dec * 1e-3) / 1e9, "ratio": U / csize,
           "kernel_ms": kernels, "phase_ms": phases,
           "roofline": dict(roof(dom), kernel=dom, peak_source=peak_src, algorithmic_bytes_per_launch=algo_bytes,
                            launches_timed=ktimes[dom][1]),
           "roofline_all": {k: roof(k) for k in kernels},
           "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": U + csize, "d2h_bytes_per_step": csize + U, "ms_per_step": float(te.item()) * 1e3,
                   "api": "zstdb200_compress_chunks + zstdb200_decompress_frames, pinned host buffers"},
           "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline and world == 1:
        threads = host_threads()
        ncpu = args.cpu_chunks or min(n, max(512, 128 * threads))
        cpu = CpuRoundTrip(data[:ncpu], args.level, threads)
        cpu.run(); cpu.check()
        runs = [cpu.run() for _ in range(3)]
        tc = float(np.median([r[0] for r in runs])); td = float(np.median([r[1] for r in runs])); kind = cpu.kind
        Uc = ncpu * CHUNK
        out["cpu_baseline"] = {"value": Uc / (tc + td) / 1e9, "unit": "GB/s", "cores": threads, "kind": kind, "cpu": cpu_model(),
                               "compress_gbs": Uc / tc / 1e9, "decompress_gbs": Uc / td / 1e9,
                               "sample": f"first {ncpu} chunks ({Uc / 2**20:.0f} MiB) of the same corpus, {threads} threads, one ctx per thread"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
