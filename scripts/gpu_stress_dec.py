"""Stress of the staged decoder's in-kernel protocols (table copies, rings, walk -> value links): many batches of varied size, level and
alignment, every result compared with the input.  usage: gpu_stress_dec.py [rounds]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from zstd_jni_b200 import corpus
from zstd_jni_b200.zstd import ZstdBatchContext

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(2026)
pool = [corpus.chunk(i).tobytes() for i in range(96)]
t0 = time.time(); frames_total = 0
with ZstdBatchContext(0) as ctx:
    for r in range(rounds):
        n = int(rng.choice([1, 2, 7, 33, 200, 777, 1500]))
        level = int(rng.choice([1, 3, 3, 3, 5, 9, -3]))
        items = []
        for _ in range(n):
            c = pool[int(rng.integers(0, len(pool)))]
            if rng.random() < 0.5:
                c = c[: int(rng.integers(0, len(c) + 1))]
            items.append(c)
        frames = ctx.compressBatch(items, level)
        back = ctx.decompressBatch(frames, [len(c) for c in items])
        assert back == items, (r, n, level)
        frames_total += n
        if r % 10 == 9:
            print(f"round {r + 1}: {frames_total} frames ok, {time.time() - t0:.1f}s", flush=True)
print("stress ok", frames_total, "frames")
