"""Shared parity inputs: sizes straddling every header-format boundary of the frame/literals/sequence
sections (T/scala/Zstd.scala:20-23 uses sizes 0..130 KiB for the same reason) x the corpus classes."""
from __future__ import annotations

import numpy as np

from zstd_jni_b200 import corpus

EDGE_SIZES = [0, 1, 2, 5, 6, 7, 8, 9, 20, 31, 32, 63, 64, 65, 100, 127, 128, 255, 256, 257, 511, 1000, 1023, 1024, 1025,
              4095, 4096, 5000, 16383, 16384, 16385, 32768, 65535, 65536, 65791, 65792, 100000, 131071, 131072]


def edge_cases(classes=(0, 2, 4, 5, 7), sizes=EDGE_SIZES):
    out = []
    for c in classes:
        full = corpus.chunk(c + 8 * 3)
        for n in sizes:
            out.append((f"class{c}-n{n}", full[:n].tobytes()))
    return out


def special_cases():
    rng = np.random.default_rng(7)
    c = []
    c.append(("zeros-128k", bytes(131072)))
    c.append(("random-128k", rng.integers(0, 256, 131072, dtype=np.uint8).tobytes()))
    c.append(("two-symbols", rng.integers(0, 2, 131072, dtype=np.uint8).tobytes()))
    c.append(("four-symbols-50k", rng.integers(0, 4, 50000, dtype=np.uint8).tobytes()))
    c.append(("period-3", (b"abc" * 50000)[:131072]))
    c.append(("period-1-then-noise", b"\x55" * 70000 + rng.integers(0, 256, 61072, dtype=np.uint8).tobytes()))
    c.append(("long-literal-run", rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + b"xyz" * 2000))   # litLength > 65535
    c.append(("long-match", b"q" * 3 + bytes(range(256)) * 4 + b"\x00" * 100000))                           # matchLength > 65535
    c.append(("skewed-bytes", rng.choice(np.arange(256, dtype=np.uint8), 131072, p=np.r_[0.6, np.full(255, 0.4 / 255)]).tobytes()))
    c.append(("ascii-digits", rng.integers(48, 58, 100000, dtype=np.uint8).tobytes()))
    # many equal symbol counts >= 165: exercises the reference's unstable quicksort tie order in HUF_sort
    c.append(("flat-64-symbols", np.tile(np.arange(64, dtype=np.uint8), 2048)[rng.permutation(131072)].tobytes()))
    c.append(("flat-200-symbols", np.resize(np.arange(200, dtype=np.uint8), 131072)[rng.permutation(131072)].tobytes()))
    return c


def corpus_cases(n=24):
    return [(f"corpus-{i}", corpus.chunk(i).tobytes()) for i in range(n)]
