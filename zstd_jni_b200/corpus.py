"""Deterministic synthetic "Silesia-mix" corpus.

The real Silesia corpus is not available offline, and nothing that runs on the
GPU box may read /root/reference, so the corpus is synthesised: ``n`` chunks of
``chunk_size`` bytes (default 131072), the class of chunk ``i`` being
``i % 8``:

  0    text-like          prose-like tokens from a Zipf vocabulary, light markup
  1    XML-like           the same vocabulary with 60 % tags / numbers
  2    structured binary  64-byte records, 75 % repeated fields, 25 % random
  3    source-like        the text pool with 5 % byte mutations
  4    low-entropy numeric int32 ramp + Gaussian noise (sigma 3), little endian
  5    long-match         a 4 KB motif repeated with 1 % mutations
  6    incompressible     uniform random bytes (-> raw blocks)
  7    degenerate         zeros / single byte / two-symbol runs (-> RLE paths)

Every chunk is a pure function of (seed, i), so ranks can generate their own
shard and tests can regenerate any chunk.  Seed 20240901 is the one quoted in
BASELINE.md.
"""
from __future__ import annotations

import numpy as np

SEED = 20240901
CHUNK = 131072
N_CLASSES = 8
_POOL_BYTES = 4 << 20


def _word_table(rng: np.random.Generator, n_words: int = 4096, max_len: int = 12):
    lens = np.clip(rng.geometric(0.28, n_words) + 1, 2, max_len).astype(np.int64)
    # letters with an English-like skew
    alphabet = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    p = 1.0 / np.arange(1, len(alphabet) + 1) ** 0.9
    p /= p.sum()
    chars = rng.choice(alphabet, size=(n_words, max_len), p=p)
    return chars, lens


def _assemble(chars: np.ndarray, lens: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Concatenate words ``ids`` (vectorised)."""
    l = lens[ids]
    starts = np.cumsum(l) - l
    total = int(l.sum())
    word_of = np.repeat(np.arange(len(ids)), l)
    j = np.arange(total) - np.repeat(starts, l)
    return chars[ids[word_of], j]


_pool_cache: dict = {}


def text_pool(seed: int = SEED, markup: float = 0.25) -> np.ndarray:
    """A 4 MiB XML-like byte pool shared by the text-derived classes.

    ``markup`` is the share of tokens that are tags/numbers rather than words:
    0.25 reads like prose with light markup (ratio ~3 at level 3), 0.6 like a
    record-oriented XML dump (ratio ~6)."""
    key = (seed, markup)
    if key in _pool_cache:
        return _pool_cache[key]
    rng = np.random.default_rng([seed, 0xC0FFEE, int(markup * 1000)])
    chars, lens = _word_table(rng)
    # vocabulary entries get a trailing separator so that concatenation reads as text
    seps = np.frombuffer(b"      \n,.=\"/", dtype=np.uint8)
    tags = [b"<row id=\"", b"\">", b"</row>\n", b"<name>", b"</name>", b"<value unit=\"kg\">", b"</value>",
            b"<!-- ", b" -->\n", b"<item key=\"", b"\" type=\"string\">", b"</item>\n  ", b"<date>2024-09-", b"</date>"]
    n_words = len(lens)
    # extend the vocabulary with tags and numbers (as fixed rows)
    extra = tags + [str(v).encode() for v in rng.integers(0, 100000, 256)]
    max_len = max(chars.shape[1] + 1, max(len(t) for t in extra))
    tab = np.zeros((n_words + len(extra), max_len), dtype=np.uint8)
    tab[:n_words, : chars.shape[1]] = chars
    ln = np.concatenate([lens + 1, np.array([len(t) for t in extra], dtype=np.int64)])
    tab[np.arange(n_words), lens] = rng.choice(seps, n_words)
    for k, t in enumerate(extra):
        tab[n_words + k, : len(t)] = np.frombuffer(t, dtype=np.uint8)
    # Zipf over words, with tags/numbers mixed in at ~25 %
    n_tokens = _POOL_BYTES // 5
    zipf = np.minimum(rng.zipf(1.25, n_tokens) - 1, n_words - 1)
    is_extra = rng.random(n_tokens) < markup
    ids = np.where(is_extra, n_words + rng.integers(0, len(extra), n_tokens), zipf)
    if markup >= 0.5:
        # record-oriented dump: a fixed tag skeleton whose slots take a word or a number
        skel = np.array([0, -1, 1, 3, -1, 4, 5, -2, 6, 9, -1, 10, -1, 11, 2], dtype=np.int64)
        reps = n_tokens // len(skel)
        ids = np.tile(skel, reps)
        slot_w = ids == -1
        slot_n = ids == -2
        ids[slot_w] = np.minimum(rng.zipf(1.5, int(slot_w.sum())) - 1, n_words - 1)
        ids[slot_n] = n_words + len(tags) + rng.integers(0, 256, int(slot_n.sum()))
        ids[~(slot_w | slot_n)] += n_words
    pool = _assemble(tab, ln, ids)
    if len(pool) < _POOL_BYTES:
        pool = np.tile(pool, _POOL_BYTES // len(pool) + 1)
    pool = np.ascontiguousarray(pool[:_POOL_BYTES])
    _pool_cache[key] = pool
    return pool


def chunk(i: int, size: int = CHUNK, seed: int = SEED) -> np.ndarray:
    """Chunk ``i`` of the corpus as a uint8 array of ``size`` bytes."""
    rng = np.random.default_rng([seed, i])
    cls = i % N_CLASSES
    if cls in (0, 1, 3):
        pool = text_pool(seed, 0.6 if cls == 1 else 0.25)
        off = int(rng.integers(0, len(pool) - size))
        out = pool[off : off + size].copy()
        if cls == 3:
            m = rng.random(size) < 0.05
            out[m] = rng.integers(32, 127, int(m.sum()), dtype=np.uint8)
        return out
    if cls == 2:
        n_rec = (size + 63) // 64
        templates = rng.integers(0, 256, (4, 64), dtype=np.uint8)
        recs = templates[rng.integers(0, 4, n_rec)]
        var_cols = rng.permutation(64)[:16]
        recs[:, var_cols] = rng.integers(0, 256, (n_rec, 16), dtype=np.uint8)
        return np.ascontiguousarray(recs.reshape(-1)[:size])
    if cls == 4:
        n = (size + 3) // 4
        ramp = (np.arange(n, dtype=np.int64) * int(rng.integers(1, 9)) + int(rng.integers(0, 1 << 20)))
        vals = (ramp + np.rint(rng.normal(0, 3, n)).astype(np.int64)).astype("<i4")
        return np.ascontiguousarray(vals.view(np.uint8)[:size])
    if cls == 5:
        motif = rng.integers(0, 256, 4096, dtype=np.uint8)
        out = np.tile(motif, size // 4096 + 1)[:size].copy()
        m = rng.random(size) < 0.01
        out[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
        return out
    if cls == 6:
        return rng.integers(0, 256, size, dtype=np.uint8)
    # cls == 7: degenerate inputs, cycling through sub-variants
    sub = (i // N_CLASSES) % 4
    if sub == 0:
        return np.zeros(size, dtype=np.uint8)
    if sub == 1:
        return np.full(size, int(rng.integers(1, 256)), dtype=np.uint8)
    if sub == 2:
        runs = rng.integers(1, 2000, size // 500 + 8)
        vals = np.resize(np.array([65, 66], dtype=np.uint8), len(runs))
        return np.ascontiguousarray(np.repeat(vals, runs)[:size])
    return (rng.random(size) < 0.03).astype(np.uint8) * 255


def corpus(n_chunks: int, size: int = CHUNK, seed: int = SEED, start: int = 0) -> np.ndarray:
    """``n_chunks`` consecutive chunks starting at ``start`` as an (n, size) array."""
    out = np.empty((n_chunks, size), dtype=np.uint8)
    for k in range(n_chunks):
        out[k] = chunk(start + k, size, seed)
    return out
