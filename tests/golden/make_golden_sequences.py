"""Generates tests/golden/sequences.json from the reference compiled in place (oracle/_ref/libzstd-oracle.so):
ZSTD_generateSequences (N/compress/zstd_compress.c:3520-3553) on deterministic inputs -- per case the number of
ZSTD_Sequence records, the SHA-256 of the record array (n x 4 little-endian u32: offset, litLength, matchLength, rep)
and its first rows.  Run in the dev container:  python -m tests.golden.make_golden_sequences
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def specs():
    out = []
    for idx in (0, 1, 2, 3, 4, 5, 7, 9, 12, 14):
        out.append(({"kind": "corpus", "index": idx, "size": 131072}, (3, 1, 5, 9) if idx in (1, 4) else (3,)))
    for idx, size in ((1, 7), (1, 8), (1, 100), (1, 1000), (2, 5000), (4, 16384), (4, 16385), (5, 65792), (0, 100000), (3, 131071)):
        out.append(({"kind": "corpus", "index": idx, "size": size}, (3, 1, 6)))
    for name in ("zeros-128k", "period-3", "long-match", "long-literal-run", "random-128k", "two-symbols"):
        out.append(({"kind": "special", "name": name}, (3, 2, -1)))
    return out


def main():
    from tests.golden.make_golden import regenerate_input
    from tests.oracle_util import ref, ref_generate_sequences
    assert ref() is not None, "oracle/_ref/libzstd-oracle.so missing: run `make -C oracle ref`"
    man = {"generator": "tests/golden/make_golden_sequences.py", "reference": "libzstd " + ref().ZSTD_versionString().decode() + " ZSTD_generateSequences", "cases": []}
    for spec, levels in specs():
        data = regenerate_input(spec)
        for level in levels:
            seqs = ref_generate_sequences(data, level)
            assert not isinstance(seqs, int), (spec, level, seqs)
            man["cases"].append({"input": spec, "input_sha256": hashlib.sha256(data).hexdigest(), "level": level, "count": int(seqs.shape[0]),
                                 "sha256": hashlib.sha256(seqs.astype("<u4").tobytes()).hexdigest(), "head": seqs[:3].tolist()})
    (HERE / "sequences.json").write_text(json.dumps(man, indent=1))
    print(len(man["cases"]), "sequence fixtures")


if __name__ == "__main__":
    main()
