"""Generates tests/golden/* from the reference compiled in place (oracle/_ref/libzstd-oracle.so).

Run in the dev container (needs /root/reference to have built oracle/_ref):
    python -m tests.golden.make_golden
The fixtures are small on purpose; inputs are regenerated from the deterministic corpus generator, only
their SHA-256 is stored.  Nothing here runs on the GPU box except regenerate_input().
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def regenerate_input(spec: dict) -> bytes:
    from zstd_jni_b200 import corpus
    from tests import cases
    if spec["kind"] == "corpus":
        return corpus.chunk(spec["index"])[: spec["size"]].tobytes()
    if spec["kind"] == "special":
        return dict(cases.special_cases())[spec["name"]]
    if spec["kind"] == "multi":
        return b"".join(corpus.chunk(i).tobytes() for i in spec["indices"])[: spec["size"]]
    raise KeyError(spec["kind"])


def main():
    from tests.oracle_util import ref, ref_compress, ref_decompress, ref_stream_compress
    assert ref() is not None, "oracle/_ref/libzstd-oracle.so missing: run `make -C oracle ref`"
    man = {"generator": "tests/golden/make_golden.py", "reference": "libzstd " + ref().ZSTD_versionString().decode() + " (luben/zstd-jni 1.5.7-16 src/main/native)",
           "oneshot": [], "decode_only": [], "errors": []}
    for f in HERE.glob("*.zst"):
        f.unlink()
    specs = []
    for cls, idx in ((0, 0), (1, 1), (2, 2), (4, 4), (5, 5), (7, 7), (7, 15), (7, 23)):
        for size in (0, 1, 6, 7, 64, 255, 256, 1000, 1024, 5000, 16384, 16385):
            if cls in (7,) and size not in (0, 7, 1000, 16385):
                continue
            specs.append({"kind": "corpus", "index": idx, "size": size})
    for idx in (1, 5, 7, 15, 23, 31):
        specs.append({"kind": "corpus", "index": idx, "size": 131072})
    for name in ("zeros-128k", "period-3", "long-match", "four-symbols-50k"):
        specs.append({"kind": "special", "name": name})
    n = 0
    for spec in specs:
        data = regenerate_input(spec)
        for level in ((3, 1) if len(data) <= 5000 or spec["kind"] == "special" else (3,)):
            frame = ref_compress(data, level)
            assert not isinstance(frame, int)
            if len(frame) > 20000:
                continue
            fn = f"oneshot_{n:03d}_L{level}.zst"; n += 1
            (HERE / fn).write_bytes(frame)
            man["oneshot"].append({"file": fn, "level": level, "input": spec, "input_sha256": hashlib.sha256(data).hexdigest(), "frame_size": len(frame)})
    # lazy levels (row match finder, cost-based table selection): the 128 KB inputs and the specials again
    for spec in specs:
        data = regenerate_input(spec)
        if len(data) <= 16384:
            continue
        for level in (9, 6, 5, 12):
            frame = ref_compress(data, level)
            assert not isinstance(frame, int)
            if len(frame) > 20000:
                continue
            fn = f"oneshot_{n:03d}_L{level}.zst"; n += 1
            (HERE / fn).write_bytes(frame)
            man["oneshot"].append({"file": fn, "level": level, "input": spec, "input_sha256": hashlib.sha256(data).hexdigest(), "frame_size": len(frame)})
    # hash-chain finder (levels 4, 6) and binary tree (level 9) on small inputs
    for spec in specs:
        data = regenerate_input(spec)
        if spec["kind"] != "corpus" or spec["size"] not in (1000, 5000, 16384):
            continue
        for level in (4, 6, 9):
            frame = ref_compress(data, level)
            assert not isinstance(frame, int)
            fn = f"oneshot_{n:03d}_L{level}.zst"; n += 1
            (HERE / fn).write_bytes(frame)
            man["oneshot"].append({"file": fn, "level": level, "input": spec, "input_sha256": hashlib.sha256(data).hexdigest(), "frame_size": len(frame)})
    # decode-only: the reference's streaming path (multi-block, unknown content size, repeat modes)
    multi = {"kind": "multi", "indices": [1, 9, 5, 17], "size": 450000}
    data = regenerate_input(multi)
    for level, checksum in ((1, False), (3, False), (3, True), (9, False)):
        z = ref_stream_compress(data, level, checksum=checksum)
        fn = f"stream_L{level}{'_xxh' if checksum else ''}.zst"
        (HERE / fn).write_bytes(z)
        man["decode_only"].append({"file": fn, "size": len(data), "sha256": hashlib.sha256(data).hexdigest(), "input": multi})
    z3 = ref_stream_compress(data[:150000], 3)
    skippable = b"\x50\x2a\x4d\x18" + (7).to_bytes(4, "little") + b"skipped"
    (HERE / "concat_skippable.zst").write_bytes(skippable + z3 + skippable + z3)
    man["decode_only"].append({"file": "concat_skippable.zst", "size": 300000, "sha256": hashlib.sha256(data[:150000] * 2).hexdigest(), "input": None})
    # error behaviour pinned by the reference
    base = ref_compress(regenerate_input({"kind": "corpus", "index": 1, "size": 20000}), 3)
    probes = {"err_truncated_end.zst": base[:-1], "err_truncated_mid.zst": base[: len(base) // 2], "err_bad_magic.zst": b"\x00" + base[1:],
              "err_reserved_bit.zst": base[:4] + bytes([base[4] | 0x08]) + base[5:], "err_trailing_garbage.zst": base + b"\x01\x02\x03\x04\x05\x06\x07\x08\x09"}
    for fn, blob in probes.items():
        r = ref_decompress(blob, 20000)
        assert isinstance(r, int), fn
        (HERE / fn).write_bytes(blob)
        man["errors"].append({"file": fn, "cap": 20000, "code": -r})
    r = ref_decompress(base, 19999)
    (HERE / "err_dst_too_small.zst").write_bytes(base)
    man["errors"].append({"file": "err_dst_too_small.zst", "cap": 19999, "code": -r})
    (HERE / "manifest.json").write_text(json.dumps(man, indent=1))
    total = sum(f.stat().st_size for f in HERE.glob("*.zst"))
    print(len(man["oneshot"]), "one-shot,", len(man["decode_only"]), "decode-only,", len(man["errors"]), "error fixtures;", total, "bytes")


if __name__ == "__main__":
    main()
