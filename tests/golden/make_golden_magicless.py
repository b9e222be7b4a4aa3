"""Generates tests/golden/magicless.json from the reference compiled in place: frames written with
ZSTD_c_format = ZSTD_f_zstd1_magicless (J/ZstdCompressCtx.setMagicless, N/jni_zstd.c:362-363) -- size and SHA-256 per case,
plus the error codes the reference's magicless decoder returns on malformed input.
Run in the dev container:  python -m tests.golden.make_golden_magicless
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def main():
    from tests.golden.make_golden import regenerate_input
    from tests.oracle_util import ref, ref_compress_flags, ref_decompress_magicless
    assert ref() is not None
    man = {"generator": "tests/golden/make_golden_magicless.py", "frames": [], "errors": []}
    for spec in ({"kind": "corpus", "index": 1, "size": 131072}, {"kind": "corpus", "index": 4, "size": 70000}, {"kind": "corpus", "index": 2, "size": 300},
                 {"kind": "corpus", "index": 0, "size": 0}, {"kind": "corpus", "index": 5, "size": 6}, {"kind": "special", "name": "random-128k"}):
        data = regenerate_input(spec)
        for level, checksum, content_size in ((3, False, True), (1, True, True), (6, False, False)):
            z = ref_compress_flags(data, level, checksum, content_size, True)
            assert z == ref_compress_flags(data, level, checksum, content_size, False)[4:]          # a magicless frame is the frame minus its magic number
            assert ref_decompress_magicless(z, len(data)) == data
            man["frames"].append({"input": spec, "level": level, "checksum": checksum, "content_size": content_size, "size": len(z),
                                  "sha256": hashlib.sha256(z).hexdigest(), "head": z[:12].hex()})
    data = regenerate_input({"kind": "corpus", "index": 1, "size": 20000})
    z = ref_compress_flags(data, 3, False, True, True)
    full = ref_compress_flags(data, 3, False, True, False)
    for name, blob, cap in (("truncated", z[:-1], 20000), ("trailing", z + b"\x00", 20000), ("tiny", z[:3], 20000), ("reserved-bit", bytes([z[0] | 8]) + z[1:], 20000),
                            ("dst-too-small", z, 19999), ("with-magic", full, 20000), ("two-frames", z + z, 40000)):
        r = ref_decompress_magicless(blob, cap)
        man["errors"].append({"name": name, "blob": blob.hex() if len(blob) < 64 else None, "cap": cap, "result": r if isinstance(r, int) else len(r)})
    (HERE / "magicless.json").write_text(json.dumps(man, indent=1))
    print(len(man["frames"]), "frames,", len(man["errors"]), "decode probes")


if __name__ == "__main__":
    main()
