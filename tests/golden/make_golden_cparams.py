"""Generates tests/golden/cparams.json from the reference compiled in place: frames written by ZSTD_compress2 after explicit
compression parameters were set (J/ZstdCompressCtx.setWindowLog / setHashLog / setChainLog / setSearchLog / setMinMatch /
setTargetLength / setStrategy -> ZSTD_c_windowLog ... ZSTD_c_strategy) -- size and SHA-256 per case.
Run in the dev container:  python -m tests.golden.make_golden_cparams
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def cases():
    rnd = random.Random(20240901)
    inputs = [{"kind": "corpus", "index": i, "size": 131072} for i in (0, 1, 2, 4, 5)] + \
             [{"kind": "corpus", "index": 1, "size": 20000}, {"kind": "corpus", "index": 0, "size": 10000}, {"kind": "corpus", "index": 3, "size": 70000},
              {"kind": "corpus", "index": 4, "size": 1000}, {"kind": "special", "name": "period-3"}]
    out = [  # hand-picked: every parameter alone, every strategy forced on the default level, typical user settings
        (inputs[1], 3, {"hashLog": 12}), (inputs[1], 3, {"chainLog": 10}), (inputs[1], 3, {"minMatch": 7}), (inputs[1], 3, {"minMatch": 3}),
        (inputs[1], 3, {"windowLog": 17}), (inputs[1], 3, {"windowLog": 27}), (inputs[1], 1, {"targetLength": 5}), (inputs[1], 6, {"searchLog": 1}),
        (inputs[1], 9, {"searchLog": 7}), (inputs[1], 6, {"targetLength": 64}), (inputs[5], 3, {"windowLog": 15, "hashLog": 20}),
        (inputs[7], 3, {"windowLog": 17, "chainLog": 18, "hashLog": 18, "searchLog": 4, "minMatch": 5, "targetLength": 16, "strategy": 4}),
    ]
    for strat in range(1, 7):
        out.append((inputs[0], 3, {"strategy": strat}))
        out.append((inputs[6], 5, {"strategy": strat, "minMatch": 4 + strat % 3}))
    for _ in range(30):
        params = {}
        for k, rng in (("windowLog", (10, 27)), ("hashLog", (6, 22)), ("chainLog", (6, 22)), ("searchLog", (1, 9)), ("minMatch", (3, 7)), ("targetLength", (0, 200)), ("strategy", (1, 6))):
            if rnd.random() < 0.4:
                params[k] = rnd.randint(*rng)
        if not params:
            params = {"hashLog": rnd.randint(6, 20)}
        out.append((rnd.choice(inputs), rnd.choice([1, 2, 3, 4, 5, 6, 7, 9, 10, 12, -3]), params))
    return out


def supported(data: bytes, level: int, params: dict) -> bool:
    """What this build takes: the window must cover the input, strategies up to btlazy2, levels with a GPU parser."""
    wl = params.get("windowLog")
    if wl and (1 << wl) < len(data):
        return False
    return not (len(data) <= 16384 and level > 10)


def main():
    from tests.golden.make_golden import regenerate_input
    from tests.oracle_util import ref, ref_compress_params
    assert ref() is not None
    man = {"generator": "tests/golden/make_golden_cparams.py", "cases": []}
    for spec, level, params in cases():
        data = regenerate_input(spec)
        z = ref_compress_params(data, level, params)
        assert not isinstance(z, int), (spec, level, params, z)
        man["cases"].append({"input": spec, "level": level, "params": params, "supported": supported(data, level, params), "size": len(z),
                             "sha256": hashlib.sha256(z).hexdigest()})
    (HERE / "cparams.json").write_text(json.dumps(man, indent=1))
    print(len(man["cases"]), "cases,", sum(1 for c in man["cases"] if not c["supported"]), "outside the supported set")


if __name__ == "__main__":
    main()
