#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — decode-side fixtures from the UNMODIFIED reference decoder.

Run in the build container (needs /root/reference):  python oracle/gen_golden_decode.py

For every stream in tests/golden/streams.npz (made by oracle/gen_golden.py with the reference's
compress()) and for the extra streams made here, the reference's own
tinyimgcodec.decompress (tinyimgcodec/codec.py:167-189) is run through oracle/ref_harness.py and the
sha256 of the decoded pixels is recorded, together with the number of blocks whose decoding raised
inside the reference's try/except (codec.py:177-185; counted by wrapping nothing — a stream is "clean"
when the oracle restatement, which is checked against the reference in
tests/test_oracle_vs_reference.py, reports zero such blocks).

Outputs:
  tests/golden/decode_streams.npz  extra streams: auto-table streams with the little-endian flag word
                                   (the only auto-table form the reference decoder can open), streams of
                                   the reference's C encoder binary (oracle/_ref/encode, flag bit 30),
                                   truncated streams
  tests/golden/decoded.json        key -> {"shape", "sha256", "clean"} for all of them
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as O  # noqa: E402
from oracle.ref_harness import load_reference  # noqa: E402
from tests.cases import ODD_CASES, make_case, synthetic_image  # noqa: E402
from tests.golden_io import Golden  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = load_reference()
    g = Golden()
    extra = {}
    # auto-table streams the reference decoder can read (flag word little-endian); the oracle's compress is
    # byte-identical to the reference's apart from those four bytes (tests/test_oracle_vs_reference.py)
    for name, img, q in (("lenna", g.images["lenna"], 50), ("47", g.images["47"], 10),
                         ("syn96", synthetic_image(64, 96, 3), 75), ("flat", np.full((24, 40), 77, np.uint8), 50),
                         ("pad", make_case(ODD_CASES["pad_37x51"]), 90)):
        extra[f"autole_{name}_q{q}"] = O.compress(img, q, True, le_flag_word=True)
    # the reference's embedded C encoder (binary built from /root/reference/c by `make -C oracle ref`)
    assert O.ref_c_available(), "run `make -C oracle ref` first"
    for name, img in (("syn64", synthetic_image(64, 64, 1)), ("syn160", synthetic_image(96, 160, 2)),
                      ("noise", make_case({"kind": "noise", "shape": (32, 48), "seed": 4}))):
        for qf in ("best", "high", "med", "low"):
            extra[f"cvar_{name}_{qf}"] = O.ref_c_compress(img, qf)
    # truncated streams: the reference zero-fills what it cannot decode
    s = g.streams["img_lenna_q50"]
    extra["trunc_lenna_half"] = s[: len(s) // 2]
    extra["trunc_lenna_hdr"] = s[:16]
    decoded = {}
    for key, data in list(sorted(g.streams.items())) + list(sorted(extra.items())):
        px = ref.decompress(data)
        opx, nerr = O.decompress(data, return_errors=True)
        assert np.array_equal(px, opx), key
        decoded[key] = {"shape": list(px.shape), "sha256": hashlib.sha256(np.ascontiguousarray(px).tobytes()).hexdigest(),
                        "clean": nerr == 0}
        print(key, px.shape, nerr, flush=True)
    np.savez_compressed(os.path.join(GOLD, "decode_streams.npz"),
                        **{k: np.frombuffer(v, dtype=np.uint8) for k, v in extra.items()})
    decoded["_provenance"] = {"reference": "clysto/tinyimgcodec decompress(), run unmodified",
                              "scipy": __import__("scipy").__version__, "numpy": np.__version__,
                              "generator": "oracle/gen_golden_decode.py"}
    with open(os.path.join(GOLD, "decoded.json"), "w") as f:
        json.dump(decoded, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
