#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — regenerates tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py

The reference has no tests, golden files or known-answer vectors of its own
(SURVEY.md §4), so the pins are outputs of the reference itself, produced here by
importing it through oracle/ref_harness.py (stand-in bidict/bitarray, SciPy 1.18.1,
numpy 2.3.5) and calling tinyimgcodec.compress / tinyimgcodec.encode
(tinyimgcodec/codec.py:26,133).  The fixtures travel to the GPU box; the reference
does not.

Outputs:
  tests/golden/kat_streams.json   size + sha256 of compress() for all 50 data/*.gif at
                                  q50, Lenna at q in {90,80,50,20,10,5}, auto-table
                                  Lenna at q in {90,50,10}  
  tests/golden/images.npz         pixels of a 6-image subset of data/*.gif (inputs)
  tests/golden/gifs_all.npz       pixels of all 50 data/*.gif (BASELINE config 2 inputs; `--images-only`
                                  regenerates just this file)
  tests/golden/streams.npz        full reference streams: subset x qualities, odd
                                  shapes / adversarial images from seeded generators,
                                  auto-table streams
  tests/golden/errors.json        cases where the reference raises (KeyError: category not
                                  in the fixed table, tinyimgcodec/huffman.py:62)
  tests/golden/coeffs.npz         encode() dc/ac arrays for Lenna q50/q90 and a padded case
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_harness import REFERENCE_ROOT, load_reference  # noqa: E402
from tests.cases import ODD_CASES, make_case  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SUBSET = ["lenna", "1", "25", "44", "47", "7"]
SUBSET_Q = [90, 50, 10]


def load_gif(name):
    return np.asarray(Image.open(os.path.join(REFERENCE_ROOT, "data", f"{name}.gif")).convert("L"))


def dump_all_gifs():
    """BASELINE config 2 needs the pixels of all 50 data/*.gif on the GPU box (the reference tree does not
    travel): the decoded grayscale pixels — data, not code — go into tests/golden/gifs_all.npz."""
    names = [str(i) for i in range(1, 50)] + ["lenna"]
    np.savez_compressed(os.path.join(GOLD, "gifs_all.npz"), **{n: load_gif(n) for n in names})


def main():
    if "--images-only" in sys.argv:
        dump_all_gifs()
        return
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    kat = {"q50": {}, "lenna_sweep": {}, "lenna_auto": {}}
    names = [str(i) for i in range(1, 50)] + ["lenna"]
    for n in names:
        out = ref.compress(load_gif(n), quality=50)
        kat["q50"][n] = {"size": len(out), "sha256": hashlib.sha256(out).hexdigest()}
        print(n, len(out), flush=True)
    lenna = load_gif("lenna")
    for q in (90, 80, 50, 20, 10, 5):
        out = ref.compress(lenna, quality=q)
        kat["lenna_sweep"][str(q)] = {"size": len(out), "sha256": hashlib.sha256(out).hexdigest()}
    for q in (90, 50, 10):
        out = ref.compress(lenna, quality=q, auto_generate_huffman_table=True)
        kat["lenna_auto"][str(q)] = {"size": len(out), "sha256": hashlib.sha256(out).hexdigest()}
    kat["_provenance"] = {
        "reference": "clysto/tinyimgcodec mounted at /root/reference, run unmodified",
        "scipy": __import__("scipy").__version__, "numpy": np.__version__,
        "generator": "oracle/gen_golden.py",
    }
    with open(os.path.join(GOLD, "kat_streams.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)

    images = {n: load_gif(n) for n in SUBSET}
    np.savez_compressed(os.path.join(GOLD, "images.npz"), **images)
    dump_all_gifs()

    streams = {}
    for n in SUBSET:
        for q in SUBSET_Q:
            streams[f"img_{n}_q{q}"] = np.frombuffer(ref.compress(images[n], quality=q), dtype=np.uint8)
    for n in ("lenna", "47"):
        for q in (50, 10):
            streams[f"auto_{n}_q{q}"] = np.frombuffer(
                ref.compress(images[n], quality=q, auto_generate_huffman_table=True), dtype=np.uint8)
    errors = {}
    for name, spec in ODD_CASES.items():
        img = make_case(spec)
        for q in spec["qualities"]:
            try:
                streams[f"case_{name}_q{q}"] = np.frombuffer(ref.compress(img, quality=q), dtype=np.uint8)
            except Exception as e:  # the reference's error behaviour is part of the contract
                errors[f"case_{name}_q{q}"] = type(e).__name__
        if spec.get("auto") and img.size:
            q = spec["qualities"][0]
            streams[f"caseauto_{name}_q{q}"] = np.frombuffer(
                ref.compress(img, quality=q, auto_generate_huffman_table=True), dtype=np.uint8)
        print("case", name, flush=True)
    np.savez_compressed(os.path.join(GOLD, "streams.npz"), **streams)
    with open(os.path.join(GOLD, "errors.json"), "w") as f:
        json.dump(errors, f, indent=1, sort_keys=True)

    coeffs = {}
    for q in (50, 90):
        e = ref.encode(lenna, quality=q)
        coeffs[f"lenna_q{q}_dc"] = e["dc"].astype(np.int32)
        coeffs[f"lenna_q{q}_ac"] = e["ac"].astype(np.int16)
    img = make_case(ODD_CASES["pad_37x51"])
    e = ref.encode(img, quality=75)
    coeffs["pad_37x51_q75_dc"] = e["dc"].astype(np.int32)
    coeffs["pad_37x51_q75_ac"] = e["ac"].astype(np.int16)
    np.savez_compressed(os.path.join(GOLD, "coeffs.npz"), **coeffs)
    print("done")


if __name__ == "__main__":
    main()
