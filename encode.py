#!/usr/bin/env python3
"""B200 encoder CLI.

    ./encode.py SRC OUT.img            one image  (the reference's CLI contract, encode.py:10-19:
                                       prints "<n> bytes" and "Compression Ratio: <r>:1", writes OUT)
    ./encode.py --batch OUTDIR SRC...  many images in one GPU batch; writes OUTDIR/<stem>.img and
                                       prints the same two lines per image, prefixed by the name

Image decoding (Pillow, convert("L")) stays on the CPU exactly as in the reference; the
fixed Huffman tables are used (auto_generate_huffman_table=False, encode.py:12).
"""
import os
import sys

import numpy as np


def _load_gray(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("L"))


def _report(n_pixels, stream, prefix=""):
    print(f"{prefix}{len(stream)} bytes")
    print(f"{prefix}Compression Ratio: {n_pixels / len(stream)}:1")


def main(argv):
    from tinyimgcodec import compress, compress_batch
    if len(argv) >= 3 and argv[0] == "--batch":
        outdir, sources = argv[1], argv[2:]
        os.makedirs(outdir, exist_ok=True)
        pixels = [_load_gray(p) for p in sources]
        for src, px, stream in zip(sources, pixels, compress_batch(pixels)):
            stem = os.path.splitext(os.path.basename(src))[0]
            _report(px.shape[0] * px.shape[1], stream, prefix=f"{stem}: ")
            with open(os.path.join(outdir, stem + ".img"), "wb") as f:
                f.write(stream)
        return 0
    if len(argv) != 2:
        sys.stderr.write(__doc__)
        return 2
    px = _load_gray(argv[0])
    stream = compress(px, auto_generate_huffman_table=False)
    _report(px.shape[1] * px.shape[0], stream)
    with open(argv[1], "wb") as f:
        f.write(stream)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
