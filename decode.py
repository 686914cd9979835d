#!/usr/bin/env python3
"""B200 decoder CLI — the non-GUI part of the reference's viewer.py (viewer.py:8-20: read the .img files,
decompress each):

    ./decode.py IN.img OUT.png                 one stream -> one 8-bit grayscale image (any format Pillow writes)
    ./decode.py --batch OUTDIR IN.img...       many streams in one GPU batch; writes OUTDIR/<stem>.png

Prints "<width>x<height> quality <q>" per stream.  Image writing (Pillow) stays on the CPU.
"""
import os
import sys


def main(argv):
    from PIL import Image
    from tinyimgcodec import decompress, decompress_batch
    from tinyimgcodec_b200 import parse_header
    if len(argv) >= 3 and argv[0] == "--batch":
        outdir, sources = argv[1], argv[2:]
        os.makedirs(outdir, exist_ok=True)
        streams = [open(p, "rb").read() for p in sources]
        for src, data, px in zip(sources, streams, decompress_batch(streams)):
            h, w, q, _ = parse_header(data)
            stem = os.path.splitext(os.path.basename(src))[0]
            print(f"{stem}: {w}x{h} quality {q}")
            Image.fromarray(px, mode="L").save(os.path.join(outdir, stem + ".png"))
        return 0
    if len(argv) != 2:
        sys.stderr.write(__doc__)
        return 2
    data = open(argv[0], "rb").read()
    h, w, q, _ = parse_header(data)
    print(f"{w}x{h} quality {q}")
    Image.fromarray(decompress(data), mode="L").save(argv[1])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
