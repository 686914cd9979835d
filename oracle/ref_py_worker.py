"""TEST INFRASTRUCTURE ONLY — one CPU worker of bench.py's Python-reference timing.

Runs the UNMODIFIED reference encoder tinyimgcodec.codec.compress (codec.py:133-164; from /root/reference in the build
container, from its bytecode under oracle/_ref/py/tinyimgcodec_ref.zip on the GPU box — oracle/Makefile `refpy`) on `count` synthetic
images and prints {"pixels", "seconds", "bytes"} as JSON.  bidict / bitarray are the pure-Python stand-ins of
oracle/standins: the real bitarray is a C extension, so the reference with its real
dependencies would be somewhat faster than what this measures.

    python oracle/ref_py_worker.py <count> <height> <width> <first_seed> [quality]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    count, h, w, seed0 = (int(v) for v in sys.argv[1:5])
    quality = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    from oracle.ref_harness import load_reference
    from tests.cases import synthetic_image
    ref = load_reference()
    imgs = [synthetic_image(h, w, seed=seed0 + i) for i in range(count)]
    ref.compress(imgs[0][:64, :64], quality)   # imports and table construction outside the timed region
    nbytes = 0
    t0 = time.perf_counter()
    for im in imgs:
        nbytes += len(ref.compress(im, quality))
    dt = time.perf_counter() - t0
    print(json.dumps({"pixels": count * h * w, "seconds": dt, "bytes": nbytes}))


if __name__ == "__main__":
    main()
