"""Loader for the committed fixtures in tests/golden/ (made by oracle/gen_golden.py)."""
import json
import os

import numpy as np

from tests.cases import ODD_CASES, make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self):
        with open(os.path.join(GOLD, "kat_streams.json")) as f:
            self.kat = json.load(f)
        with open(os.path.join(GOLD, "errors.json")) as f:
            self.errors = json.load(f)
        self.images = dict(np.load(os.path.join(GOLD, "images.npz")))
        self.streams = {k: v.tobytes() for k, v in np.load(os.path.join(GOLD, "streams.npz")).items()}
        self.coeffs = dict(np.load(os.path.join(GOLD, "coeffs.npz")))
        # decode side (oracle/gen_golden_decode.py): pixels of the reference's decompress() by hash
        with open(os.path.join(GOLD, "decoded.json")) as f:
            self.decoded = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        self.decode_streams = {k: v.tobytes()
                               for k, v in np.load(os.path.join(GOLD, "decode_streams.npz")).items()}

    def decode_cases(self, clean=True):
        """Yield (key, stream bytes, shape, sha256 of the reference decoder's pixels).  clean: the reference
        decoded every block without an internal exception (codec.py:177-185)."""
        for key, info in sorted(self.decoded.items()):
            if info["clean"] == clean:
                data = self.streams.get(key) or self.decode_streams[key]
                yield key, data, tuple(info["shape"]), info["sha256"]

    def all_gifs(self):
        """name -> pixels for all 50 data/*.gif (BASELINE config 2); loaded on demand (8 MB)."""
        return dict(np.load(os.path.join(GOLD, "gifs_all.npz")))

    def stream_cases(self, auto=False):
        """Yield (key, image, quality, expected bytes) for default- or auto-table streams."""
        for key, data in sorted(self.streams.items()):
            kind, rest = key.split("_", 1)
            name, q = rest.rsplit("_q", 1)
            q = int(q)
            if kind in ("img", "auto"):
                img = self.images[name]
            else:
                img = make_case(ODD_CASES[name])
            if (kind in ("auto", "caseauto")) == auto:
                yield key, img, q, data

    def error_cases(self):
        for key, exc in sorted(self.errors.items()):
            _, rest = key.split("_", 1)
            name, q = rest.rsplit("_q", 1)
            yield key, make_case(ODD_CASES[name]), int(q), exc
