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
