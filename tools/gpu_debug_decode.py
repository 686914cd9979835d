"""Developer aid: run a few decode cases on the GPU and print where they diverge from the CPU restatement."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import tinyimgcodec_b200 as tic
from oracle import oracle_lib as O
from tests.cases import make_case, synthetic_image


def check(name, img, q, **kw):
    t = time.time()
    s = O.compress(img, q, **kw)
    want = O.decompress(s)
    try:
        got = tic.decompress(s, strict=False)
    except Exception as ex:
        print(name, "raised", type(ex).__name__, ex, flush=True)
        return
    bad = np.argwhere(got != want)
    msg = f"{name} q{q} {img.shape} stream {len(s)} B: {'OK' if len(bad) == 0 else 'BAD'}"
    if len(bad):
        by, bx = bad[:, 0] // 8, bad[:, 1] // 8
        blocks = np.unique(by * ((img.shape[1] + 7) // 8) + bx)
        msg += (f" {len(bad)} pixels in {len(blocks)} blocks, first blocks {blocks[:6].tolist()}, first px {bad[0].tolist()} "
                f"got {got[tuple(bad[0])]} want {want[tuple(bad[0])]} maxdiff {np.abs(got.astype(int) - want.astype(int)).max()}")
    print(msg, tic.get_encoder().decode_stats(), f"[{time.time() - t:.2f}s]", flush=True)


check("flat8", np.full((8, 8), 128, np.uint8), 50)
check("tiny8", make_case({"kind": "noise", "shape": (8, 8), "seed": 1}), 50)
check("one", make_case({"kind": "noise", "shape": (1, 1), "seed": 1}), 50)
check("syn64", synthetic_image(64, 64, 1), 50)
check("syn256x384", synthetic_image(256, 384, 1), 50)
check("pad37x51", make_case({"kind": "noise", "shape": (37, 51), "seed": 1}), 75)
check("noise256", make_case({"kind": "noise", "shape": (256, 256), "seed": 13}), 90)
check("syn1024", synthetic_image(1024, 1024, 0), 50)
check("syn1024q10", synthetic_image(1024, 1024, 0), 10)
check("syn1024q95", synthetic_image(1024, 1024, 0), 95)
check("flat2048", np.full((2048, 2048), 77, np.uint8), 50)
check("auto-le", synthetic_image(256, 384, 1), 50, auto_generate_huffman_table=True, le_flag_word=True)
