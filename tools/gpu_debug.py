"""Developer aid: run a few encode cases on the GPU and print where they diverge from the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tinyimgcodec_b200 as tic
from oracle import oracle_lib as O
from tests.cases import make_case, synthetic_image

def check(name, img, q):
    t = time.time()
    try:
        e, eo = tic.encode(img, q), O.encode(img, q)
    except Exception as ex:
        print(name, "encode raised", type(ex).__name__, ex); return
    dcok, acok = np.array_equal(e["dc"], eo["dc"]), np.array_equal(e["ac"], eo["ac"])
    msg = f"{name} q{q}: dc {dcok} ac {acok}"
    if not dcok:
        bad = np.nonzero(e["dc"] != eo["dc"])[0]
        msg += f" dcbad n={len(bad)} first={bad[:5]} got={e['dc'][bad[:5]]} want={eo['dc'][bad[:5]]}"
    if not acok:
        bad = np.argwhere(e["ac"] != eo["ac"])
        msg += f" acbad n={len(bad)} first={bad[:5].tolist()}"
    try:
        out, want = tic.compress(img, q), O.compress(img, q)
        same = out == want
        msg += f" stream {same} ({len(out)} vs {len(want)})"
        if not same:
            n = min(len(out), len(want))
            d = next((i for i in range(n) if out[i] != want[i]), n)
            msg += f" firstdiff byte {d}: got {out[d:d+8].hex()} want {want[d:d+8].hex()}"
    except Exception as ex:
        msg += f" compress raised {type(ex).__name__} {ex}"
    print(msg, f"[{time.time()-t:.2f}s] stats={tic.get_encoder().stats()}", flush=True)

check("tiny8", make_case({"kind": "noise", "shape": (8, 8), "seed": 1}), 50)
check("one", make_case({"kind": "noise", "shape": (1, 1), "seed": 1}), 50)
check("syn64", synthetic_image(64, 64, 1), 50)
check("syn256x384", synthetic_image(256, 384, 1), 50)
check("pad37x51", make_case({"kind": "noise", "shape": (37, 51), "seed": 1}), 75)
check("noise256", make_case({"kind": "noise", "shape": (256, 256), "seed": 13}), 90)
check("syn1024", synthetic_image(1024, 1024, 0), 50)
check("syn1024q95", synthetic_image(1024, 1024, 0), 95)
check("flat", np.full((64, 64), 128, np.uint8), 50)
