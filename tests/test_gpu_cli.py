"""GPU: the command-line contract.  ./encode.py SRC OUT.img prints exactly the reference's two lines
(encode.py:15-16 of the reference) and writes the reference's bytes; ./decode.py reads them back (the non-GUI
half of viewer.py:8-20)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_lib as O
from tests.cases import synthetic_image

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv):
    return subprocess.run([sys.executable, *argv], cwd=ROOT, capture_output=True, text=True, timeout=300)


def test_encode_then_decode_cli(tmp_path):
    from PIL import Image
    img = synthetic_image(120, 200, seed=4)
    src, out, back = tmp_path / "in.png", tmp_path / "out.img", tmp_path / "back.png"
    Image.fromarray(img, mode="L").save(src)
    r = _run("encode.py", str(src), str(out))
    assert r.returncode == 0, r.stderr
    want = O.compress(img, 50)
    assert out.read_bytes() == want
    n = len(want)
    assert r.stdout.splitlines() == [f"{n} bytes", f"Compression Ratio: {200 * 120 / n}:1"]   # encode.py:15-16
    r = _run("decode.py", str(out), str(back))
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == ["200x120 quality 50"]
    assert np.array_equal(np.asarray(Image.open(back)), O.decompress(want))


def test_batch_cli(tmp_path):
    from PIL import Image
    imgs = {"a": synthetic_image(64, 64, 1), "b": synthetic_image(40, 72, 2), "c": np.full((9, 17), 200, np.uint8)}
    for k, im in imgs.items():
        Image.fromarray(im, mode="L").save(tmp_path / f"{k}.png")
    r = _run("encode.py", "--batch", str(tmp_path / "enc"), *[str(tmp_path / f"{k}.png") for k in imgs])
    assert r.returncode == 0, r.stderr
    for k, im in imgs.items():
        assert (tmp_path / "enc" / f"{k}.img").read_bytes() == O.compress(im, 50), k
    r = _run("decode.py", "--batch", str(tmp_path / "dec"), *[str(tmp_path / "enc" / f"{k}.img") for k in imgs])
    assert r.returncode == 0, r.stderr
    for k, im in imgs.items():
        got = np.asarray(Image.open(tmp_path / "dec" / f"{k}.png"))
        assert np.array_equal(got, O.decompress(O.compress(im, 50))), k


def test_usage_errors():
    assert _run("encode.py").returncode == 2
    assert _run("decode.py", "only-one-argument").returncode == 2
