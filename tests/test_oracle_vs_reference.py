"""CPU, build container only: the C restatement against the UNMODIFIED reference run
in-process (oracle/ref_harness.py).  This is what pins the oracle; the GPU box sees
only the golden vectors derived from the same runs."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle_lib as O
from oracle.ref_harness import REFERENCE_ROOT, load_reference
from tests.cases import make_case

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def _gif(name):
    from PIL import Image
    return np.asarray(Image.open(os.path.join(REFERENCE_ROOT, "data", f"{name}.gif")).convert("L"))


def test_all_data_images_hashes_match_kat(golden):
    """Oracle on all 50 data/*.gif == the reference's streams (by the committed hashes)."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(REFERENCE_ROOT, "data", "*.gif")))
    assert len(names) == 50
    for n in names:
        out = O.compress(_gif(n), 50)
        assert hashlib.sha256(out).hexdigest() == golden.kat["q50"][n]["sha256"], n


def test_kat_regenerated_from_reference(ref, golden):
    """Do not trust the table blindly: regenerate a few entries from the reference now."""
    for n in ("lenna", "23", "44"):
        out = ref.compress(_gif(n), quality=50)
        assert hashlib.sha256(out).hexdigest() == golden.kat["q50"][n]["sha256"]


@pytest.mark.parametrize("q", [97, 90, 61, 50, 49, 33, 10, 1])
def test_random_shapes_streams_and_coeffs(ref, q):
    rng = np.random.default_rng(q)
    for _ in range(4):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        kind = ["noise", "synthetic", "binary", "impulse"][int(rng.integers(0, 4))]
        img = make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30))})
        try:
            want = ref.compress(img, quality=q)
        except KeyError:
            with pytest.raises(O.OracleError):
                O.compress(img, q)
            continue
        assert O.compress(img, q) == want, (h, w, kind, q)
        e_ref, e_or = ref.encode(img, quality=q), O.encode(img, q)
        assert np.array_equal(e_ref["dc"], e_or["dc"]) and np.array_equal(e_ref["ac"], e_or["ac"])


def test_auto_table_random(ref):
    rng = np.random.default_rng(123)
    for q in (90, 50, 10):
        for kind in ("noise", "synthetic", "impulse"):
            h, w = int(rng.integers(8, 90)), int(rng.integers(8, 90))
            img = make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30))})
            assert O.compress(img, q, True) == ref.compress(img, quality=q, auto_generate_huffman_table=True)


def test_default_tables_equal_reference_literals(ref):
    """The canonical BITS/HUFFVAL expansion equals constants.py:54-241 string by string."""
    import tinyimgcodec_reference.constants as C
    for cat, code in C.HUFFMAN_CATEGORY_CODEWORD[C.DC].items():
        assert O.default_code(0, cat) == code
    for (run, size), code in C.HUFFMAN_CATEGORY_CODEWORD[C.AC].items():
        assert O.default_code(1, run * 16 + size) == code
    assert len(C.HUFFMAN_CATEGORY_CODEWORD[C.AC]) == 162


def test_roundtrip_decodes_with_reference_decoder(ref):
    from oracle.ref_harness import reference_psnr
    psnr = reference_psnr()
    img = _gif("lenna")
    out = O.compress(img, 50)
    dec = ref.decompress(out)
    assert dec.shape == img.shape
    assert abs(psnr(img, dec) - 35.83) < 0.01


def test_le_flag_word_makes_auto_streams_decodable(ref):
    """SURVEY.md §8(f)4.  The reference writes the auto-table flag MSB-first (codec.py:111) and reads it
    little-endian (codec.py:119): its decoder takes its own auto-table streams for fixed-table ones and fails
    or returns garbage.  The opt-in little-endian flag word differs from the reference's stream in header
    bytes 12..15 only, and the reference decoder then reconstructs exactly what it reconstructs from the
    fixed-table stream of the same quality (same quantised coefficients)."""
    img = _gif("lenna")
    for q in (50, 10):
        plain = O.compress(img, q, True)
        compat = O.compress(img, q, True, le_flag_word=True)
        assert plain == ref.compress(img, quality=q, auto_generate_huffman_table=True)
        assert plain[12:16] == bytes([0x80, 0, 0, 0]) and compat[12:16] == bytes([0, 0, 0, 0x80])
        assert plain[:12] == compat[:12] and plain[16:] == compat[16:]
        want = ref.decompress(O.compress(img, q))
        got = ref.decompress(compat)
        assert np.array_equal(got, want)
        try:
            broken = ref.decompress(plain)
            assert not np.array_equal(broken, want)
        except Exception:
            pass   # the reference decoder may also just fail on its own auto-table stream


@pytest.mark.parametrize("q", [97, 75, 50, 49, 20, 3, 1])
def test_decompress_random_shapes(ref, q):
    """Oracle decompress == the reference's decompress (codec.py:167-189) on its own streams."""
    rng = np.random.default_rng(1000 + q)
    for _ in range(4):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        kind = ["noise", "synthetic", "binary", "impulse", "flat"][int(rng.integers(0, 5))]
        img = make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30)),
                         "value": int(rng.integers(0, 256))})
        try:
            s = ref.compress(img, quality=q)
        except KeyError:
            continue
        got, nerr = O.decompress(s, return_errors=True)
        assert nerr == 0 and np.array_equal(ref.decompress(s), got), (h, w, kind, q)


def test_decompress_data_images(ref):
    for n in ("lenna", "5", "31"):
        for q in (90, 50, 10):
            s = ref.compress(_gif(n), quality=q)
            assert np.array_equal(ref.decompress(s), O.decompress(s)), (n, q)


def test_decompress_damaged_streams_like_the_reference(ref):
    """Truncated streams, and auto-table streams whose flag word the reference misreads: the reference
    swallows one exception per damaged block (codec.py:177-185); the restatement follows it bit for bit."""
    img = make_case({"kind": "synthetic", "shape": (64, 96), "seed": 3})
    s = ref.compress(img, quality=50)
    for cut in (len(s) // 2, len(s) - 3, 20, 16):
        got, nerr = O.decompress(s[:cut], return_errors=True)
        assert nerr > 0 and np.array_equal(ref.decompress(s[:cut]), got), cut
    s = ref.compress(img, quality=50, auto_generate_huffman_table=True)
    assert np.array_equal(ref.decompress(s), O.decompress(s))


def test_decompress_c_variant_streams(ref):
    """Streams of the reference's C encoder binary (flag bit 30): decode() takes the scaled-DCT branch
    (codec.py:58-62)."""
    if not O.ref_c_available():
        pytest.skip("oracle/_ref/encode not built")
    img = make_case({"kind": "synthetic", "shape": (64, 64), "seed": 1})
    for qf in ("best", "high", "med", "low"):
        s = O.ref_c_compress(img, qf)
        got, nerr = O.decompress(s, return_errors=True)
        assert nerr == 0 and np.array_equal(ref.decompress(s), got), qf


def test_decompress_random_sweep(ref):
    """60 seeded random cases: 7 image kinds, shapes 1..90, qualities 1..99, a third with per-image tables (the
    little-endian flag word, the only auto-table form the reference decoder can open)."""
    rng = np.random.default_rng(99)
    n = 0
    while n < 60:
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = ["noise", "synthetic", "binary", "impulse", "flat", "checker", "blockalt"][int(rng.integers(0, 7))]
        q = int(rng.integers(1, 100))
        img = make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30)),
                         "value": int(rng.integers(0, 256))})
        auto = bool(rng.integers(0, 3) == 0)
        try:
            s = O.compress(img, q, auto, le_flag_word=auto) if auto else ref.compress(img, quality=q)
        except (KeyError, O.OracleError):
            continue
        got, nerr = O.decompress(s, return_errors=True)
        assert nerr == 0 and np.array_equal(ref.decompress(s), got), (kind, h, w, q, auto)
        n += 1
