"""CPU: the C restatement (oracle/tic_oracle.c) against the committed golden vectors
that the unmodified reference produced (oracle/gen_golden.py).  Runs anywhere."""
import hashlib

import numpy as np
import pytest

from oracle import oracle_lib as O


def test_streams_default_table(golden):
    n = 0
    for key, img, q, want in golden.stream_cases(auto=False):
        got = O.compress(img, q)
        assert got == want, key
        n += 1
    assert n >= 50


def test_streams_auto_table(golden):
    n = 0
    for key, img, q, want in golden.stream_cases(auto=True):
        got = O.compress(img, q, True)
        assert got == want, key
        n += 1
    assert n >= 6


def test_error_cases(golden):
    for key, img, q, exc in golden.error_cases():
        assert exc == "KeyError"
        with pytest.raises(O.OracleError) as ei:
            O.compress(img, q)
        assert ei.value.status == 1, key


def test_kat_subset_hashes(golden):
    """Appendix D known-answer hashes for the images whose pixels are committed."""
    for name, img in golden.images.items():
        out = O.compress(img, 50)
        kat = golden.kat["q50"][name]
        assert len(out) == kat["size"]
        assert hashlib.sha256(out).hexdigest() == kat["sha256"]
    lenna = golden.images["lenna"]
    for q, kat in golden.kat["lenna_sweep"].items():
        out = O.compress(lenna, int(q))
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"])
    for q, kat in golden.kat["lenna_auto"].items():
        out = O.compress(lenna, int(q), True)
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"])
    assert golden.kat["q50"]["lenna"]["sha256"] == \
        "4596d8bb0d5577d2d4321e5e2ff8c086ce8fbd313b00d4879d4c65baaa8048a9"


def test_kat_all_50_gifs(golden):
    """The oracle on all 50 data/*.gif against the reference's recorded size + sha256 (config 2 inputs)."""
    gifs = golden.all_gifs()
    assert len(gifs) == 50
    for name, img in gifs.items():
        out = O.compress(img, 50)
        kat = golden.kat["q50"][name]
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"]), name


def test_coefficients(golden):
    from tests.cases import ODD_CASES, make_case
    lenna = golden.images["lenna"]
    for q in (50, 90):
        e = O.encode(lenna, q)
        assert np.array_equal(e["dc"], golden.coeffs[f"lenna_q{q}_dc"])
        assert np.array_equal(e["ac"], golden.coeffs[f"lenna_q{q}_ac"].astype(np.int32))
    e = O.encode(make_case(ODD_CASES["pad_37x51"]), 75)
    assert np.array_equal(e["dc"], golden.coeffs["pad_37x51_q75_dc"])
    assert np.array_equal(e["ac"], golden.coeffs["pad_37x51_q75_ac"].astype(np.int32))
    assert e["height"] == 37 and e["width"] == 51


def test_quality_zero():
    with pytest.raises(O.OracleError) as ei:
        O.compress(np.zeros((8, 8), np.uint8), 0)
    assert ei.value.status == 3


def test_decompress_matches_reference_decoder_hashes(golden):
    """Oracle decompress() == the reference's decompress() (by the committed pixel hashes), including the
    streams the reference only half-decodes (truncated, auto-table streams it misreads)."""
    import hashlib
    n = 0
    for clean in (True, False):
        for key, data, shape, sha in golden.decode_cases(clean):
            px, nerr = O.decompress(data, return_errors=True)
            assert px.shape == shape, key
            assert hashlib.sha256(px.tobytes()).hexdigest() == sha, key
            assert (nerr == 0) == clean, key
            n += 1
    assert n > 50


def test_round_trip_psnr_is_sane(golden):
    """decompress(compress(x)) is close to x: guards against a decoder that matches only by accident."""
    img = golden.images["lenna"]
    px = O.decompress(O.compress(img, 50)).astype(np.float64)
    mse = np.mean((px - img.astype(np.float64)) ** 2)
    assert 30.0 < 10 * np.log10(255.0 ** 2 / mse) < 40.0
