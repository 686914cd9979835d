"""CPU: the oracle's FDCT restatement is bit-identical to scipy.fftpack.dct — the
third-party arithmetic the reference calls (tinyimgcodec/utils.py:4,32-37).  Runs on
the GPU box's host too: the oracle's arithmetic is a property of the installed SciPy
binary (platform note)."""
import numpy as np
from scipy.fftpack import dct, idct

from oracle import oracle_lib as O


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_dct8_integer_rows_bitwise():
    rng = np.random.default_rng(0)
    x = rng.integers(-128, 128, (1_000_000, 8)).astype(np.float64)
    assert np.array_equal(_bits(dct(x, norm="ortho", axis=-1)), _bits(O.dct8_rows(x)))


def test_dct8_real_rows_bitwise():
    rng = np.random.default_rng(1)
    x = rng.normal(0, 400, (500_000, 8))
    assert np.array_equal(_bits(dct(x, norm="ortho", axis=-1)), _bits(O.dct8_rows(x)))


def test_dct8_extremes_bitwise():
    rng = np.random.default_rng(2)
    x = rng.choice([-128.0, 127.0], (200_000, 8))
    assert np.array_equal(_bits(dct(x, norm="ortho", axis=-1)), _bits(O.dct8_rows(x)))


def test_quant_table_matches_numpy_expression():
    q_tab = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
                      [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
                      [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                      [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]])
    for quality in range(1, 100):
        factor = 5000 / quality if quality < 50 else 200 - 2 * quality  # utils.py:50
        want = q_tab * factor / 100                                       # utils.py:53
        assert np.array_equal(_bits(want), _bits(O.quant_table(quality))), quality


def test_idct8_rows_bitwise():
    """The decode side: scipy.fftpack.idct (tinyimgcodec/utils.py:40-45) = ducc0's DCT-III."""
    rng = np.random.default_rng(3)
    for x in (rng.normal(0, 400, (500_000, 8)),
              rng.integers(-1024, 1024, (500_000, 8)).astype(np.float64) * 0.4,
              rng.integers(-200, 200, (300_000, 8)).astype(np.float64) * 16.0):
        assert np.array_equal(_bits(idct(x, norm="ortho", axis=-1)), _bits(O.idct8_rows(x)))


def test_idct_2d_blocks_bitwise():
    """Both passes in the reference's order (axis -2, then axis -1) on dequantised integer blocks."""
    rng = np.random.default_rng(4)
    blocks = rng.integers(-40, 40, (20000, 8, 8)).astype(np.float64) * O.quant_table(50)
    want = idct(idct(blocks, norm="ortho", axis=-2), norm="ortho", axis=-1)
    cols = O.idct8_rows(np.ascontiguousarray(blocks.transpose(0, 2, 1))).transpose(0, 2, 1)
    got = O.idct8_rows(np.ascontiguousarray(cols))
    assert np.array_equal(_bits(want), _bits(got))
