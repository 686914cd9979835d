"""The two identities the encode kernel's exact path rests on for outputs 0 and 4 of the 8-point DCT (9 in 10 of its
items: the true .5 ties sit at (0,0), (4,0), (0,4), (4,4)), checked bit for bit on the CPU against the oracle's
restatement of ducc0's DCT-II (`oracle/tic_oracle.c:dct8_ducc0`, itself pinned against SciPy in test_oracle_vs_scipy.py):

  column pass (csrc/tic_kernels.cuh, exact_item, TIC_EXACT_INT_COLS; also settle_rational / dc_exact_from_colsums):
      on small integers every operation in front of the last multiplication is exact, so
      y0 = RN(RN(0.5 * (x0+..+x7)) * HSQ),   y4 = RN(RN(0.5 * (x0+x7+x3+x4 - x1-x2-x5-x6)) * TW3)
  row pass (exact_item, TIC_EXACT_ROW_INLINE; dct8_exact's early return): on arbitrary doubles
      y0 / y4 = RN(RN(0.25 * RN(RN(RN(2c0 + 2c7) + 2 RN(c3+c4)) +- 2 RN(RN(c1+c2) + RN(c5+c6)))) * {HSQ, TW3})
"""
import numpy as np

from oracle import oracle_lib as O

HSQ = float.fromhex("0x1.6a09e667f3bcdp-1")
TW3 = float.fromhex("0x1.6a09e667f3bccp-1")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_integer_column_pass_for_outputs_0_and_4():
    rng = np.random.default_rng(7)
    x = rng.integers(-128, 128, (400_000, 8)).astype(np.float64)
    x[:8] = np.array([[-128] * 8, [127] * 8, [127, -128] * 4, [-128, 127] * 4, [0] * 8, [1] * 8,
                      [127, 127, -128, -128, 127, 127, -128, -128], [-128, 127, 127, -128, -128, 127, 127, -128]])
    y = O.dct8_rows(x.copy())
    xi = x.astype(np.int64)
    s0 = xi.sum(axis=1)
    s4 = xi[:, 0] + xi[:, 7] + xi[:, 3] + xi[:, 4] - xi[:, 1] - xi[:, 2] - xi[:, 5] - xi[:, 6]
    y0 = (0.5 * s0.astype(np.float64)) * HSQ      # numpy float64 arithmetic is IEEE round-to-nearest: one rounding per operation
    y4 = (0.5 * s4.astype(np.float64)) * TW3
    assert np.array_equal(_bits(y[:, 0]), _bits(y0))
    assert np.array_equal(_bits(y[:, 4]), _bits(y4))


def test_row_pass_sequence_for_outputs_0_and_4():
    rng = np.random.default_rng(8)
    # what the row pass sees: column results, i.e. doubles of a few hundred in magnitude with full mantissas
    c = rng.standard_normal((400_000, 8)) * rng.choice([1.0, 30.0, 700.0], (400_000, 1))
    c[:1000] = (0.5 * rng.integers(-1024, 1017, (1000, 8)).astype(np.float64)) * HSQ   # column results of real blocks
    y = O.dct8_rows(c.copy())
    c0, c7 = 2.0 * c[:, 0], 2.0 * c[:, 7]
    c1, c3, c5 = c[:, 1] + c[:, 2], c[:, 3] + c[:, 4], c[:, 5] + c[:, 6]
    h0, h3, h1 = c0 + c7, 2.0 * c3, c1 + c5
    a, e1 = h0 + h3, 2.0 * h1
    assert np.array_equal(_bits(y[:, 0]), _bits((0.25 * (a + e1)) * HSQ))
    assert np.array_equal(_bits(y[:, 4]), _bits((0.25 * (a - e1)) * TW3))
