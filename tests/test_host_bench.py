"""bench.py's host-side pieces that need no GPU: the workload generator is BASELINE.md §4's numpy generator."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tests.cases import synthetic_image  # noqa: E402


@pytest.mark.timeout(120)
def test_bench_generator_is_the_numpy_generator():
    """Host draws (forked workers, shared memory) + the integer upsample in torch == tests/cases.py::synthetic_image
    (the BASELINE.md §4 generator the golden fixtures were made with), seed = global image index."""
    first, count, h, w = 5, 6, 64, 96
    draws = bench.HostDraws(first, count, h, w, workers=3)
    try:
        imgs = bench.synth_images_numpy(draws, "cpu", chunk=4).numpy()
    finally:
        draws.close()
    assert imgs.shape == (count, h, w) and imgs.dtype == np.uint8
    for i in range(count):
        assert np.array_equal(imgs[i], synthetic_image(h, w, first + i)), i


def test_bench_generator_full_size_image():
    g, n = bench.synth_draws(4095)
    img = bench.synth_assemble(g[None], n[None], "cpu")[0].numpy()
    assert np.array_equal(img, synthetic_image(1024, 1024, 4095))
