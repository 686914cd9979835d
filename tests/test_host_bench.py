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


def test_bench_helpers_are_not_shadowed():
    """run_gpu_arm's nested helpers (step, barrier) are called by every later leg of the bench: a loop variable of the
    same name silently turns the leg into an error entry of the JSON line (it happened to the decode leg once)."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "run_gpu_arm")
    defs = {n.name for n in ast.walk(fn) if isinstance(n, ast.FunctionDef) and n is not fn}
    assigned = set()
    for n in ast.walk(fn):
        targets = []
        if isinstance(n, ast.Assign):
            targets = n.targets
        elif isinstance(n, (ast.For, ast.comprehension, ast.AugAssign, ast.AnnAssign, ast.NamedExpr)):
            targets = [n.target]
        elif isinstance(n, ast.With):
            targets = [i.optional_vars for i in n.items if i.optional_vars is not None]
        for t in targets:
            assigned |= {x.id for x in ast.walk(t) if isinstance(x, ast.Name)}
    assert defs and not (defs & assigned), sorted(defs & assigned)
