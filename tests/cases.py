"""Seeded input generators shared by the golden-vector generator and the parity tests.

All generators are integer-only and deterministic so that the build container (where
the reference runs) and the GPU box (where only the fixtures travel) see the same
pixels.  `synthetic_image` is the BASELINE.md §4 generator for configs C3/C4/C5.
"""
import numpy as np


def synthetic_image(h, w, seed=0):
    """Natural-like synthetic grayscale image."""
    rng = np.random.default_rng(seed)
    gh, gw = h // 32 + 2, w // 32 + 2
    g = rng.integers(0, 256, (gh, gw)).astype(np.int64)
    ys, xs = np.arange(h), np.arange(w)
    gy, fy = ys // 32, (ys % 32)[:, None]
    gx, fx = xs // 32, (xs % 32)[None, :]
    a = g[gy][:, gx]
    b = g[gy][:, gx + 1]
    c = g[gy + 1][:, gx]
    d = g[gy + 1][:, gx + 1]
    base = ((a * (32 - fx) + b * fx) * (32 - fy) + (c * (32 - fx) + d * fx) * fy) // 1024
    noise = rng.integers(-10, 11, (h, w))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def big_synthetic(h, w, seed=0, cell=4096):
    """A large natural-like image in seconds: one `cell` x `cell` synthetic_image tiled over h x w, with a
    different brightness offset per tile so that no two tile rows give the same stream."""
    base = synthetic_image(cell, cell, seed)
    ny, nx = (h + cell - 1) // cell, (w + cell - 1) // cell
    img = np.tile(base, (ny, nx))[:h, :w]
    for ty in range(ny):
        for tx in range(nx):
            off = (37 * ty + 11 * tx + seed) % 23 - 11
            blk = img[ty * cell:(ty + 1) * cell, tx * cell:(tx + 1) * cell]
            np.clip(blk.astype(np.int16) + off, 0, 255, out=blk, casting="unsafe")
    return img


def make_case(spec):
    kind = spec["kind"]
    h, w = spec["shape"]
    rng = np.random.default_rng(spec.get("seed", 0))
    if kind == "noise":
        return rng.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == "flat":
        return np.full((h, w), spec["value"], dtype=np.uint8)
    if kind == "checker":  # alternating 0/255 pixels
        yy, xx = np.mgrid[0:h, 0:w]
        return (((yy + xx) & 1) * 255).astype(np.uint8)
    if kind == "blockalt":  # alternating 0/255 8x8 blocks: maximum DC differences
        yy, xx = np.mgrid[0:h, 0:w]
        return ((((yy // 8) + (xx // 8)) & 1) * 255).astype(np.uint8)
    if kind == "impulse":  # sparse impulses: long zero runs (ZRL) in the AC scan
        img = np.full((h, w), 128, dtype=np.uint8)
        n = max(1, (h * w) // 97)
        img.reshape(-1)[rng.choice(h * w, n, replace=False)] = rng.integers(0, 256, n)
        return img
    if kind == "binary":  # random 0/255 pixels: largest AC magnitudes
        return (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    if kind == "synthetic":
        return synthetic_image(h, w, spec.get("seed", 0))
    if kind == "ramp":
        yy, xx = np.mgrid[0:h, 0:w]
        return ((yy * 3 + xx * 5) % 256).astype(np.uint8)
    raise ValueError(kind)


# Cases whose reference streams are committed in tests/golden/streams.npz.
ODD_CASES = {
    "pad_37x51": {"kind": "noise", "shape": (37, 51), "seed": 1, "qualities": [75, 50, 5], "auto": True},
    "one_px": {"kind": "noise", "shape": (1, 1), "seed": 2, "qualities": [50]},
    "row_1x17": {"kind": "noise", "shape": (1, 17), "seed": 3, "qualities": [50, 90]},
    "col_19x1": {"kind": "noise", "shape": (19, 1), "seed": 4, "qualities": [50]},
    "tiny_3x5": {"kind": "noise", "shape": (3, 5), "seed": 5, "qualities": [50, 10]},
    "tiny_7x9": {"kind": "ramp", "shape": (7, 9), "qualities": [50]},
    "empty_0x8": {"kind": "noise", "shape": (0, 8), "seed": 6, "qualities": [50]},
    "exact_8x8": {"kind": "noise", "shape": (8, 8), "seed": 7, "qualities": [50, 95]},
    "wide_8x1048": {"kind": "synthetic", "shape": (8, 1048), "seed": 8, "qualities": [50]},
    "unaligned_100x100": {"kind": "synthetic", "shape": (100, 100), "seed": 9, "qualities": [50, 20]},
    "w_mult4_64x60": {"kind": "noise", "shape": (64, 60), "seed": 10, "qualities": [50]},
    "flat0_64x64": {"kind": "flat", "shape": (64, 64), "value": 0, "qualities": [50]},
    "flat255_40x72": {"kind": "flat", "shape": (40, 72), "value": 255, "qualities": [90]},
    "flat128_64x64": {"kind": "flat", "shape": (64, 64), "value": 128, "qualities": [50]},
    "checker_64x64": {"kind": "checker", "shape": (64, 64), "qualities": [50, 90]},
    "blockalt_64x128": {"kind": "blockalt", "shape": (64, 128), "qualities": [50, 95]},
    "impulse_128x128": {"kind": "impulse", "shape": (128, 128), "seed": 11, "qualities": [50, 90, 95], "auto": True},
    "binary_64x64": {"kind": "binary", "shape": (64, 64), "seed": 12, "qualities": [50, 90, 95]},
    "noise_256x256": {"kind": "noise", "shape": (256, 256), "seed": 13, "qualities": [90, 50, 10], "auto": True},
    "synthetic_256x512": {"kind": "synthetic", "shape": (256, 512), "seed": 14, "qualities": [90, 50, 10, 1, 99]},
    "synthetic_1024x1024": {"kind": "synthetic", "shape": (1024, 1024), "seed": 0, "qualities": [50]},
}
