"""GPU parity tests: the CUDA path, called through the C ABI (ctypes, via the host mirror in
tinyimgcodec_b200/codec.py), against the committed golden vectors of the reference and against
the CPU oracle (oracle/tic_oracle.c) on the same seeded inputs.  Bit-exact: integer coefficients
and byte streams must be identical."""
import hashlib

import numpy as np
import pytest

from oracle import oracle_lib as O
from tests.cases import make_case, synthetic_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tic():
    import tinyimgcodec_b200 as m
    return m


def _first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n


def _assert_same(got, want, what):
    assert got == want, f"{what}: len {len(got)} vs {len(want)}, first difference at byte {_first_diff(got, want)}"


def test_native_library_loaded(tic):
    import tinyimgcodec_b200._lib as L
    lib = L.load()
    assert b"sm_100a" in lib.tic_version()
    with open("/proc/self/maps") as f:
        assert "libtinyimgcodec_cuda.so" in f.read()


def test_lenna_kat(tic, golden):
    out = tic.compress(golden.images["lenna"])
    assert len(out) == 20765
    assert hashlib.sha256(out).hexdigest() == "4596d8bb0d5577d2d4321e5e2ff8c086ce8fbd313b00d4879d4c65baaa8048a9"


def test_golden_streams(tic, golden):
    n = 0
    for key, img, q, want in golden.stream_cases(auto=False):
        _assert_same(tic.compress(img, q), want, key)
        n += 1
    assert n >= 50


def test_golden_error_cases(tic, golden):
    for key, img, q, exc in golden.error_cases():
        with pytest.raises(KeyError):
            tic.compress(img, q)


def test_golden_coefficients(tic, golden):
    lenna = golden.images["lenna"]
    for q in (50, 90):
        e = tic.encode(lenna, q)
        assert e["dc"].dtype == np.int32 and e["ac"].dtype == np.int32
        assert np.array_equal(e["dc"], golden.coeffs[f"lenna_q{q}_dc"])
        assert np.array_equal(e["ac"], golden.coeffs[f"lenna_q{q}_ac"].astype(np.int32))
        assert (e["height"], e["width"], e["quality"]) == (512, 512, q)


def test_subset_images_batch_equals_reference_streams(tic, golden):
    """BASELINE config 2 shape (data/*.gif as one batch), on the committed subset."""
    names = sorted(golden.images)
    outs = tic.compress_batch([golden.images[n] for n in names], 50)
    for n, out in zip(names, outs):
        kat = golden.kat["q50"][n]
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"]), n


def test_all_50_gifs_one_batch_equals_reference(tic, golden):
    """BASELINE config 2 in full: all 50 data/*.gif as ONE batch, every stream checked against the
    reference's own bytes (size + sha256 recorded by oracle/gen_golden.py from the unmodified reference)."""
    gifs = golden.all_gifs()
    names = sorted(gifs)
    assert len(names) == 50
    outs = tic.compress_batch([gifs[n] for n in names], 50)
    for n, out in zip(names, outs):
        kat = golden.kat["q50"][n]
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"]), n


@pytest.mark.parametrize("q", [1, 5, 10, 20, 35, 49, 50, 51, 65, 80, 90, 95])
def test_random_shapes_vs_oracle(tic, q):
    rng = np.random.default_rng(1000 + q)
    for _ in range(6):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        kind = ["noise", "synthetic", "binary", "impulse", "checker", "ramp"][int(rng.integers(0, 6))]
        img = make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30))})
        try:
            want = O.compress(img, q)
        except O.OracleError:
            with pytest.raises(KeyError):
                tic.compress(img, q)
            continue
        _assert_same(tic.compress(img, q), want, f"{kind} {h}x{w} q{q}")
        e, eo = tic.encode(img, q), O.encode(img, q)
        assert np.array_equal(e["dc"], eo["dc"]) and np.array_equal(e["ac"], eo["ac"])


def test_all_qualities_vs_oracle(tic):
    img = synthetic_image(96, 160, seed=5)
    noise = make_case({"kind": "noise", "shape": (64, 72), "seed": 77})
    for q in range(1, 100):
        for im in (img, noise):
            try:
                want = O.compress(im, q)
            except O.OracleError:
                with pytest.raises(KeyError):
                    tic.compress(im, q)
                continue
            _assert_same(tic.compress(im, q), want, f"q{q}")


def test_ragged_batch_vs_oracle(tic):
    rng = np.random.default_rng(7)
    shapes = [(512, 512), (8, 8), (1, 1), (0, 8), (37, 51), (1024, 1024), (64, 2048), (3, 5), (1032, 520),
              (16, 16), (129, 1025), (256, 8)]
    imgs = []
    for i, (h, w) in enumerate(shapes):
        kind = ["synthetic", "noise", "impulse"][i % 3]
        imgs.append(make_case({"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30))}))
    outs = tic.compress_batch(imgs, 50)
    assert len(outs) == len(imgs)
    for im, out in zip(imgs, outs):
        _assert_same(out, O.compress(im, 50), f"batch {im.shape}")


def test_1024_batch_vs_oracle(tic):
    """BASELINE config 4 shape at a size the oracle finishes in seconds: 24 x 1024^2."""
    imgs = [synthetic_image(1024, 1024, seed=i) for i in range(24)]
    for q in (50, 90, 10):
        outs = tic.compress_batch(imgs, q)
        for i, (im, out) in enumerate(zip(imgs, outs)):
            _assert_same(out, O.compress(im, q), f"1024^2 seed {i} q{q}")


def test_8k_image_vs_oracle(tic):
    """BASELINE config 3: one 7680x4320 image."""
    img = synthetic_image(4320, 7680, seed=0)
    _assert_same(tic.compress(img, 50), O.compress(img, 50), "8K q50")


def _device_stream(tic, d_img, q):
    """One image resident in HBM through the batch C ABI; returns the stream bytes."""
    import torch
    res = tic.get_encoder(0).encode_batch_device([d_img], q).finish()
    off, size = int(res.offsets[0]), int(res.sizes[0])
    got = res.out[off: off + size].cpu().numpy().tobytes()
    del res
    torch.cuda.empty_cache()
    return got


def test_32k_image_quality_sweep(tic):
    """BASELINE config 5: one 32768 x 32768 image (16.8 M blocks, 131,072 tiles in ONE stream — the
    scan / compaction stress) at the reference's benchmarked qualities (tests/benchmark.py:13).
    The extremes and q50 are compared byte for byte with the C oracle (about 20 s of host time each); the other
    qualities through a size-independent property: blocks are coded in raster order with a running DC
    predictor and no resets, so the stream of the top 1024 rows alone is, header aside, a bit-prefix of
    the full image's stream."""
    import torch
    from tests.cases import big_synthetic
    img = big_synthetic(32768, 32768, seed=5)
    d_img = torch.from_numpy(img).cuda()
    for q in (90, 50, 5):
        _assert_same(_device_stream(tic, d_img, q), O.compress(img, q), f"32768^2 q{q}")
    top = np.ascontiguousarray(img[:1024])
    for q in (80, 20, 10):
        full, part = _device_stream(tic, d_img, q), O.compress(top, q)
        assert full[:4] == (32768).to_bytes(4, "little") and full[4:16] == part[4:16]
        assert full[16: len(part) - 1] == part[16:-1], f"prefix property q{q}"   # last byte: padding bits
        assert len(full) > 30 * len(part)


def test_stream_longer_than_2_pow_32_bits_vs_oracle(tic):
    """Uniform noise at 32768^2, quality 90: more than 2^32 bits in one stream (64-bit bit positions)."""
    import torch
    img = np.random.default_rng(99).integers(0, 256, (32768, 32768), dtype=np.uint8)
    got = _device_stream(tic, torch.from_numpy(img).cuda(), 90)
    assert len(got) * 8 > (1 << 32)
    _assert_same(got, O.compress(img, 90), "noise 32768^2 q90")


def test_noise_high_quality_stress_vs_oracle(tic):
    """Uniform noise: worst case for scan/pack (6+ bits per pixel at q90) and for the tie guard."""
    img = make_case({"kind": "noise", "shape": (1024, 2048), "seed": 3})
    for q in (90, 50):
        _assert_same(tic.compress(img, q), O.compress(img, q), f"noise q{q}")
    flat = np.full((1024, 1024), 77, np.uint8)
    _assert_same(tic.compress(flat, 50), O.compress(flat, 50), "flat")


def test_device_tensor_batch_and_stats(tic):
    """(N, H, W) tensor resident in HBM -> streams in HBM (the bench's call), plus the counters of
    tic_last_stats."""
    import torch
    imgs = np.stack([synthetic_image(256, 384, seed=s) for s in range(9)])
    enc = tic.get_encoder(0)
    res = enc.encode_batch_device(torch.from_numpy(imgs).cuda(), 50).finish()
    for im, out in zip(imgs, res.to_bytes()):
        _assert_same(out, O.compress(im, 50), "tensor batch")
    offs = res.offsets.cpu().numpy()
    assert np.all(offs % 16 == 0) and np.all(np.diff(offs) > 0)          # dense, 16-byte aligned, in order
    st = enc.stats()
    assert st["launches"] == 4 and st["blocks"] == 9 * 32 * 48 and st["tiles"] == 9 * 12   # small (N,H,W) batch: prep, encode, fused scan, compact
    assert st["timed_batches"] == 1 and st["encode_kernel_ms_sum"] > 0 and st["compact_kernel_ms_sum"] > 0


def test_unaligned_pixel_pointers_and_output_alignment(tic):
    """Pixels at odd device addresses take the byte-wise load path (and the worklist for every halo DC); the
    output buffer must be 16-byte aligned (streams start on 16-byte boundaries)."""
    import torch
    enc = tic.get_encoder(0)
    imgs = [synthetic_image(64, 128, seed=3), make_case({"kind": "noise", "shape": (40, 72), "seed": 8}),
            np.full((256, 256), 255, np.uint8), synthetic_image(37, 51, seed=5)]
    raw = torch.empty(sum(im.size for im in imgs) + 64, dtype=torch.uint8, device="cuda")
    views, pos = [], 1
    for im in imgs:
        v = raw[pos: pos + im.size].view(im.shape)
        v.copy_(torch.from_numpy(im))
        assert v.data_ptr() % 8 != 0
        views.append(v)
        pos += im.size + (3 if (pos + im.size) % 8 == 5 else 2)   # keep every start off the 8-byte grid
        pos += 1 if pos % 8 == 0 else 0
    for q in (50, 90):
        outs = enc.encode_batch_device(views, q).finish().to_bytes()
        for im, out in zip(imgs, outs):
            _assert_same(out, O.compress(im, q), f"unaligned {im.shape} q{q}")
    buf = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    with pytest.raises(tic.TicError) as ei:
        enc.encode_batch_device(views[:1], 50, out=buf[4:])
    assert ei.value.code == -1


def test_output_capacity_is_respected(tic):
    """TIC_E_CAPACITY: a too-small d_out is reported, and nothing past out_capacity is written."""
    import torch
    img = make_case({"kind": "noise", "shape": (512, 512), "seed": 21})
    need = len(O.compress(img, 90))
    enc = tic.get_encoder(0)
    d_img = torch.from_numpy(img).cuda()
    for cap in (need // 3, need - 5, 64):
        cap16 = cap & ~15
        buf = torch.full((cap16 + 4096,), 0xA5, dtype=torch.uint8, device="cuda")
        with pytest.raises(tic.TicError) as ei:
            enc.encode_batch_device([d_img], 90, out=buf[:cap16]).finish()
        assert ei.value.code == -4
        assert bool((buf[cap16:] == 0xA5).all()), f"write past capacity {cap16}"
    buf = torch.full((((need + 3) & ~3) + 4096,), 0xA5, dtype=torch.uint8, device="cuda")
    res = enc.encode_batch_device([d_img], 90, out=buf[: (need + 3) & ~3]).finish()       # exactly enough
    assert res.to_bytes()[0] == O.compress(img, 90) and bool((buf[(need + 3) & ~3:] == 0xA5).all())


def test_error_behaviour_matches_reference(tic):
    import struct
    img = np.zeros((16, 16), np.uint8)
    with pytest.raises(ValueError):
        tic.compress(np.zeros((2, 3, 4), np.uint8))            # codec.py:27
    with pytest.raises(ZeroDivisionError):
        tic.compress(img, 0)                                   # utils.py:50
    with pytest.raises(struct.error):
        tic.compress(img, 50.0)                                # codec.py:103
    with pytest.raises(KeyError):
        tic.compress(img, 100)                                 # huffman.py:62 via utils.py:53
    assert isinstance(tic.compress(img.astype(np.int64)), bytes)


def test_dropin_package_names(tic):
    import tinyimgcodec
    from tinyimgcodec.codec import compress, encode
    img = synthetic_image(64, 64, seed=1)
    assert tinyimgcodec.compress(img) == compress(img) == O.compress(img, 50)
    assert set(encode(img)) == {"height", "width", "quality", "dc", "ac"}


def test_golden_auto_table_streams(tic, golden):
    """auto_generate_huffman_table=True (codec.py:146-148): tables built on the device must reproduce
    the reference's heapq tree and its serialised header bit for bit."""
    n = 0
    for key, img, q, want in golden.stream_cases(auto=True):
        _assert_same(tic.compress(img, q, True), want, key)
        n += 1
    assert n >= 6
    for q, kat in golden.kat["lenna_auto"].items():
        out = tic.compress(golden.images["lenna"], int(q), auto_generate_huffman_table=True)
        assert (len(out), hashlib.sha256(out).hexdigest()) == (kat["size"], kat["sha256"])


def test_auto_table_random_vs_oracle(tic):
    rng = np.random.default_rng(4242)
    for q in (95, 90, 50, 10, 2):
        for kind in ("noise", "synthetic", "impulse", "flat", "binary"):
            h, w = int(rng.integers(1, 150)), int(rng.integers(1, 150))
            spec = {"kind": kind, "shape": (h, w), "seed": int(rng.integers(0, 1 << 30)), "value": 128}
            img = make_case(spec)
            try:
                want = O.compress(img, q, True)
            except O.OracleError as e:
                assert e.status == 2
                with pytest.raises(OverflowError):
                    tic.compress(img, q, True)
                continue
            _assert_same(tic.compress(img, q, True), want, f"auto {kind} {h}x{w} q{q}")


def test_auto_table_le_flag_word_vs_oracle(tic):
    """TIC_FLAG_AUTO_LE_FLAG: only the four flag bytes change (decodability by the reference decoder is checked
    in tests/test_oracle_vs_reference.py where the reference is mounted)."""
    img = synthetic_image(200, 312, seed=4)
    for q in (50, 10):
        plain, compat = tic.compress(img, q, True), tic.compress(img, q, True, le_flag_word=True)
        _assert_same(compat, O.compress(img, q, True, le_flag_word=True), f"le flag q{q}")
        assert plain[12:16] == bytes([0x80, 0, 0, 0]) and compat[12:16] == bytes([0, 0, 0, 0x80])
        assert plain[:12] == compat[:12] and plain[16:] == compat[16:]
    with pytest.raises(ValueError):
        tic.compress(img, 50, False, le_flag_word=True)


def test_auto_table_batch_vs_oracle(tic):
    imgs = [synthetic_image(1024, 1024, seed=3), make_case({"kind": "noise", "shape": (200, 312), "seed": 5}),
            np.full((64, 64), 128, np.uint8), synthetic_image(8, 8, seed=1), synthetic_image(520, 1032, seed=9)]
    outs = tic.compress_batch(imgs, 50, auto_generate_huffman_table=True)
    for im, out in zip(imgs, outs):
        _assert_same(out, O.compress(im, 50, True), f"auto batch {im.shape}")
    with pytest.raises(IndexError):
        tic.compress(np.zeros((0, 8), np.uint8), 50, True)


@pytest.mark.parametrize("chunk", [1, 3, 4])
def test_compress_batch_pinned_vs_oracle(tic, chunk):
    """SURVEY §8(f)1, reference caller encode.py:10-19: the pipelined host API (pinned H2D / encode / D2H in
    chunks) that produces bench.py's e2e number.  Every (offset, size) slice of the returned pinned buffer is
    the reference's stream, for N below, at and above the chunk size (partial last chunk, more chunks than
    buffers), twice in a row on the same encoder (buffer reuse)."""
    import torch
    enc = tic.get_encoder(0)
    h, w = 64, 96
    for n in sorted({1, max(chunk - 1, 1), chunk, 2 * chunk + 3, 5 * chunk + 1}):
        imgs = np.stack([synthetic_image(h, w, seed=100 * chunk + i) for i in range(n)])
        imgs[n // 2] = make_case({"kind": "noise", "shape": (h, w), "seed": n})   # one image far above the mean size
        h_images = torch.from_numpy(imgs).pin_memory()
        want = [O.compress(im, 50) for im in imgs]
        for rep in range(2):
            h_out, index = enc.compress_batch_pinned(h_images, 50, chunk=chunk, out_bytes_per_pixel=3.3)
            assert len(index) == n
            host = h_out.numpy()
            for i, (off, size) in enumerate(index):
                assert off % 16 == 0
                _assert_same(host[off: off + size].tobytes(), want[i], f"chunk={chunk} n={n} rep={rep} image {i}")


def test_compress_batch_pinned_capacity_error_names_the_knob(tic):
    """A chunk that compresses worse than out_bytes_per_pixel must fail loudly (TIC_E_CAPACITY), not return a
    truncated stream (codec.py:272)."""
    import torch
    enc = tic.get_encoder(0)
    imgs = np.stack([make_case({"kind": "noise", "shape": (64, 64), "seed": i}) for i in range(6)])
    with pytest.raises(Exception) as ei:
        enc.compress_batch_pinned(torch.from_numpy(imgs).pin_memory(), 95, chunk=4, out_bytes_per_pixel=0.05)
    assert "out_bytes_per_pixel" in str(ei.value) or "too small" in str(ei.value)


@pytest.mark.parametrize("q", [1, 50, 90, 95, 99])
def test_tie_guard_is_sound_all_coefficients_exact(tic, q):
    """SURVEY App. B (last bullet), VERDICT r1 weak #2: the guard band of the fast transform is ASSERTED, not
    argued.  With TIC_FLAG_DEBUG_ALL_EXACT every coefficient — flagged or not, live group or not — is recomputed
    by the float64 exact path (the reference's arithmetic, utils.py:32-37,53); `guard_misses` counts those whose
    exact value differs from the fast value WITHOUT having been flagged.  Must be 0 on adversarial content."""
    import torch
    enc = tic.get_encoder(0)
    h, w = 128, 256
    kinds = [{"kind": "binary", "shape": (h, w), "seed": 3}, {"kind": "checker", "shape": (h, w)},
             {"kind": "blockalt", "shape": (h, w)}, {"kind": "noise", "shape": (h, w), "seed": 5}]
    imgs = [make_case(k) for k in kinds] + [synthetic_image(h, w, seed=s) for s in range(4)]
    # blocks whose pixels follow the sign of one basis function each: the largest sum |pixel x basis| there is
    yy, xx = np.mgrid[0:h, 0:w]
    n = ((yy // 8) * (w // 8) + xx // 8) % 64
    u, v = n // 8, n % 8
    basis = np.cos((2 * (yy % 8) + 1) * u * np.pi / 16) * np.cos((2 * (xx % 8) + 1) * v * np.pi / 16)
    imgs.append(np.where(basis > 0, 255, 0).astype(np.uint8))
    d = torch.from_numpy(np.stack(imgs)).cuda()
    try:
        res = enc.encode_batch_device(d, q, debug_all_exact=True).finish()
    except KeyError:
        res = None   # q >= 97 on 0/255 content: category outside the fixed tables, like the reference (counted anyway)
    st = enc.stats()
    assert st["exact_items"] >= len(imgs) * (h // 8) * (w // 8) * 64      # every coefficient went through the exact path
    assert enc.guard_misses() == 0
    if res is not None:
        for got, im in zip(res.to_bytes(), imgs):
            _assert_same(got, O.compress(im, q), f"debug-all-exact stream q={q}")


def test_bench_images_guard_is_sound(tic):
    """The same assertion on 16 images of the benchmark's kind (1024 x 1024, BASELINE config 4) at q50."""
    import torch
    enc = tic.get_encoder(0)
    imgs = np.stack([synthetic_image(1024, 1024, seed=s) for s in range(16)])
    enc.encode_batch_device(torch.from_numpy(imgs).cuda(), 50, debug_all_exact=True).finish()
    assert enc.guard_misses() == 0
    assert enc.stats()["exact_items"] >= 16 * 16384 * 64


def test_back_to_back_batches_without_finish(tic):
    """ADVICE r1 (tic_encode.cu:943): two DIFFERENT batches enqueued back to back on a busy stream, one finish.
    Both must be correct (the pinned descriptor staging is a ring), and an error of the FIRST batch must still be
    reported by the finish that follows the second (sticky status)."""
    import torch
    enc = tic.get_encoder(0)
    a = [synthetic_image(40 + 8 * i, 64 + 16 * i, seed=i) for i in range(5)]
    b = [synthetic_image(96, 72 - 8 * i, seed=50 + i) for i in range(3)]
    da = [torch.from_numpy(x).cuda() for x in a]
    db = [torch.from_numpy(x).cuda() for x in b]
    big = torch.from_numpy(np.stack([synthetic_image(1024, 1024, seed=s) for s in range(8)])).cuda()
    for _ in range(3):
        enc.encode_batch_device(big, 50)            # keeps the stream busy while the next two are enqueued
    ra = enc.encode_batch_device(da, 50)
    rb = enc.encode_batch_device(db, 60)
    rb.finish()
    ra.total_bytes = int((ra.offsets + ra.sizes).max().item())
    for got, im in zip(ra.to_bytes(), a):
        _assert_same(got, O.compress(im, 50), "first of two back-to-back batches")
    for got, im in zip(rb.to_bytes(), b):
        _assert_same(got, O.compress(im, 60), "second of two back-to-back batches")
    # sticky error: the first batch overflows its (tiny) output buffer, the second is fine
    small_out = torch.empty(64, dtype=torch.uint8, device="cuda")
    enc.encode_batch_device(da, 50, out=small_out)
    rb2 = enc.encode_batch_device(db, 60)
    with pytest.raises(Exception) as ei:
        rb2.finish()
    assert "too small" in str(ei.value)
    enc.encode_batch_device(db, 60).finish()        # and the handle is usable again
