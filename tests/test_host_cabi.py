"""CPU: the C-ABI library loads, exports every symbol include/tinyimgcodec_cuda.h declares, and
its pure-host entry points agree with the oracle.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tinyimgcodec_b200 import build
    build.build()
    from tinyimgcodec_b200 import _lib
    return _lib.load()


def test_exports_match_header(lib):
    from tinyimgcodec_b200 import _lib
    with open(os.path.join(ROOT, "include", "tinyimgcodec_cuda.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(tic_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_host_arithmetic(lib):
    from oracle import oracle_lib as O
    for h, w in [(0, 8), (1, 1), (8, 8), (37, 51), (512, 512), (4320, 7680), (32768, 32768)]:
        nblk = ((h + 7) // 8) * ((w + 7) // 8) if h and w else 0
        assert lib.tic_num_blocks(h, w) == nblk
        assert lib.tic_max_out_bytes(h, w) == O.lib().tico_max_out_bytes(h, w)


def test_parse_header_is_pure_host(lib, golden):
    from tinyimgcodec_b200 import parse_header
    data = golden.streams["img_lenna_q50"]
    h, w, q, f = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_uint32())
    buf = np.frombuffer(data, dtype=np.uint8)
    assert lib.tic_parse_header(buf.ctypes.data, buf.size, ctypes.byref(h), ctypes.byref(w), ctypes.byref(q),
                                ctypes.byref(f)) == 0
    assert (h.value, w.value, q.value, f.value) == (512, 512, 50, 0) == parse_header(data)
    assert lib.tic_parse_header(buf.ctypes.data, 15, None, None, None, None) != 0
    import struct
    with pytest.raises(struct.error):
        parse_header(data[:15])


def test_null_handle_is_rejected_everywhere(lib):
    """Every entry point that takes a handle returns TIC_E_INVALID for NULL before touching CUDA."""
    from tinyimgcodec_b200 import _lib
    z = ctypes.c_void_p(None)
    i64 = (ctypes.c_int64 * 12)()
    assert lib.tic_destroy(z) == _lib.TIC_E_INVALID
    assert lib.tic_encode_finish(z, None, None) == _lib.TIC_E_INVALID
    assert lib.tic_last_stats(z, i64) == _lib.TIC_E_INVALID
    assert lib.tic_decode_batch(z, None, None, None, None, 0, 0, None, None, None) == _lib.TIC_E_INVALID
    assert lib.tic_decode_finish(z, None) == _lib.TIC_E_INVALID
    assert lib.tic_decode_stats(z, i64) == _lib.TIC_E_INVALID
    assert lib.tic_decode_coeffs(z, None, None, 8, 8, 50, 0, None, None) == _lib.TIC_E_INVALID
    assert lib.tic_decompress_host(z, None, 0, 0, None, 0, None) == _lib.TIC_E_INVALID
    assert lib.tic_last_error(z) == b"null handle"


def test_no_cpu_fallback(lib):
    """Without a GPU the product must fail loudly, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tinyimgcodec_b200 as tic
    with pytest.raises(tic.TicError):
        tic.compress(np.zeros((8, 8), np.uint8))
    with pytest.raises(tic.TicError):
        tic.decompress(bytes(16))
    h = ctypes.c_void_p()
    assert lib.tic_create(0, ctypes.byref(h)) != 0


def test_argument_errors_before_any_gpu_work():
    import struct
    from tinyimgcodec_b200.codec import _as_u8_image, _check_quality
    with pytest.raises(ZeroDivisionError):
        _check_quality(0)
    with pytest.raises(struct.error):
        _check_quality(50.0)
    with pytest.raises(KeyError):
        _check_quality(100)
    with pytest.raises(ValueError):
        _as_u8_image(np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        _as_u8_image(np.full((4, 4), 300))
    img, h, w = _as_u8_image(np.full((4, 6), 7.9))
    assert img.dtype == np.uint8 and (h, w) == (4, 6) and int(img[0, 0]) == 7


def test_product_does_not_import_oracle():
    """The product path must not reference oracle/ in any way."""
    pkg = os.path.join(ROOT, "tinyimgcodec_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, fn)
