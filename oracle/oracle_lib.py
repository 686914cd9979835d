"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/libtic_oracle.so (tic_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtic_oracle.so")
_lib = None

STATUS = {0: "ok", 1: "category", 2: "field", 3: "quality0", 4: "capacity", 5: "empty_auto"}


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status} ({STATUS.get(status, '?')})")
        self.status = status


def build(force=False):
    src = os.path.join(_HERE, "tic_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.tico_compress.restype = ctypes.c_int64
        L.tico_compress.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.POINTER(ctypes.c_int)]
        L.tico_encode_coeffs.restype = ctypes.c_int
        L.tico_encode_coeffs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
        L.tico_max_out_bytes.restype = ctypes.c_int64
        L.tico_max_out_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
        L.tico_dct8_rows.restype = None
        L.tico_dct8_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.tico_quant_table.restype = ctypes.c_int
        L.tico_quant_table.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.tico_default_code.restype = ctypes.c_int
        L.tico_default_code.argtypes = [ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int)]
        L.tico_idct8_rows.restype = None
        L.tico_idct8_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.tico_parse_header.restype = ctypes.c_int
        L.tico_parse_header.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.POINTER(ctypes.c_int64)] * 3 + [
            ctypes.POINTER(ctypes.c_uint32)]
        L.tico_decompress.restype = ctypes.c_int
        L.tico_decompress.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_int64)]
        del u8p
        _lib = L
    return _lib


def _as_u8(image):
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("2-D image expected")
    return np.ascontiguousarray(image.astype(np.uint8, copy=False))


def compress(image, quality=50, auto_generate_huffman_table=False, le_flag_word=False):
    """Restatement of tinyimgcodec.codec.compress (codec.py:133-164) for uint8 input.  le_flag_word: the
    opt-in header form whose flag word the reference's own decoder can read (auto-table streams only)."""
    img = _as_u8(image)
    h, w = img.shape
    L = lib()
    cap = int(L.tico_max_out_bytes(h, w)) + (4096 if auto_generate_huffman_table else 0)
    if auto_generate_huffman_table:
        cap = cap * 2
    out = np.empty(cap, dtype=np.uint8)
    status = ctypes.c_int(0)
    n = L.tico_compress(img.ctypes.data, h, w, int(quality),
                        (2 if le_flag_word else 1) if auto_generate_huffman_table else 0,
                        out.ctypes.data, cap, ctypes.byref(status))
    if n < 0:
        raise OracleError(status.value)
    return out[:n].tobytes()


def encode(image, quality=50):
    """Restatement of tinyimgcodec.codec.encode (codec.py:26-43) for uint8 input."""
    img = _as_u8(image)
    h, w = img.shape
    nblk = ((h + 7) // 8) * ((w + 7) // 8) if h and w else 0
    dc = np.zeros(nblk, dtype=np.int32)
    ac = np.zeros((nblk, 63), dtype=np.int32)
    rc = lib().tico_encode_coeffs(img.ctypes.data, h, w, int(quality), dc.ctypes.data, ac.ctypes.data)
    if rc:
        raise OracleError(rc)
    return {"height": h, "width": w, "quality": quality, "dc": dc, "ac": ac}


def parse_header(data):
    """Restatement of tinyimgcodec.codec.parse_header's fixed part (codec.py:117-122): (h, w, quality, flag)."""
    buf = np.frombuffer(bytes(data[:16]), dtype=np.uint8)
    h, w, q = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    flag = ctypes.c_uint32(0)
    if lib().tico_parse_header(buf.ctypes.data, buf.size, ctypes.byref(h), ctypes.byref(w), ctypes.byref(q),
                               ctypes.byref(flag)):
        raise OracleError(4)
    return h.value, w.value, q.value, flag.value


def decompress(data, return_errors=False):
    """Restatement of tinyimgcodec.codec.decompress (codec.py:167-189) -> uint8 H x W."""
    h, w, _, _ = parse_header(data)
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros((h, w), dtype=np.uint8)
    nerr = ctypes.c_int64(0)
    rc = lib().tico_decompress(buf.ctypes.data, buf.size, out.ctypes.data, ctypes.byref(nerr))
    if rc:
        raise OracleError(rc)
    return (out, nerr.value) if return_errors else out


def idct8_rows(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).copy()
    assert x.shape[-1] == 8
    lib().tico_idct8_rows(x.ctypes.data, x.size // 8)
    return x


def dct8_rows(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).copy()
    assert x.shape[-1] == 8
    lib().tico_dct8_rows(x.ctypes.data, x.size // 8)
    return x


def quant_table(quality):
    qt = np.zeros(64, dtype=np.float64)
    rc = lib().tico_quant_table(int(quality), qt.ctypes.data)
    if rc:
        raise OracleError(rc)
    return qt.reshape(8, 8)


def default_code(is_ac, sym):
    code = ctypes.c_uint32(0)
    ln = ctypes.c_int(0)
    rc = lib().tico_default_code(int(is_ac), int(sym), ctypes.byref(code), ctypes.byref(ln))
    if rc:
        return None
    return format(code.value, f"0{ln.value}b")


# ---- the reference's own C encoder (oracle/_ref/encode, built from /root/reference/c by `make ref`) -------
REF_C_ENCODER = os.path.join(_HERE, "_ref", "encode")


def ref_c_available():
    return os.path.isfile(REF_C_ENCODER)


def ref_c_compress(img, qfactor="med"):
    """stdout of `encode <width> <height> <qfactor>` fed the raw rows of `img` from a FILE on stdin
    (c/encode.c:13-66) — the reference binary itself, unmodified."""
    import subprocess
    import tempfile
    import numpy as np
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    with tempfile.NamedTemporaryFile(suffix=".raw") as f:
        f.write(img.tobytes())
        f.flush()
        with open(f.name, "rb") as fin:
            return subprocess.run([REF_C_ENCODER, str(w), str(h), qfactor], stdin=fin, stdout=subprocess.PIPE,
                                  check=True).stdout
