"""ctypes binding of libtinyimgcodec_cuda.so (include/tinyimgcodec_cuda.h).

There is no CPU fallback: if the library is missing or no B200-class device is usable,
loading / handle creation raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TIC_LIB_PATH") or os.path.join(_HERE, "libtinyimgcodec_cuda.so")   # override: kernel experiments

TIC_OK = 0
TIC_E_INVALID = -1
TIC_E_CUDA = -2
TIC_E_QUALITY = -3
TIC_E_CAPACITY = -4
TIC_E_CATEGORY = -5
TIC_E_UNSUPPORTED = -6
TIC_E_TABLE = -7
TIC_E_STREAM = -8
TIC_FLAG_AUTO_HUFFMAN = 1
TIC_FLAG_C_VARIANT = 2
TIC_FLAG_AUTO_LE_FLAG = 4
TIC_FLAG_DEBUG_ALL_EXACT = 8
TIC_STATUS_CATEGORY = 1
TIC_STATUS_TABLE = 2
TIC_STATUS_LONGCODE = 4
TIC_DFLAG_ACCEPT_BE_FLAG = 1
TIC_DFLAG_EXACT_ONLY = 2
TIC_DFLAG_FUSED = 4
TIC_DFLAG_NO_EARLY_STOP = 8
TIC_DFLAG_SYNC_ROUNDS = 16
TIC_DSTATUS_HEADER = 1
TIC_DSTATUS_CODE = 2
TIC_DSTATUS_TRUNCATED = 4
TIC_DSTATUS_TABLE = 8
TIC_DSTATUS_QUALITY = 16
TIC_DSTATUS_RANGE = 32
AUTO_HEADER_SLACK = 1664

# every symbol include/tinyimgcodec_cuda.h declares
EXPORTS = ["tic_version", "tic_create", "tic_destroy", "tic_last_error", "tic_max_out_bytes",
           "tic_num_blocks", "tic_encode_batch", "tic_encode_finish", "tic_encode_coeffs",
           "tic_compress_host", "tic_last_stats", "tic_last_guard_misses", "tic_parse_header", "tic_decode_batch", "tic_decode_finish",
           "tic_decompress_host", "tic_decode_coeffs", "tic_decode_stats"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m tinyimgcodec_b200.build` "
            "(nvcc, sm_100a).  tinyimgcodec_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32
    L.tic_version.restype = ctypes.c_char_p
    L.tic_version.argtypes = []
    L.tic_create.restype = ctypes.c_int
    L.tic_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.tic_destroy.restype = ctypes.c_int
    L.tic_destroy.argtypes = [vp]
    L.tic_last_error.restype = ctypes.c_char_p
    L.tic_last_error.argtypes = [vp]
    L.tic_max_out_bytes.restype = i64
    L.tic_max_out_bytes.argtypes = [i32, i32]
    L.tic_num_blocks.restype = i64
    L.tic_num_blocks.argtypes = [i32, i32]
    L.tic_encode_batch.restype = ctypes.c_int
    L.tic_encode_batch.argtypes = [vp, vp, vp, vp, i32, i32, u32, vp, i64, vp, vp, vp, vp]
    L.tic_encode_finish.restype = ctypes.c_int
    L.tic_encode_finish.argtypes = [vp, vp, ctypes.POINTER(i64)]
    L.tic_encode_coeffs.restype = ctypes.c_int
    L.tic_encode_coeffs.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.tic_compress_host.restype = ctypes.c_int
    L.tic_compress_host.argtypes = [vp, vp, i32, i32, i32, u32, vp, i64, ctypes.POINTER(i64),
                                    ctypes.POINTER(i32)]
    L.tic_last_stats.restype = ctypes.c_int
    L.tic_last_stats.argtypes = [vp, ctypes.POINTER(i64)]
    L.tic_last_guard_misses.restype = i64
    L.tic_last_guard_misses.argtypes = [vp]
    L.tic_parse_header.restype = ctypes.c_int
    L.tic_parse_header.argtypes = [vp, i64, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32),
                                   ctypes.POINTER(u32)]
    L.tic_decode_batch.restype = ctypes.c_int
    L.tic_decode_batch.argtypes = [vp, vp, vp, vp, vp, i32, u32, vp, vp, vp]
    L.tic_decode_finish.restype = ctypes.c_int
    L.tic_decode_finish.argtypes = [vp, vp]
    L.tic_decompress_host.restype = ctypes.c_int
    L.tic_decompress_host.argtypes = [vp, vp, i64, u32, vp, i64, ctypes.POINTER(i32)]
    L.tic_decode_coeffs.restype = ctypes.c_int
    L.tic_decode_coeffs.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    L.tic_decode_stats.restype = ctypes.c_int
    L.tic_decode_stats.argtypes = [vp, ctypes.POINTER(i64)]
    _lib = L
    return L
