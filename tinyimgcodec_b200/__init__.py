"""tinyimgcodec_b200 — B200-native (sm_100a) implementation of tinyimgcodec's codec path.

Public names mirror the reference package (tinyimgcodec/__init__.py:1-5): `encode`, `compress` (the encode
hot path) and `decode`, `decompress` (the GPU decoder).  `compress_c`
produces the stream variant of the reference's embedded C encoder (c/encode.c).
"""
from .codec import (DeviceBatchResult, Encoder, TicError, TicStreamError, compress, compress_batch, compress_c,
                    decode, decompress, decompress_batch, encode, get_encoder, parse_header)

__version__ = "0.1.0"
__all__ = ["encode", "decode", "compress", "decompress", "compress_batch", "decompress_batch", "compress_c",
           "parse_header", "Encoder", "get_encoder", "DeviceBatchResult", "TicError", "TicStreamError"]
