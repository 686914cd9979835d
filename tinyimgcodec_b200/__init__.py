"""tinyimgcodec_b200 — B200-native (sm_100a) implementation of tinyimgcodec's encode path.

Public names mirror the reference package (tinyimgcodec/__init__.py:1-5) for the path this
repo covers: `encode` and `compress`.  `decode` / `decompress` are outside the hot path
 and are not provided.  `compress_c` produces the stream variant of the reference's
embedded C encoder (c/encode.c).
"""
from .codec import (DeviceBatchResult, Encoder, TicError, compress, compress_batch, compress_c, encode,
                    get_encoder)

__version__ = "0.1.0"
__all__ = ["encode", "compress", "compress_batch", "compress_c", "Encoder", "get_encoder", "DeviceBatchResult", "TicError"]
