"""Drop-in name for the reference package: `from tinyimgcodec import encode, decode, compress, decompress`
(tinyimgcodec/__init__.py:1-5 of clysto/tinyimgcodec) resolves to the B200 path.
"""
from tinyimgcodec_b200 import compress, compress_batch, decode, decompress, decompress_batch, encode  # noqa: F401

__version__ = "0.0.1"
__all__ = ["encode", "decode", "compress", "decompress"]
