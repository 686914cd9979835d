"""`tinyimgcodec.codec` entry points (tinyimgcodec/codec.py:26,46,133,167)."""
from tinyimgcodec_b200.codec import compress, decode, decompress, encode  # noqa: F401
