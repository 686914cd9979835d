"""`tinyimgcodec.codec` entry points of the encode path (tinyimgcodec/codec.py:26,133)."""
from tinyimgcodec_b200.codec import compress, encode  # noqa: F401
