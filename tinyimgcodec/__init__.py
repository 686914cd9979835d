"""Drop-in name for the reference package: `from tinyimgcodec import compress, encode`
(tinyimgcodec/__init__.py:1-5 of clysto/tinyimgcodec) resolves to the B200 encode path.

Only the encode hot path is rebuilt here; `decode` and `decompress` belong to
the reference's decoder and raise, pointing at it.
"""
from tinyimgcodec_b200 import compress, compress_batch, encode  # noqa: F401

__version__ = "0.0.1"
__all__ = ["encode", "decode", "compress", "decompress"]


def _decoder_out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(
            f"tinyimgcodec.{name} is the reference's decoder; this repo rebuilds only the encode path "
            "(use clysto/tinyimgcodec to decode the .img streams produced here)")
    fn.__name__ = name
    return fn


decode = _decoder_out_of_scope("decode")
decompress = _decoder_out_of_scope("decompress")
