"""TEST INFRASTRUCTURE ONLY — stand-in for bitarray.util (see __init__.py)."""
from . import bitarray


def ba2int(a):
    if len(a) == 0:
        raise ValueError("non-empty bitarray expected")
    n = 0
    for v in a._b:
        n = (n << 1) | v
    return n


def int2ba(n, length=None):
    n = int(n)
    if n < 0:
        raise OverflowError("unsigned integer expected")
    bits = bin(n)[2:] if n else "0"
    if length is not None:
        if len(bits) > length:
            raise OverflowError(f"int too big to convert: {n} does not fit in {length} bits")
        bits = bits.rjust(length, "0")
    return bitarray(bits)
