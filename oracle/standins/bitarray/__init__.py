"""TEST INFRASTRUCTURE ONLY — stand-in for the third-party `bitarray` C extension.

Implements exactly the surface tinyimgcodec/bitbuffer.py:1-72 uses, with the
semantics of bitarray(endian="big"): MSB-first bits, tobytes() zero-pads the last
byte.  Pure Python + numpy; used only through oracle/ref_harness.py to run the
reference unmodified in the build container.
"""
import numpy as np


class bitarray:
    def __init__(self, initial=None, endian="big"):
        if endian != "big":
            raise NotImplementedError("stand-in supports endian='big' only")
        self._b = bytearray()
        if initial is not None:
            self.extend(initial)

    # -- construction -----------------------------------------------------
    def frombytes(self, data):
        bits = np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8))
        self._b.extend(bits.tobytes())

    def extend(self, x):
        if isinstance(x, bitarray):
            self._b.extend(x._b)
        elif isinstance(x, str):
            for ch in x:
                if ch == "0":
                    self._b.append(0)
                elif ch == "1":
                    self._b.append(1)
                else:
                    raise ValueError(f"expected '0' or '1', got {ch!r}")
        else:
            for v in x:
                v = int(v)
                if v not in (0, 1):
                    raise ValueError(f"bit must be 0 or 1, got {v}")
                self._b.append(v)

    # -- export -----------------------------------------------------------
    def tobytes(self):
        if not self._b:
            return b""
        return np.packbits(np.frombuffer(bytes(self._b), dtype=np.uint8)).tobytes()

    def to01(self):
        return "".join("1" if v else "0" for v in self._b)

    def invert(self):
        self._b = bytearray(1 - v for v in self._b)

    # -- container protocol -----------------------------------------------
    def __len__(self):
        return len(self._b)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            out = bitarray()
            out._b = self._b[idx]
            return out
        return self._b[idx]

    def __eq__(self, other):
        return isinstance(other, bitarray) and self._b == other._b

    def __repr__(self):
        return f"bitarray('{self.to01()}')"
