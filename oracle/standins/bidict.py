"""TEST INFRASTRUCTURE ONLY — stand-in for the third-party `bidict` package.

The reference (clysto/tinyimgcodec) imports `bidict` (requirements.txt:1), which is
not installed in this image and cannot be fetched (no network).  The reference only
uses an insertion-ordered mapping plus `.inverse` (tinyimgcodec/constants.py:54,70,
tinyimgcodec/codec.py:88, tinyimgcodec/huffman.py:83,168), so this 10-line class is
enough to import and run the reference UNMODIFIED from /root/reference.
It is put on sys.path only by oracle/ref_harness.py; the product never imports it.
"""


class bidict(dict):
    @property
    def inverse(self):
        return {v: k for k, v in self.items()}
