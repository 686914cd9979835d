"""TEST INFRASTRUCTURE ONLY — compiles the reference's Python package to bytecode inside a zip archive.

    python oracle/stage_ref_bytecode.py /root/reference oracle/_ref/py/tinyimgcodec_ref.zip

Outputs only: the reference's .py sources stay where they lie; the archive holds tinyimgcodec/<module>.pyc, which the
same interpreter imports as a sourceless package through zipimport (oracle/ref_harness.py).  A zip rather than loose
.pyc files because loose bytecode does not travel to the GPU box."""
import glob
import os
import py_compile
import sys
import tempfile
import zipfile

ref_root, out = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as td, zipfile.ZipFile(out, "w", zipfile.ZIP_STORED) as z:
    for f in sorted(glob.glob(os.path.join(ref_root, "tinyimgcodec", "*.py"))):
        name = os.path.basename(f)
        c = os.path.join(td, name + "c")
        py_compile.compile(f, cfile=c, dfile="reference/tinyimgcodec/" + name, doraise=True)
        z.write(c, "tinyimgcodec/" + name + "c")
print(out)
