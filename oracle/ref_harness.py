"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

The reference needs `bidict` and `bitarray` (requirements.txt:1-2); neither is
installed and there is no network, so oracle/standins/ provides the minimal API
(SURVEY.md Appendix E).  /root/reference exists only in the build container: it
does not travel to the GPU box, so anything that calls load_reference() must be
skipped when it is absent (tests do `pytest.importorskip`-style gating through
`reference_available()`).  Nothing in the product imports this file.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STANDINS = os.path.join(_HERE, "standins")
# the build container has the reference itself; the GPU box only has its bytecode (oracle/Makefile `refpy`:
# oracle/_ref/py/tinyimgcodec_ref.zip, a sourceless package — the encoder only, for CPU timing)
REFERENCE_ROOT = os.environ.get("TIC_REFERENCE_ROOT", "/root/reference")
_STAGED_ZIP = os.path.join(_HERE, "_ref", "py", "tinyimgcodec_ref.zip")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "tinyimgcodec", "codec.py")) and os.path.isfile(_STAGED_ZIP):
    REFERENCE_ROOT = _STAGED_ZIP   # zipimport: tinyimgcodec/*.pyc inside


def reference_available():
    """The reference SOURCES (needed by the tests that also use its data/ and tests/ files)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tinyimgcodec", "codec.py"))


def reference_python_available():
    """The reference's Python package, as source or as bytecode: enough to call compress() / decompress()."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tinyimgcodec", "codec.py")) or REFERENCE_ROOT == _STAGED_ZIP


def load_reference():
    """Return the reference `tinyimgcodec` package (tinyimgcodec/__init__.py:1-5).

    The repo root also holds a drop-in package called `tinyimgcodec`; to make sure
    the REFERENCE one is imported, it is loaded under its own name from
    REFERENCE_ROOT with the repo root temporarily removed from sys.path, then
    re-registered as `tinyimgcodec_reference` so both can coexist in a process.
    """
    if "tinyimgcodec_reference" in sys.modules:
        return sys.modules["tinyimgcodec_reference"]
    if not reference_python_available():
        raise ImportError(f"reference not present at {REFERENCE_ROOT}")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k == "tinyimgcodec" or k.startswith("tinyimgcodec.")}
    for k in saved_mods:
        del sys.modules[k]
    try:
        repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path = [_STANDINS, REFERENCE_ROOT] + [
            p for p in sys.path
            if os.path.abspath(p or os.getcwd()) != repo_root
        ]
        ref = importlib.import_module("tinyimgcodec")
        assert os.path.abspath(ref.__file__).startswith(os.path.abspath(REFERENCE_ROOT))
        ref_mods = {k: v for k, v in sys.modules.items()
                    if k == "tinyimgcodec" or k.startswith("tinyimgcodec.")}
    finally:
        sys.path = saved_path
        for k in list(sys.modules):
            if k == "tinyimgcodec" or k.startswith("tinyimgcodec."):
                del sys.modules[k]
        sys.modules.update(saved_mods)
    for k, v in ref_mods.items():
        sys.modules[k.replace("tinyimgcodec", "tinyimgcodec_reference", 1)] = v
    return ref_mods["tinyimgcodec"]


def reference_psnr():
    """The reference's own PSNR helper, verbatim (tests/psnr.py:5-9, uint8-wrapping)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "tinyimgcodec_reference_psnr", os.path.join(REFERENCE_ROOT, "tests", "psnr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.psnr
