"""SASS opcode histogram per kernel of libtinyimgcodec_cuda.so (static instruction counts; cuobjdump -sass).

    python tools/sass_histogram.py [lib.so] > profiles/<round>_sass_histogram.txt

What to look for (B200_PROFILING.md): UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDGSTS = cp.async, HMMA would be the legacy mma.sync path (absent)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "tinyimgcodec_b200/libtinyimgcodec_cuda.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["cu++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(hist, demangle)) if len(demangle) == len(hist) else {k: k for k in hist}
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "UTMALDG", "HMMA", "IMMA", "DFMA", "DADD", "DMUL")
for k, c in hist.items():
    total = sum(c.values())
    short = re.sub(r"\((?!int\)).*", "", names[k]).replace("(int)", "")
    print(f"== {short}   [{total} instructions]")
    print("   tensor/async: " + (", ".join(f"{op} {c[op]}" for op in KEY if c[op]) or "-"))
    print("   top: " + ", ".join(f"{op} {n}" for op, n in c.most_common(14)))
