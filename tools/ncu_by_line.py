"""Join `ncu --page source --csv` (per-SASS counts) with `nvdisasm -g` line info of the same binary and
aggregate executed warp-instructions / stall samples per CUDA source line and per source function."""
import csv, re, subprocess, sys, os, collections, tempfile
rep_csv, lib, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
td = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
dis = []
for cubin in sorted(os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")):   # the one that holds the kernel
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    if kernel in out:
        dis = out.splitlines()
        break
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(rep_csv)))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
if len(body) != len(lines):
    print(f"WARNING: {len(body)} profiled instructions vs {len(lines)} disassembled — binaries differ?")
n = min(len(body), len(lines))
ex = collections.Counter(); sm = collections.Counter()
for i in range(n):
    e = float(body[i][col["Instructions Executed"]] or 0); s = float(body[i][col["# Samples"]] or 0)
    ex[lines[i]] += e; sm[lines[i]] += s
tot_e, tot_s = sum(ex.values()), sum(sm.values())
src = {}
def text(f, l):
    if f not in src:
        p = [os.path.join(d, f) for d in (os.environ.get("TIC_SRC_DIR", "tinyimgcodec_b200/csrc"),) if os.path.exists(os.path.join(d, f))]
        src[f] = open(p[0]).read().splitlines() if p else []
    return src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
print(f"total warp-inst {tot_e:.3e}, samples {tot_s:.0f}")
print(f"top {top} lines by executed instructions:")
for (k, e) in ex.most_common(top):
    if k is None: continue
    print(f"  {e/tot_e*100:5.1f}% inst {sm[k]/tot_s*100:5.1f}% samp  {k[0]}:{k[1]:<4d} {text(*k)}")
print(f"\ntop {top//2} lines by stall samples:")
for (k, s) in sm.most_common(top // 2):
    if k is None: continue
    print(f"  {s/tot_s*100:5.1f}% samp {ex[k]/tot_e*100:5.1f}% inst  {k[0]}:{k[1]:<4d} {text(*k)}")
# per-function aggregation: nearest preceding "__device__"/"__global__" definition line in the file
def func_of(f, l):
    text(f, 1)
    for j in range(min(l, len(src[f])) - 1, -1, -1):
        m = re.match(r"\s*(?:template.*>\s*)?(?:__device__|__global__).*?(\w+)\s*\(", src[f][j])
        if m: return m.group(1)
        m = re.match(r"^(\w[\w\s\*&:<>]*?)\b(\w+)\s*\(.*\)\s*\{?\s*$", src[f][j])
    return f
fe = collections.Counter(); fs = collections.Counter()
for k in ex:
    if k is None: continue
    fn = func_of(*k); fe[fn] += ex[k]; fs[fn] += sm[k]
print("\nby function:")
for fn, e in fe.most_common(20):
    print(f"  {e/tot_e*100:5.1f}% inst {fs[fn]/tot_s*100:5.1f}% samp  {fn}")
