"""Summarise `ncu --page source --csv` output: stall totals, hottest SASS, instruction mix."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot_samples = sum(num(r, "# Samples") for r in body)
tot_inst = sum(num(r, "Instructions Executed") for r in body)
print(f"instructions(static)={len(body)} warp-inst executed={tot_inst:.3e} samples={tot_samples:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r, s) for r in body) for s in stalls}
print("stall totals (all samples):", ", ".join(f"{k[6:]}={v/tot_samples*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print(f"\ntop {top} by samples:")
order = sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:top]
for i in order:
    r = body[i]
    st = sorted(((s[6:], num(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f"  #{i:5d} {num(r,'# Samples')/tot_samples*100:5.2f}%  exec={num(r,'Instructions Executed'):.2e} thr/inst={num(r,'Avg. Threads Executed'):4.1f}  {r[col['Source']].strip()[:70]:70s} {st}")
mix = collections.Counter()
for r in body:
    op = r[col["Source"]].strip().split()
    op = [o for o in op if not o.startswith("@")]
    if op: mix[op[0].split(".")[0]] += num(r, "Instructions Executed")
print("\ninstruction mix (warp-inst executed):")
for k, v in mix.most_common(28):
    print(f"  {k:10s} {v:.3e} {v/tot_inst*100:5.1f}%")
# cumulative executed by static position deciles
acc = 0; marks = []
for i, r in enumerate(body):
    acc += num(r, "Instructions Executed")
    marks.append(acc)
print("\nexecuted share by static position (every 5%):")
n = len(body)
print("  " + " ".join(f"{marks[min(n-1,int(n*p/20))]/tot_inst*100:.0f}" for p in range(1, 21)))
