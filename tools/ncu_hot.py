"""Hot SASS instructions and stall reasons from an ncu source page:
   ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], rows[start + 1:]
ix = {h: i for i, h in enumerate(hdr)}
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError, KeyError):
        return 0.0


S = "Warp Stall Sampling (All Samples)"
tot = sum(f(r, S) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f"{k:28s} {v:10.0f} {v / max(tot, 1) * 100:5.1f}%")
cum = 0
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -f(r, S))[:top_n]:
    st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"#{r[-1]:5d} {f(r, S):7.0f} exec {f(r, 'Instructions Executed'):8.0f}  {r[ix['Source']].strip()[:64]:64s} {st}")
