"""Top CUDA source lines of a kernel by stall samples / executed instructions from an .ncu-rep."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and r and r[0].isdigit() and len(r) > max(si, ii):
        key = (int(r[0]), r[1][:120])
        a = data.setdefault(key, [0, 0])
        a[0] += int(r[ii]) if r[ii].isdigit() else 0
        a[1] += int(r[si]) if r[si].isdigit() else 0
tot = sum(v[0] for v in data.values()) or 1
ts = sum(v[1] for v in data.values()) or 1
print("total warp-inst", tot, "samples", ts)
for (ln, src), v in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[0]/tot*100:5.1f}% inst {v[1]/ts*100:5.1f}% samp  L{ln}: {src}")
