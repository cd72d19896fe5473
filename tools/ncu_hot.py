"""Top stall sites of an .ncu-rep (SASS view): python tools/ncu_hot.py file.ncu-rep [n]"""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr): continue
    s = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[k]] or 0), k) for k in stalls), reverse=True)[:2]
    data.append((s, idx, r[ci["Source"]].strip()[:90], top))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, idx, src, top in sorted(data, reverse=True)[:n]:
    print(f"{s:7d} {100*s/tot:5.1f}%  #{idx:4d} {src:90s} {top}")
