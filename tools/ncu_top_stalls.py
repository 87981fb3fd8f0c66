"""Top stall lines of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv > f.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data) or 1
data.sort(key=lambda t: -t[0])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total samples", tot)
for n, r in data[:top]:
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print(f"{100*n/tot:5.1f}% {r[col['Source']][:90]:90s} " + " ".join(f"{s[6:]}={c}" for c, s in st if c))
