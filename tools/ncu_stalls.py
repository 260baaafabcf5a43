"""Top stall sites of one kernel from `ncu --page source --csv` output (SASS view with -lineinfo).
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-skip N --launch-count 1 > s.csv
       python tools/ncu_stalls.py s.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        h = r
        continue
    if h and len(r) == len(h):
        data.append(r)
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("instructions", len(data), "samples", tot)
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
data.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for r in data[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = {s[6:]: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0)}
    print(f"{n:6d} {100.0*n/max(tot,1):5.1f}%  {r[col['Source']][:90]:90s} {why}")
