"""Summarise an ncu --set full report into profiles/ (per-launch duration, DRAM traffic, tensor %, ...)
and profiles/traffic.json (dram bytes per launch per stage, consumed by bench.py's roofline.traffic)."""
import csv, json, subprocess, sys, collections
rep, out_csv, traffic_json = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [h.index(c) for c in cols]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
# stage mapping by launch order within one forward of the tcgen05 engine
order = ["condition", "in_linear"] + ["qkv", "attention", "out_proj_ln", "ff1", "ff2_ln"] * 4 + ["rnn_ih", "rnn", "head"]
names = [r[h.index("Kernel Name")] for r in rows[2:]]
start = next((i for i, n in enumerate(names) if "condition" in n), None)
traffic = collections.defaultdict(list)
if start is not None:
    ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    for k, st in enumerate(order):
        if start + k < len(rows) - 2:
            r = rows[2 + start + k]
            traffic[st].append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(traffic_json, "w"), indent=1)
print(open(traffic_json).read())
