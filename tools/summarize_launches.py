"""Fold an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals.
    python tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/<name>.csv"""
import csv, io, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    text = f.read()
start = text.find('"ID"')
rd = csv.DictReader(io.StringIO(text[start:]))
agg = OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"<.*", "", name)                     # drop template arguments
    name = re.sub(r"\(.*", "", name).strip()
    if name.startswith("void at::native::"):
        m = re.search(r"(\w+Functor|\w+_kernel_cuda|\w+_kernel)", r["Kernel Name"][len("void at::native::"):])
        name = "at::" + r["Kernel Name"][len("void at::native::"):].split("<")[0] + (":" + m.group(1) if m else "")
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("kernel,launches,total_us,share")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{us:.1f},{us / total:.4f}")
print(f"TOTAL,{sum(v[0] for v in agg.values())},{total:.1f},1.0000")
