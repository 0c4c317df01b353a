"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    name = row["Kernel Name"]
    m = re.search(r"(\w+_kernel)", name)
    short = m.group(1) if m else name[:48]
    t = re.search(r"_kernel<([^>]*)>", name)
    if t and ("conv3x3" in short or "apply" in short or "gemm" in short):
        short += "<" + t.group(1).replace("(int)", "").replace("__nv_bfloat16", "bf16") + ">"
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'n':>5s} {'total_us':>10s} {'share':>6s} {'avg_us':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {v[0]:5d} {v[1]:10.1f} {v[1] / tot:6.3f} {v[1] / v[0]:8.1f}")
print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):5d} {tot:10.1f}")
