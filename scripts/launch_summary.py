"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python scripts/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
seen = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    seen += 1
    if seen <= skip:
        continue
    name = row["Kernel Name"]
    m = re.match(r".*?(gemm_bf16_tcgen05_kernel<[^>]*>|[A-Za-z_0-9]+_kernel)", name)
    key = m.group(1) if m else name[:70]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.3f} ms {v[0]:5d}x {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] * 1e3:9.1f} us  {k}")
print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
