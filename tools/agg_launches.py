"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and gemm_tc by grid)."""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
rows = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    k = row["Kernel Name"].split("(")[0].replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += v
    rows.append((k, row["Grid Size"], v))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(rows)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:56]:56s} n={v[0]:5d} us={v[1]:10.1f} share={v[1] / tot:.3f} avg={v[1] / v[0]:.1f}")
if "--grids" in sys.argv:
    g = collections.defaultdict(lambda: [0, 0.0])
    for k, grid, v in rows:
        if "gemm_tc" in k:
            g[grid][0] += 1
            g[grid][1] += v
    print("--- gemm_tc_kernel by grid")
    for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:30]:
        print(k, v[0], round(v[1], 1), round(v[1] / v[0], 1))
