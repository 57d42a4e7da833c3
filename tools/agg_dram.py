"""Aggregates an ncu `--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of
one agent step by kernel: DRAM bytes actually moved (per step and per launch) next to the time, as JSON on stdout.
bench.py reads profiles/<tag>_dram_traffic.json for `roofline.traffic`."""
import collections
import csv
import json
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}
agg = collections.defaultdict(lambda: {"launches": 0, "read": 0.0, "write": 0.0, "us": 0.0})
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("void ", "")
    k = "gn::gemm_tc_kernel" if "gemm_tc_kernel" in k else k
    v = float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        agg[k]["us"] += v
        agg[k]["launches"] += 1
    elif m == "dram__bytes_read.sum":
        agg[k]["read"] += v
    elif m == "dram__bytes_write.sum":
        agg[k]["write"] += v
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    tot = a["read"] + a["write"]
    out[k] = dict(launches=a["launches"], us_per_step=round(a["us"], 1), dram_read_bytes_per_step=int(a["read"]),
                  dram_write_bytes_per_step=int(a["write"]), dram_bytes_per_launch=int(tot / max(1, a["launches"])),
                  dram_gb_per_s=round(tot / max(a["us"], 1e-9) / 1e3, 1))
print(json.dumps(dict(source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                             "--clock-control none, python tools/one_step.py (one eager agent step, cold-cache, serialised)",
                      kernels=out), indent=1))
