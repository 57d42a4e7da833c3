"""Summarises an .ncu-rep (one kernel) to the metrics the roofline needs.  Usage: python tools/ncu_summary.py rep [title]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_xu.sum"]
print(f"# {title}\n# ncu --set full --clock-control none ({rep.split('/')[-1]}); one block per captured kernel")
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    d = dict(zip(hdr, vals))
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h} [{u}] = {v}")
    try:   # derived: achieved DRAM bandwidth and its fraction of the measured HBM peak
        t_ns = float(d["gpu__time_duration.sum"].replace(",", ""))
        unit_t = units[hdr.index("gpu__time_duration.sum")]
        t_s = t_ns * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9}.get(unit_t, 1e-9)
        def nbytes(key):
            v = float(d[key].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(key)], 1)
        tot = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
        print(f"derived: HBM traffic {tot / 1e6:.2f} MB in {t_s * 1e6:.2f} us = {tot / t_s / 1e9:.0f} GB/s "
              f"({tot / t_s / 6.54e12 * 100:.1f} % of the measured 6540 GB/s)")
    except Exception as e:  # noqa: BLE001
        print("derived: n/a", e)
    print()
