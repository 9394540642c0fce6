#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-launch table committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "header comment" > profiles/<name>.csv"""
import csv
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main():
    rep, comment = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    out = csv.writer(sys.stdout)
    if comment:
        out.writerow(["# " + comment])
    out.writerow(["Kernel Name"] + cols)
    out.writerow([""] + [units[idx[c]] for c in cols])
    for d in data:
        out.writerow([d[idx["Kernel Name"]][:90]] + [d[idx[c]] for c in cols])


if __name__ == "__main__":
    main()
