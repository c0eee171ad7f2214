#!/usr/bin/env python
"""Compact markdown summary of an ncu report (one row per captured launch).   python tools/ncu_summary.py rep.ncu-rep"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "xu pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefront %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp insts")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")


def find(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name) or h.endswith(name):
            return i
    return None


idx = [(find(n), lab) for n, lab in COLS]
print("| # | kernel | " + " | ".join(f"{lab} [{units[i]}]" if i is not None and units[i] else lab for i, lab in idx) + " |")
print("|---|---|" + "---|" * len(idx))
for n, r in enumerate(data):
    vals = []
    for i, lab in idx:
        v = r[i] if i is not None else "-"
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.4g}"
        except Exception:
            pass
        vals.append(v)
    print(f"| {n} | {r[ki].split('(')[0][-40:]} | " + " | ".join(vals) + " |")
