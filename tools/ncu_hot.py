#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: headline metrics (raw page) and the hottest SASS lines (source page).
  python tools/ncu_hot.py gpurun_out/x.ncu-rep [min_share_percent]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
keys += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:3]:
    print(r[hdr.index("Kernel Name")][:120])
    for k in keys:
        if k in hdr:
            v = r[hdr.index(k)]
            try:
                if float(v.replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            print(f"  {k:84s} {v} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
S, I, SRC = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Source")
first, seen = [], set()
for r in rows[hi + 1:]:
    if len(r) <= I or not r[0].startswith("0x"):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    first.append(r)
tot = sum(int(r[S]) for r in first) or 1
toti = sum(int(r[I]) for r in first) or 1
print(f"  SASS lines {len(first)}, samples {tot}, warp instructions {toti}")
for k, r in enumerate(first):
    s = int(r[S])
    if s >= tot * thr / 100:
        print(f"  {k:5d} {100 * s / tot:5.1f}% ex {int(r[I]):9d}  {r[SRC].strip()[:100]}")
