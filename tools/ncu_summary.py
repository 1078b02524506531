#!/usr/bin/env python
"""Condense ncu exports into the summaries kept under profiles/.

  ncu_summary.py launches <launch-list.csv>   per-kernel totals / share of the captured window
  ncu_summary.py full <raw-page.csv>          one row per captured launch: time, DRAM bytes, throughputs, stalls
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"(gcl::)?(<unnamed>|unnamed>)::", "", name)
    return re.sub(r"\(.*$", "", name)[:70]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            ns = float(r[14])
        except ValueError:
            continue
        k = short(r[4])
        tot[k][0] += 1
        tot[k][1] += ns
    total = sum(v[1] for v in tot.values())
    print(f"# {len(rows)} launches, {total / 1e3:.1f} us of kernel time in the window (cold-cache, serialised: shares, not absolutes)")
    print("kernel,launches,total_us,avg_us,share")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{n},{ns / 1e3:.1f},{ns / n / 1e3:.2f},{ns / total:.4f}')


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
            ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
            ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
            ("smsp__inst_executed.sum", "warp_inst"),
            ("sm__inst_executed_pipe_tc.sum", "tc_inst"), ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tc_pct")]
    cols = [(m, n) for m, n in cols if m in hdr]
    units = rows[1]
    to_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}       # ncu picks the byte unit per report
    print("kernel," + ",".join(n for _, n in cols))
    for r in data:
        vals = []
        for m, n in cols:
            v = r[hdr.index(m)].replace(",", "")
            try:
                f = float(v)
                if n.endswith("_MB"):
                    f *= to_mb.get(units[hdr.index(m)], 1.0)
                vals.append(f"{f:.4g}")
            except ValueError:
                vals.append(v)
        print('"' + short(r[hdr.index("Kernel Name")]) + '",' + ",".join(vals))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
