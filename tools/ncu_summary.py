#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full) into a small text file for profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/prof_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = []
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append("kernel: " + name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("  %-72s %s %s" % (k, row[i], units[i]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    cols = src[h]
    isrc, iex, ismp = cols.index("Source"), cols.index("Instructions Executed"), cols.index("# Samples")
    ops, smp = Counter(), Counter()
    stall_cols = [i for i, c in enumerate(cols) if c.startswith("stall_") and "Not Issued" not in c]
    stalls = Counter()
    for r in src[h + 1:]:
        if len(r) <= iex or not r[iex].isdigit():
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[iex]); smp[op] += int(r[ismp] or 0)
        for i in stall_cols:
            if r[i].isdigit():
                stalls[cols[i]] += int(r[i])
    tot = sum(ops.values())
    lines.append("instruction mix of the first kernel (warp instructions executed, share, stall samples):")
    for op, n in ops.most_common(14):
        lines.append("  %-10s %14d %6.2f%%  %d" % (op, n, 100.0 * n / tot, smp[op]))
    st = sum(stalls.values()) or 1
    lines.append("warp stall samples (all): " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / st) for k, v in stalls.most_common(8)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
