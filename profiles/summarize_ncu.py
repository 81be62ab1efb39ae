#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU) into the text files kept under profiles/.
usage: summarize_ncu.py <report.ncu-rep> <out_prefix> [iterations_counter_line_regex]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "dram__sectors_write.sum",
    "dram__bytes_read.sum.per_second", "sm__cycles_elapsed.max",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    with open(out + "_metrics.txt", "w") as f:
        for row in raw[2:]:
            name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"== kernel: {name}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct")):
                    f.write(f"{h:95s} {units[i]:16s} {row[i]}\n")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]))))
    cur, agg = None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or not r[0].isdigit():
            continue
        try:
            agg.append((int(r[7]), int(r[4]) if r[4] != "-" else 0, cur, int(r[0]), r[1][:100]))
        except ValueError:
            pass
    tot = sum(a[0] for a in agg) or 1
    ts = sum(a[1] for a in agg) or 1
    with open(out + "_source_top.txt", "w") as f:
        f.write(f"# total warp instructions {tot}, stall samples {ts}\n# inst%  stall%  file:line  source\n")
        for ie, s, fn, ln, text in sorted(agg, reverse=True)[:60]:
            f.write(f"{100 * ie / tot:6.2f} {100 * s / ts:6.2f}  {fn}:{ln}  {text}\n")


if __name__ == "__main__":
    main()
