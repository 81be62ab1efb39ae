#!/usr/bin/env python3
"""profiles/traffic.json entry from a summarised ncu capture.
usage: traffic_from_ncu.py <*_metrics.txt> <key> <kmers_per_launch> [note]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, key, kmers = sys.argv[1], sys.argv[2], int(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    rd = wr = None
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
            rd = float(f[-1].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
        if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
            wr = float(f[-1].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[f[1]]
    assert rd is not None and wr is not None, "dram__bytes_* not found"
    p = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(p))
    tj["kernels"][key] = dict(dram_bytes_per_kmer=round((rd + wr) / kmers, 2),
                              source=f"{os.path.relpath(path, ROOT)} (ncu --set full, {kmers} k-mers: {rd + wr:.0f} B{'; ' + note if note else ''})")
    json.dump(tj, open(p, "w"), indent=1)
    print(key, tj["kernels"][key])


if __name__ == "__main__":
    main()
