#!/usr/bin/env python3
"""Whole-process start-up cost at human scale on the GPU box: a 3.1 Gbp index is built on the GPU and saved in the
reference's format; then `fmsi query -O` on ONE record is run in fresh processes with $FMSI_GPU_TIMING (stage
times of fmsi_gpu_index_load), next to the reference binary on the same files.
usage: load_time_cli.py [--genome 3100000000]"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
from bench import device_genome  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=3_100_000_000)
args = ap.parse_args()
k = 31
dev = torch.device("cuda", 0)
d = tempfile.mkdtemp(prefix="fmsi_load_")
prefix = os.path.join(d, "ms.fa")
codes, ascii_ = device_genome(args.genome, 4, k, dev)
t0 = time.perf_counter()
gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=0)
t1 = time.perf_counter()
gi.save(prefix)
print(f"built in {t1 - t0:.2f} s, saved in {time.perf_counter() - t1:.2f} s: "
      + ", ".join(f"{e} {os.path.getsize(prefix + '.fmsi.' + e) >> 20} MiB" for e in ("ac_gt", "ac", "gt", "klcp", "mask")), flush=True)
gi.close()
del codes, ascii_
torch.cuda.empty_cache()
q = os.path.join(d, "q.fa")
open(q, "w").write(">q\n" + "ACGT" * 8 + "\n")
cli = os.path.join(ROOT, "fmsi_b200", "bin", "fmsi")
env = dict(os.environ, FMSI_GPU_TIMING="1")
for flags in (["-O"], ["-O", "-S"], ["-O"]):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "query", *flags, "-q", q, prefix], capture_output=True, env=env)
    dt = time.perf_counter() - t0
    print(f"fmsi query {' '.join(flags)} (1 record), whole process: {dt:.2f} s -> {r.stdout.decode().strip()}")
    print("   " + "\n   ".join(l for l in r.stderr.decode().splitlines() if "timing" in l), flush=True)
ref = os.path.join(ROOT, "oracle", "_ref", "fmsi")
if os.path.exists(ref):
    for flags in (["-O"], ["-O", "-S"]):
        t0 = time.perf_counter()
        r = subprocess.run([ref, "query", *flags, "-q", q, prefix], capture_output=True)
        print(f"reference fmsi query {' '.join(flags)} (1 record), whole process: {time.perf_counter() - t0:.2f} s -> {r.stdout.decode().strip()}", flush=True)
