#!/usr/bin/env python3
"""Index load timing on the GPU box: `fmsi_gpu_index_load` on reference-format files (file read +
device-side conversion + suffix table / dictionary) vs the reference's load_index (its `fmsi query`
on a one-record file). usage: load_time.py <prefix> [k]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402

prefix = sys.argv[1]
fg.device_count()
import torch  # noqa: E402
torch.cuda.init()
for label, kw in (("backward-search layout only (dict=0)", dict(dict=0)), ("with dictionary tier", dict(dict=1)), ("dict=0 again (page cache warm)", dict(dict=0))):
    t0 = time.perf_counter()
    gi = fg.Index.load(prefix, use_klcp=False, **kw)
    dt = time.perf_counter() - t0
    print(f"{label}: {dt:.2f} s  (N={gi.n}, prefix_t={gi.prefix_t}, hbm={gi.hbm_bytes / 1e9:.2f} GB)", flush=True)
    gi.close()
ref = os.path.join(ROOT, "oracle", "_ref", "fmsi")
if os.path.exists(ref):
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(">q\n" + "A" * 31 + "\n")
    t0 = time.perf_counter()
    subprocess.run([ref, "query", "-O", "-q", f.name, prefix], capture_output=True)
    print(f"reference load_index + 1 query: {time.perf_counter() - t0:.2f} s")
