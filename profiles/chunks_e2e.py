#!/usr/bin/env python3
"""End-to-end rate of fmsi_gpu_query_chunks with HOST buffers on 150-bp reads (BASELINE configs[1] shape): one call
over all reads (text pieces, queries and result copies overlap inside the call) against the same reads sent as calls
below the pipelining threshold (each: upload, query, download, one after the other); pinned and pageable buffers.

    python profiles/chunks_e2e.py [--genome 5000000] [--reads 1000000] [--dict 2]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
from bench import device_genome  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome", type=int, default=5_000_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--dict", type=int, default=2)
args = ap.parse_args()
k, L = 31, 150
dev = torch.device("cuda", 0)
codes, ascii_ = device_genome(args.genome, 4, k, dev)
gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=args.dict)
gen = torch.Generator(device=dev)
gen.manual_seed(9)
R = args.reads
pos = torch.randint(0, args.genome - L + 1, (R,), device=dev, generator=gen)
rd = codes[(pos[:, None] + torch.arange(L, device=dev)[None, :])]
sub = torch.rand(R, L, device=dev, generator=gen) < 0.01
shift = torch.randint(1, 4, (R, L), device=dev, generator=gen, dtype=torch.uint8)
rd = torch.where(sub, (rd + shift) & 3, rd)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
nk = L - k + 1
spans = [(0, 64), (64, nk - 64)]
base = np.arange(R, dtype=np.uint64) * L
off = np.stack([base + p for p, m in spans], 1).reshape(-1)
ln = np.tile(np.array([m + k - 1 for p, m in spans], dtype=np.uint32), R)
roff = np.stack([np.arange(R, dtype=np.uint64) * nk + p for p, m in spans], 1).reshape(-1)
n_res = R * nk
lib = fg.lib()
out = {"reads": R, "kmers": n_res, "dict": args.dict, "rows": []}
for pinned in (True, False):
    bases_t = lut[rd.long()].reshape(-1).cpu()
    res_t = torch.empty(n_res, dtype=torch.uint8)
    if pinned:
        bases_t, res_t = bases_t.pin_memory(), res_t.pin_memory()

    def prepare(c0, c1):  # per-call arguments, built outside the timed region
        b0 = int(off[c0])
        b1 = int(off[c1 - 1] + ln[c1 - 1])
        r0 = int(roff[c0])
        r1 = int(roff[c1 - 1]) + int(ln[c1 - 1]) - k + 1
        arrs = [np.ascontiguousarray(off[c0:c1] - np.uint64(b0)).view(np.int64), np.ascontiguousarray(ln[c0:c1]).view(np.int32),
                np.ascontiguousarray(roff[c0:c1] - np.uint64(r0)).view(np.int64)]
        ts = [torch.from_numpy(a) for a in arrs]
        if pinned:
            ts = [t.pin_memory() for t in ts]
        return (b0, b1, r0, r1, ts[0], ts[1], ts[2], c1 - c0)

    def call(a, streaming):
        b0, b1, r0, r1, o, l, r, n = a
        rc = lib.fmsi_gpu_query_chunks(gi._h, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, streaming, bases_t.data_ptr() + b0, b1 - b0,
                                       o.data_ptr(), l.data_ptr(), r.data_ptr(), n, r1 - r0, k, res_t.data_ptr() + r0, fg.MEM_HOST, None)
        assert rc == 0, lib.fmsi_gpu_last_error()

    for label, step in (("one call (pipelined inside)", len(off)), ("calls of 40 k chunks (serial stages)", 40_000)):
        calls = [prepare(c, min(len(off), c + step)) for c in range(0, len(off), step)]
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            for a in calls:
                call(a, 1)
            best = min(best, time.perf_counter() - t0)
        row = {"buffers": "pinned" if pinned else "pageable", "mode": label, "ms": round(best * 1e3, 2), "gkmers_s": round(n_res / best / 1e9, 2),
               "present_frac": round(float(res_t.float().mean()), 4)}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
print(json.dumps(out))
