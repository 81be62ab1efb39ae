#!/usr/bin/env python3
"""A/B of the backward-search tier at human scale (one process per variant; the variant is the environment):
single 31-mers `query -O` and 150-bp reads `query -O -S`, device-resident, with the kernels' own probe counts.

    [FMSI_GPU_LIB=...] [FMSI_GPU_MULTISTEP=0|2|3] [FMSI_GPU_PREFIX_T=t] [FMSI_GPU_L2_PERSIST=table:96] \
        python profiles/backward_ab.py --label NAME [--genome 3100000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label", required=True)
    ap.add_argument("--genome", type=int, default=3_100_000_000)
    ap.add_argument("--batch", type=int, default=1 << 26)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--dict", type=int, default=0)
    args = ap.parse_args()
    a = argparse.Namespace(batch=0, no_parity=True)
    cx = bench.Ctx(a)
    k = 31
    codes, ascii_ = bench.device_genome(args.genome, 4, k, cx.dev)
    t0 = time.time()
    gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=args.dict)
    build_s = time.time() - t0
    del ascii_
    torch.cuda.empty_cache()
    wl = dict(name="human", k=k, codes=codes, genome=None, reads=args.reads)
    out = dict(label=args.label, env={e: os.environ[e] for e in os.environ if e.startswith("FMSI_GPU_")}, genome=args.genome, build_s=round(build_s, 2),
               tier=gi.dict_kind, prefix_t=gi.prefix_t, multistep=gi.multistep, hbm_gb=round(gi.hbm_bytes / 1e9, 2))
    r = bench.bench_kmers(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O", args.batch, args.steps, 3, 1000, True)
    out["query_O"] = dict(gkmers_s=round(r["value"] / 1e9, 3), ms=round(r["ms_per_step"], 4), probes_per_kmer=r["roofline"]["algorithmic"]["probes_per_kmer"],
                          gprobes_s=r["roofline"]["request_rate"]["gprobes_s"], frac=r["roofline"]["frac"], e2e_gkmers_s=round(r["e2e"]["value"] / 1e9, 3))
    r = bench.bench_kmers(cx, gi, wl, fg.MODE_OR, fg.OUT_ORDERS, "lookup", args.batch, args.steps, 3, 2000, False)
    out["lookup"] = dict(gkmers_s=round(r["value"] / 1e9, 3), probes_per_kmer=r["roofline"]["algorithmic"]["probes_per_kmer"])
    r, _ = bench.bench_reads(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, True, "query -O -S", args.reads, args.steps, 3, 3000)
    out["reads_S"] = dict(gkmers_s=round(r["value"] / 1e9, 3), ms=round(r["ms_per_step"], 4), probes_per_kmer=r["roofline"]["algorithmic"]["probes_per_kmer"],
                          gprobes_s=r["roofline"]["request_rate"]["gprobes_s"], e2e_gkmers_s=round(r["e2e"]["value"] / 1e9, 3))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
