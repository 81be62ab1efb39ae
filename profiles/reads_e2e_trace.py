#!/usr/bin/env python3
"""Host-side timeline of the e2e reads path (fmsi_gpu_query_chunks_packed with pinned host buffers): run with
FMSI_GPU_TRACE=1. 1 M reads of 150 bp against a GPU-built index; prints device-resident vs host-buffer rates."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=400_000_000)
    ap.add_argument("--dict", type=int, default=-1)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--locality", type=int, default=0, help="1 = build the minimizer-bucketed dictionary at load, -1 = never")
    args = ap.parse_args()
    cx = bench.Ctx(argparse.Namespace(batch=0, no_parity=True))
    codes, ascii_ = bench.device_genome(args.genome, 4, 31, cx.dev)
    gi = fg.Index.build(ascii_.data_ptr(), 31, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=args.dict, locality=args.locality)
    del ascii_
    torch.cuda.empty_cache()
    wl = dict(name="trace", k=31, codes=codes, genome=None, reads=args.reads)
    r, _ = bench.bench_reads(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, True, "query -O -S", args.reads, 5, 3, 3000)
    gi.refresh_info()
    print(json.dumps(dict(tier=gi.dict_kind, locality=gi.locality, hbm_gb=round(gi.hbm_bytes / 1e9, 1), probes_per_kmer=r["roofline"]["algorithmic"]["probes_per_kmer"],
                          device_ms=round(r["ms_per_step"], 3), device_gkmers_s=round(r["value"] / 1e9, 2), e2e_gkmers_s=round(r["e2e"]["value"] / 1e9, 2), e2e=r["e2e"])))


if __name__ == "__main__":
    main()
