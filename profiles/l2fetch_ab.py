#!/usr/bin/env python3
"""A/B of cudaLimitMaxL2FetchGranularity at human scale (one process per value): does asking L2 for 32-byte fills
change what a random 32-byte probe moves from DRAM, or its cost? The kernels already ask for `.L2::64B` fills per load
(device_index.cuh); ncu r02c shows 64 bytes moved per 32-byte probe (`wasted` 1.79 on fold_query_kernel).

    python profiles/l2fetch_ab.py --gran 0|32|64|128 --dict 2|0 --label NAME      (0 = leave the driver default)
Run once plain (timed with CUDA events) and once under
    ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'fold_query|query_kmers_kernel' -c 8
for the bytes (timings under ncu are not bench values).
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
import bench  # noqa: E402

CUDA_LIMIT_MAX_L2_FETCH_GRANULARITY = 0x05


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label", required=True)
    ap.add_argument("--gran", type=int, default=0)
    ap.add_argument("--dict", type=int, default=2)
    ap.add_argument("--genome", type=int, default=3_100_000_000)
    ap.add_argument("--batch", type=int, default=1 << 26)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    a = argparse.Namespace(batch=0, no_parity=True)
    cx = bench.Ctx(a)
    torch.zeros(1, device=cx.dev)
    rt = ctypes.CDLL("libcudart.so.12")
    got = ctypes.c_size_t(0)
    if args.gran:
        rc = rt.cudaDeviceSetLimit(CUDA_LIMIT_MAX_L2_FETCH_GRANULARITY, ctypes.c_size_t(args.gran))
        assert rc == 0, rc
    rt.cudaDeviceGetLimit(ctypes.byref(got), CUDA_LIMIT_MAX_L2_FETCH_GRANULARITY)
    k = 31
    codes, ascii_ = bench.device_genome(args.genome, 4, k, cx.dev)
    t0 = time.time()
    gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=args.dict)
    build_s = time.time() - t0
    del ascii_
    torch.cuda.empty_cache()
    wl = dict(name="human", k=k, codes=codes, genome=None, reads=0)
    out = dict(label=args.label, l2_fetch_granularity=got.value, genome=args.genome, build_s=round(build_s, 2), tier=gi.dict_kind,
               prefix_t=gi.prefix_t, multistep=gi.multistep, hbm_gb=round(gi.hbm_bytes / 1e9, 2))
    r = bench.bench_kmers(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O", args.batch, args.steps, 3, 1000, True)
    out["query_O"] = dict(gkmers_s=round(r["value"] / 1e9, 3), ms=round(r["ms_per_step"], 4), probes_per_kmer=r["roofline"]["algorithmic"]["probes_per_kmer"],
                          gprobes_s=r["roofline"]["request_rate"]["gprobes_s"], frac=r["roofline"]["frac"])
    rt.cudaDeviceGetLimit(ctypes.byref(got), CUDA_LIMIT_MAX_L2_FETCH_GRANULARITY)
    out["l2_fetch_granularity_after"] = got.value
    print(json.dumps(out))


if __name__ == "__main__":
    main()
