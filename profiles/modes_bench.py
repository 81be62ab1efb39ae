#!/usr/bin/env python3
"""Device-resident throughput of every query mode (not bench.py's headline line): single k-mers in
-O / or / lookup mode through the strand-folded dictionary, the SA-ordered dictionary and plain backward search, and 150 bp reads
with 1 % substitutions through the kLCP streaming kernel vs the single-k-mer path, LAZY and BOTH
strands. BASELINE configs[1] / configs[2] shapes on a GPU-built index.

    python profiles/modes_bench.py [--genome 5000000] [--kmers 33554432] [--reads 1000000]
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
from bench import device_genome, device_queries  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--kmers", type=int, default=1 << 25)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--copies", type=int, default=1, help="pangenome-like: this many mutated copies of the genome, concatenated")
    ap.add_argument("--snp", type=float, default=0.01, help="per-base substitution rate of every copy")
    ap.add_argument("--variants", type=int, default=0,
                    help="BASELINE configs[4] at its named scale: a masked superstring of the base genome + this many SNP variants, "
                         "each contributing its k new k-mers as a 2k-1 window (ON) joined to the next by k-1 OFF positions")
    ap.add_argument("--tiers", default="fold,dict,backward")
    args = ap.parse_args()
    k = args.k
    dev = torch.device("cuda", 0)
    codes, ascii_ = device_genome(args.genome, 4, k, dev)
    if args.copies > 1:  # BASELINE configs[4] shape, scaled: every k-mer occurs ~copies times (large SA intervals)
        gen0 = torch.Generator(device=dev)
        gen0.manual_seed(77)
        parts = []
        for c in range(args.copies):
            sub = torch.rand(args.genome, device=dev, generator=gen0) < args.snp
            shift = torch.randint(1, 4, (args.genome,), device=dev, generator=gen0, dtype=torch.uint8)
            parts.append(torch.where(sub, (codes + shift) & 3, codes))
        codes = torch.cat(parts)
        lut0 = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        ascii_ = lut0[codes.long()]
        ascii_[codes.numel() - (k - 1):] += 32
        args.genome = codes.numel()
    if args.variants > 0:
        # Pangenome as a masked superstring (what kmercamel emits for many near-identical genomes): the base genome
        # once, then for every SNP the window of 2k-1 bases around it (its k new k-mers, mask ON), windows back to
        # back, so the k-1 k-mers that straddle two windows are OFF occurrences: a mask-heavy superstring whose
        # distinct represented k-mers number ~ genome + variants * k.
        gen0 = torch.Generator(device=dev)
        gen0.manual_seed(78)
        V, W = args.variants, 2 * k - 1
        n_total = args.genome + V * W
        sup = torch.empty(n_total, dtype=torch.uint8, device=dev)
        upper = torch.ones(n_total, dtype=torch.bool, device=dev)
        sup[:args.genome] = codes
        upper[args.genome - (k - 1):args.genome] = False  # k-mers running from the genome into the first window
        step = 1 << 22
        ar = torch.arange(W, device=dev)
        for a in range(0, V, step):
            b = min(V, a + step)
            pos = torch.randint(k - 1, args.genome - k, (b - a,), device=dev, generator=gen0)
            win = codes[(pos[:, None] + (ar[None, :] - (k - 1)))]
            shift = torch.randint(1, 4, (b - a,), device=dev, generator=gen0, dtype=torch.uint8)
            win[:, k - 1] = (win[:, k - 1] + shift) & 3
            sup[args.genome + a * W:args.genome + b * W] = win.reshape(-1)
            del win, pos, shift
        wpos = (torch.arange(V * W, device=dev) % W)
        upper[args.genome:] = wpos < k  # the k k-mers that start inside a window's first k positions contain the SNP
        del wpos
        upper[n_total - (k - 1):] = False
        lut0 = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        ascii_ = lut0[sup.long()] + (~upper).to(torch.uint8) * 32
        codes = sup
        on_positions = upper
        args.genome = n_total
        del sup
        torch.cuda.empty_cache()
    out = {"genome": args.genome, "copies": args.copies, "variants": args.variants, "k": k, "rows": []}
    idx = {}
    for name, dct in (("fold", 2), ("dict", 1), ("backward", 0)):
        if name not in args.tiers.split(","):
            continue
        t0 = time.time()
        idx[name] = fg.Index.build(ascii_.data_ptr(), k, with_klcp=True, device=0, n=args.genome, mem=fg.MEM_DEVICE, dict=dct)
        print(f"index[{name}]: built in {time.time() - t0:.2f}s, t={idx[name].prefix_t}, hbm={idx[name].hbm_bytes / 1e9:.2f} GB", flush=True)
    stream = torch.cuda.current_stream(dev).cuda_stream
    n = args.kmers
    q = device_queries(codes, k, n, 1, dev)
    res8 = torch.empty(n, dtype=torch.uint8, device=dev)
    res64 = torch.empty(2 * n, dtype=torch.int64, device=dev)
    for label, mode, outp in (("query -O", fg.MODE_ALL, fg.OUT_PRESENCE), ("query (or)", fg.MODE_OR, fg.OUT_PRESENCE), ("lookup", fg.MODE_OR, fg.OUT_ORDERS)):
        for sname, strands in (("lazy", fg.STRANDS_LAZY), ("both", fg.STRANDS_BOTH)):
            row = {"case": f"single 31-mers, {label}, {sname}"}
            for name in idx:
                dst = res8 if outp == fg.OUT_PRESENCE else res64
                dt = timed(lambda: idx[name].query_kmers_ptr(q.data_ptr(), n, dst.data_ptr(), k, mode, outp, strands, fg.MEM_DEVICE, stream))
                row[f"{name}_gkmers_s"] = round(n / dt / 1e9, 2)
            out["rows"].append(row)
            print(json.dumps(row), flush=True)
    # reads: 150 bp, 1 % substitutions, random strand; chunks of <= 64 k-mers overlapping by k-1
    R, L = args.reads, 150
    gen = torch.Generator(device=dev)
    gen.manual_seed(9)
    pos = torch.randint(0, args.genome - L + 1, (R,), device=dev, generator=gen)
    rd = codes[(pos[:, None] + torch.arange(L, device=dev)[None, :])]
    flip = torch.rand(R, device=dev, generator=gen) < 0.5
    rd = torch.where(flip[:, None], 3 - rd.flip(1), rd)
    sub = torch.rand(R, L, device=dev, generator=gen) < 0.01
    shift = torch.randint(1, 4, (R, L), device=dev, generator=gen, dtype=torch.uint8)
    rd = torch.where(sub, (rd + shift) & 3, rd).contiguous()
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = lut[rd.long()].contiguous().view(-1)
    nk = L - k + 1
    first = min(64, nk)
    offs, lens, roff = [], [], []
    base = torch.arange(R, device=dev, dtype=torch.int64) * L
    rbase = torch.arange(R, device=dev, dtype=torch.int64) * nk
    p, chunks = 0, []
    while p < nk:
        m = min(64, nk - p)
        chunks.append((p, m))
        p += m
    off_t = torch.stack([base + p for p, m in chunks], 1).contiguous().view(-1)
    len_t = torch.tensor([m + k - 1 for p, m in chunks], device=dev, dtype=torch.int32).repeat(R).contiguous()
    roff_t = torch.stack([rbase + p for p, m in chunks], 1).contiguous().view(-1)
    n_res = R * nk
    r8 = torch.empty(n_res, dtype=torch.uint8, device=dev)
    r64 = torch.empty(2 * n_res, dtype=torch.int64, device=dev)
    L_ = fg.lib()
    for label, mode, outp in (("query -O", fg.MODE_ALL, fg.OUT_PRESENCE), ("lookup", fg.MODE_OR, fg.OUT_ORDERS)):
        for sname, strands in (("lazy", fg.STRANDS_LAZY), ("both", fg.STRANDS_BOTH)):
            row = {"case": f"150 bp reads (1% subs), {label}, {sname}"}
            for name in idx:
                for sm, streaming in (("S", 1), ("single", 0)):
                    dst = r8 if outp == fg.OUT_PRESENCE else r64

                    def call():
                        rc = L_.fmsi_gpu_query_chunks(idx[name]._h, mode, outp, strands, streaming, bases.data_ptr(), bases.numel(), off_t.data_ptr(),
                                                      len_t.data_ptr(), roff_t.data_ptr(), off_t.numel(), n_res, k, dst.data_ptr(), fg.MEM_DEVICE, stream)
                        assert rc == 0, L_.fmsi_gpu_last_error()
                    dt = timed(call)
                    row[f"{name}_{sm}_gkmers_s"] = round(n_res / dt / 1e9, 2)
            out["rows"].append(row)
            print(json.dumps(row), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
