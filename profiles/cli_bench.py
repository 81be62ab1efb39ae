#!/usr/bin/env python3
"""End-to-end CLI timing on the GPU box (not part of bench.py's contract): BASELINE configs[1] and
configs[2] shapes through `fmsi_b200/bin/fmsi` vs the reference binary on a subsample.

    python profiles/cli_bench.py [--genome 5000000] [--reads 1000000] [--ref-reads 20000]

Builds the index with the GPU builder (byte-identical to `fmsi index`), writes FASTA query files,
times `fmsi query -O`, `query -O -S`, `lookup`, `lookup -S` (default exact mode and
FMSI_GPU_STRANDS=lazy) and checks the first --ref-reads records byte for byte against oracle/_ref/fmsi.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
from fmsi_b200 import synth  # noqa: E402

CLI = os.path.join(ROOT, "fmsi_b200", "bin", "fmsi")
REF = os.path.join(ROOT, "oracle", "_ref", "fmsi")


def timed(cmd, env=None, stdout=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=stdout or subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{cmd}: {r.stderr.decode()[-500:]}")
    timed.last_stderr = r.stderr.decode(errors="replace")
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--kmers", type=int, default=10_000_000)
    ap.add_argument("--ref-reads", type=int, default=20_000)
    ap.add_argument("--k", type=int, default=31)
    args = ap.parse_args()
    k = args.k
    d = tempfile.mkdtemp(prefix="fmsi_cli_")
    g = synth.random_codes(args.genome, 11)
    ms = synth.genome_superstring(g, k)
    prefix = os.path.join(d, "ms.fa")
    synth.write_fasta_single(prefix, "ms", ms)
    t0 = time.time()
    gi = fg.Index.build(ms, k, with_klcp=True)
    gi.save(prefix)
    gi.close()
    print(f"index: {args.genome} bp built+saved on GPU in {time.time() - t0:.1f}s", flush=True)

    reads = synth.read_queries(g, 150, args.reads, 2)
    rfa = os.path.join(d, "reads.fa")
    with open(rfa, "wb") as f:
        f.write(synth.rows_to_fasta(reads, "r"))
    kq = synth.kmer_queries(g, k, args.kmers, 3)
    kfa = os.path.join(d, "kmers.fa")
    with open(kfa, "wb") as f:
        f.write(synth.rows_to_fasta(kq, "q"))
    sub_r, sub_k = os.path.join(d, "reads_sub.fa"), os.path.join(d, "kmers_sub.fa")
    with open(sub_r, "wb") as f:
        f.write(synth.rows_to_fasta(reads[:args.ref_reads], "r"))
    with open(sub_k, "wb") as f:
        f.write(synth.rows_to_fasta(kq[:args.ref_reads * 20], "q"))
    n_read_kmers = args.reads * (150 - k + 1)
    out = {"genome": args.genome, "k": k, "reads": args.reads, "read_kmers": n_read_kmers, "single_kmers": args.kmers, "runs": []}
    cases = [("query -O (single 31-mers)", ["query", "-O"], kfa, sub_k, args.kmers, args.ref_reads * 20),
             ("lookup (single 31-mers)", ["lookup"], kfa, sub_k, args.kmers, args.ref_reads * 20),
             ("query -O -S (150 bp reads)", ["query", "-O", "-S"], rfa, sub_r, n_read_kmers, args.ref_reads * (150 - k + 1)),
             ("query -O (150 bp reads)", ["query", "-O"], rfa, sub_r, n_read_kmers, args.ref_reads * (150 - k + 1)),
             ("lookup -S (150 bp reads)", ["lookup", "-S"], rfa, sub_r, n_read_kmers, args.ref_reads * (150 - k + 1))]
    for name, flags, qf, subf, units, sub_units in cases:
        row = {"case": name}
        for label, env in (("exact", dict(os.environ, FMSI_GPU_TIMING="1")), ("lazy", dict(os.environ, FMSI_GPU_STRANDS="lazy", FMSI_GPU_TIMING="1"))):
            dt = timed([CLI, *flags, "-q", qf, prefix], env=env)
            row[f"{label}_s"] = round(dt, 3)
            row[f"{label}_mkmers_s"] = round(units / dt / 1e6, 2)
            # the CLI's own phase clock: index load (incl. CUDA context creation), then the query pipeline
            ph = {ln.split("] ")[1].split(":")[0]: float(ln.rsplit(":", 1)[1].split()[0]) for ln in timed.last_stderr.splitlines() if ln.startswith("[fmsi timing]")}
            row[f"{label}_phases_s"] = ph
            if "pipeline drained" in ph and ph["pipeline drained"] > 0:
                row[f"{label}_mkmers_s_after_load"] = round(units / ph["pipeline drained"] / 1e6, 2)
        # parity + reference rate on the subsample
        a = os.path.join(d, "a.txt")
        b = os.path.join(d, "b.txt")
        with open(a, "wb") as fa_:
            timed([CLI, *flags, "-q", subf, prefix], stdout=fa_)
        with open(b, "wb") as fb_:
            dt_ref = timed([REF, *flags, "-q", subf, prefix], stdout=fb_)
        row["ref_1core_mkmers_s"] = round(sub_units / dt_ref / 1e6, 3)
        row["byte_identical_on_subsample"] = open(a, "rb").read() == open(b, "rb").read()
        out["runs"].append(row)
        print(json.dumps(row), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
