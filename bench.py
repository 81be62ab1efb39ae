#!/usr/bin/env python3
"""bench.py — 31-mer queries/s of the FMS-index query path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Headline (`metric` / `value` / `e2e` / `roofline` / `cpu_baseline`): BASELINE.json configs[3] — the human-scale
index (3.1 Gbp, k = 31), one "step" = one pass of the hot path (`fmsi query -O`: fmsi_gpu_query_kmers, mode ALL,
strands LAZY) over one batch of 2^26 synthetic packed 31-mers, 50 % present. `value` is measured with the batch
resident in HBM (CUDA events on the launching stream), `e2e` through the C-ABI with pinned HOST buffers
(host<->device copies inside the timed region).

The same JSON line carries
  `tiers`  the same workload on the backward-search tier (dict = 0: query_kmers_kernel / stream_kernel — the tier for
           wide indexes and small memory), with its own roofline and parity sample;
  `modes`  the other BASELINE.json configs — E. coli-sized `query -O` (configs[0]), `-S` reads (configs[1]), `lookup`
           (configs[2]), human-scale reads and lookup, pangenome-like k = 23 / 31 streaming reads (configs[4]) — each
           with value, e2e, roofline, cpu_baseline (the reference binary on the same sample) and an oracle parity
           sample;
  `cli`    whole-process `fmsi query -O` of this repo's drop-in binary against the reference binary on the same
           FASTA, outputs compared byte for byte.
Every roofline describes the TIMED kernel: algorithmic bytes = what that kernel must move per k-mer (query in,
result out, 32 B per dependent probe it issues — counted live by the kernels' probe accounting), against the
measured HBM peak; `vs_reference_algorithm` keeps SURVEY 8(d)'s figure for the reference's backward search.

With N > 1 (torchrun) every rank owns one GPU, a full replica of the index and its own batch (weak scaling, no
data-path collective); times are the max over ranks; a sample of every rank's answers is checked against the
oracle on rank 0. `--impl reference` times the unmodified reference CPU binary (oracle/_ref/fmsi, one process per
host core over FASTA shards of a bounded sample) on the same index files and query distribution.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from fmsi_b200 import synth  # noqa: E402

DATA = os.path.join(ROOT, "data")
REF_FMSI = os.path.join(ROOT, "oracle", "_ref", "fmsi")
REF_KMERCAMEL = os.path.join(ROOT, "oracle", "_ref", "kmercamel")
OUR_FMSI = os.path.join(ROOT, "fmsi_b200", "bin", "fmsi")
METRIC = "31-mer queries/sec"
UNIT = "kmers/s"
READ_LEN = 150
# measured ceiling of dependent random 32-byte reads that miss L2 on this part (tools/randbw2.cu,
# profiles/r01b_randbw2_b200.jsonl): the DRAM activate rate, whatever each request returns
RANDOM_REQUEST_CEILING = 36.9e9

WORKLOADS = {
    # BASELINE.json configs[0..2]: E. coli-sized synthetic, k=31, masked superstring via bundled kmercamel
    "ecoli": dict(genome_len=5_000_000, k=31, seed=1, superstring="kmercamel", batch=1 << 26, reads=1_000_000,
                  desc="5 Mbp random genome, kmercamel -c + optimize -a ones, fmsi index -k 31"),
    # BASELINE.json configs[3]: human-scale synthetic: 3.1 Gbp i.i.d. sequence (itself a valid max-ones masked
    # superstring: upper case except the last k-1 letters), k=31, index replicated per GPU. The reference's `fmsi index`
    # needs hours and ~53 GB for this, so the index is built on the GPU (fmsi_gpu_index_build: byte-identical files,
    # tests/test_gpu_build.py up to 100 Mbp) and saved for the CPU arm.
    "human": dict(genome_len=3_100_000_000, k=31, seed=4, superstring="genome", batch=1 << 26, reads=1_000_000, device_built=True,
                  desc="3.1 Gbp i.i.d. sequence as max-ones masked superstring, k=31, index built on GPU"),
    "human_small": dict(genome_len=400_000_000, k=31, seed=4, superstring="genome", batch=1 << 26, reads=1_000_000, device_built=True,
                        desc="400 Mbp i.i.d. sequence (reduced human-scale shape, index >> L2), k=31"),
    # BASELINE.json configs[4]: pangenome-like, ~1.2 G distinct k-mers: the base genome once, then one window of 2k-1
    # bases per SNP variant (its k new k-mers ON, the k-1 k-mers that straddle two windows OFF) — the mask-heavy
    # superstring kmercamel emits for many near-identical genomes, generated directly (its hash tables do not fit)
    "pangenome_k31": dict(genome_len=5_000_000, k=31, seed=4, superstring="variants", variants=39_000_000, batch=1 << 26, reads=1_000_000,
                          device_built=True, desc="pangenome-like masked superstring: 5 Mbp genome + 39 M SNP windows of 61 bases (2.38 Gbp, ~1.2 G represented 31-mers)"),
    "pangenome_k23": dict(genome_len=5_000_000, k=23, seed=4, superstring="variants", variants=52_000_000, batch=1 << 26, reads=1_000_000,
                          device_built=True, desc="pangenome-like masked superstring: 5 Mbp genome + 52 M SNP windows of 45 bases (2.35 Gbp, ~1.2 G represented 23-mers)"),
    # reduced shapes for quick checks
    "pangenome_small": dict(genome_len=1_000_000, k=31, seed=4, superstring="variants", variants=1_000_000, batch=1 << 22, reads=100_000,
                            device_built=True, desc="pangenome-like, reduced (debug)"),
    "tiny": dict(genome_len=200_000, k=31, seed=3, superstring="contigs", batch=1 << 22, reads=50_000, desc="200 kbp random genome (debug)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {r.stderr.decode(errors='replace')[-2000:]}")
    return r


def static_config(name: str) -> dict:
    """The part of `config` both arms print identically."""
    w = WORKLOADS[name]
    return dict(workload=name, desc=w["desc"], k=w["k"], genome_len=w["genome_len"], superstring=w["superstring"],
                queries="single k-mers, 50% present (uniform positions of the indexed sequence, random strand), 50% i.i.d. random",
                mode="query -O (MODE_ALL, STRANDS_LAZY)",
                l2="GPU arm: inputs larger than L2 (2 alternating batches of 512 MiB); CPU arm: FASTA shards of a bounded sample")


# ================================================================================================ workloads
def prepare_file_index(name: str) -> dict:
    """Genome -> masked superstring -> reference `fmsi index`; cached under data/<name>/ (CPU only)."""
    w = WORKLOADS[name]
    d = os.path.join(DATA, name)
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "ms.fa")
    k = w["k"]
    genome = synth.random_codes(w["genome_len"], w["seed"])
    if not os.path.exists(prefix + ".fmsi.misc") or not os.path.exists(prefix + ".fmsi.klcp"):
        if not os.path.exists(REF_FMSI):
            raise RuntimeError("oracle/_ref/fmsi is needed to build the benchmark index (index construction is the "
                               "reference's, unchanged); run __graft_entry__.build() where /root/reference exists")
        t0 = time.time()
        how = w["superstring"]
        if how == "kmercamel" and os.path.exists(REF_KMERCAMEL):
            gfa = os.path.join(d, "genome.fa")
            synth.write_fasta_single(gfa, "genome", synth.codes_to_ascii(genome))
            run([REF_KMERCAMEL, "-c", "-k", str(k), "-p", gfa, "-o", os.path.join(d, "ms.raw.fa")])
            run([REF_KMERCAMEL, "optimize", "-c", "-a", "ones", "-k", str(k), "-p", os.path.join(d, "ms.raw.fa"), "-o", prefix])
            os.remove(gfa)
            os.remove(os.path.join(d, "ms.raw.fa"))
        else:
            synth.write_fasta_single(prefix, "ms", synth.contig_superstring(genome, k, 64, w["seed"] + 100, "max"))
            how = "contigs"
        run([REF_FMSI, "index", "-k", str(k), prefix])
        with open(os.path.join(d, "how.txt"), "w") as f:
            f.write(how)
        log(f"[bench] built index {name} ({how}) in {time.time() - t0:.1f}s")
    how = open(os.path.join(d, "how.txt")).read().strip() if os.path.exists(os.path.join(d, "how.txt")) else "?"
    return dict(name=name, prefix=prefix, k=k, genome=genome, codes=None, how=how, built_by="reference `fmsi index`", **{kk: w[kk] for kk in ("batch", "desc", "genome_len", "reads")})


def device_genome(n: int, seed: int, k: int, dev):
    """Seeded i.i.d. base codes on the device + the mask-cased ASCII superstring (device)."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    codes = torch.empty(n, dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_ = torch.empty(n, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for a in range(0, n, step):
        b = min(n, a + step)
        codes[a:b] = torch.randint(0, 4, (b - a,), dtype=torch.uint8, device=dev, generator=gen)
        ascii_[a:b] = lut[codes[a:b].long()]
    ascii_[n - (k - 1):] += 32  # last k-1 letters lower case (mask convention, parser.h:31-37)
    return codes, ascii_


def device_pangenome(genome_len: int, variants: int, seed: int, k: int, dev):
    """configs[4]: base genome + `variants` SNP windows of 2k-1 bases, mask ON for the k k-mers that contain the SNP."""
    import torch
    codes, _ = device_genome(genome_len, seed, k, dev)
    gen0 = torch.Generator(device=dev)
    gen0.manual_seed(78)
    V, W = variants, 2 * k - 1
    n_total = genome_len + V * W
    sup = torch.empty(n_total, dtype=torch.uint8, device=dev)
    sup[:genome_len] = codes
    step = 1 << 22
    ar = torch.arange(W, device=dev)
    for a in range(0, V, step):
        b = min(V, a + step)
        pos = torch.randint(k - 1, genome_len - k, (b - a,), device=dev, generator=gen0)
        win = codes[(pos[:, None] + (ar[None, :] - (k - 1)))]
        shift = torch.randint(1, 4, (b - a,), device=dev, generator=gen0, dtype=torch.uint8)
        win[:, k - 1] = (win[:, k - 1] + shift) & 3
        sup[genome_len + a * W:genome_len + b * W] = win.reshape(-1)
        del win, pos, shift
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_ = torch.empty(n_total, dtype=torch.uint8, device=dev)
    for a in range(0, n_total, 1 << 28):
        b = min(n_total, a + (1 << 28))
        ascii_[a:b] = lut[sup[a:b].long()]
    # OFF positions: the k-1 k-mers that run from the genome into the first window, the k-1 that straddle two windows,
    # the last k-1 letters
    ascii_[genome_len - (k - 1):genome_len] += 32
    tail = ascii_[genome_len:].view(V, W)
    tail[:, k:] += 32
    del codes
    torch.cuda.empty_cache()
    return sup, ascii_


def prepare_device_built(name: str, dev, local_rank: int, with_klcp: bool, **build_kw):
    """Human-scale / pangenome shapes: sequence and index built on the GPU."""
    import torch
    import fmsi_b200 as fg
    w = WORKLOADS[name]
    k = w["k"]
    d = os.path.join(DATA, name)
    prefix = os.path.join(d, "ms.fa")
    t0 = time.time()
    if w["superstring"] == "variants":
        codes, ascii_ = device_pangenome(w["genome_len"], w["variants"], w["seed"], k, dev)
    else:
        codes, ascii_ = device_genome(w["genome_len"], w["seed"], k, dev)
    n = codes.numel()
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    t1 = time.time()
    gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=with_klcp, device=local_rank, n=n, mem=fg.MEM_DEVICE, **build_kw)
    t2 = time.time()
    log(f"[bench] {name}: sequence {t1 - t0:.1f}s, GPU index build {t2 - t1:.1f}s (tier {gi.dict_kind}, multistep {gi.multistep}, {gi.hbm_bytes / 1e9:.1f} GB)")
    wl = dict(name=name, prefix=prefix, k=k, genome=None, codes=codes, ascii=ascii_, how=w["superstring"] + " (GPU-built index)",
              built_by="fmsi_gpu_index_build (byte-identical to `fmsi index`: tests/test_gpu_build.py)", batch=w["batch"], reads=w["reads"],
              desc=w["desc"], genome_len=n, build_s=round(t2 - t1, 2))
    return gi, wl


def ensure_files(gi, wl: dict) -> bool:
    """The reference-format files of a device-built index, for the CPU arm and the oracle (written once)."""
    prefix = wl["prefix"]
    if os.path.exists(prefix + ".fmsi.misc"):
        return True
    t0 = time.time()
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    gi.save(prefix)
    log(f"[bench] {wl['name']}: index files saved in {time.time() - t0:.1f}s")
    return True


# ================================================================================================ queries
def device_kmer_queries(codes, k: int, batch: int, seed: int, dev, frac_present: float = 0.5):
    """Packed k-mers on the device: present ones from uniform positions of the indexed sequence on a random strand,
    absent ones i.i.d. uniform (same distribution as synth.packed_kmer_queries)."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    n = codes.numel()
    pos = torch.randint(0, n - k + 1, (batch,), device=dev, generator=gen)
    fw = torch.zeros(batch, dtype=torch.int64, device=dev)
    rc = torch.zeros(batch, dtype=torch.int64, device=dev)
    for t in range(k):
        c = codes[pos + t].long()
        fw = (fw << 2) | c
        rc = rc | ((3 - c) << (2 * t))
    flip = torch.rand(batch, device=dev, generator=gen) < 0.5
    km = torch.where(flip, rc, fw)
    rnd = torch.randint(0, 1 << (2 * k), (batch,), dtype=torch.int64, device=dev, generator=gen)
    present = torch.rand(batch, device=dev, generator=gen) < frac_present
    return torch.where(present, km, rnd)


device_queries = device_kmer_queries  # name used by the tests and the profiles/ scripts


def device_reads(codes, n_reads: int, seed: int, dev):
    """(n_reads, 150) base codes: uniform positions, random strand, 1 % substitutions (BASELINE configs[1])."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    n = codes.numel()
    L = READ_LEN
    pos = torch.randint(0, n - L + 1, (n_reads,), device=dev, generator=gen)
    rd = codes[(pos[:, None] + torch.arange(L, device=dev)[None, :])]
    flip = torch.rand(n_reads, device=dev, generator=gen) < 0.5
    rd = torch.where(flip[:, None], 3 - rd.flip(1), rd)
    sub = torch.rand(n_reads, L, device=dev, generator=gen) < 0.01
    shift = torch.randint(1, 4, (n_reads, L), device=dev, generator=gen, dtype=torch.uint8)
    return torch.where(sub, (rd + shift) & 3, rd).contiguous()


def reads_layout(n_reads: int, k: int):
    """Chunks of <= 64 k-mers overlapping by k-1 (what the -S kernel takes), results back to back in read order."""
    nk = READ_LEN - k + 1
    pieces, p = [], 0
    while p < nk:
        m = min(64, nk - p)
        pieces.append((p, m))
        p += m
    base = np.arange(n_reads, dtype=np.uint64) * np.uint64(READ_LEN)
    rbase = np.arange(n_reads, dtype=np.uint64) * np.uint64(nk)
    off = np.stack([base + np.uint64(p) for p, m in pieces], 1).reshape(-1)
    ln = np.tile(np.array([m + k - 1 for p, m in pieces], dtype=np.uint32), n_reads)
    roff = np.stack([rbase + np.uint64(p) for p, m in pieces], 1).reshape(-1)
    return off, ln, roff, n_reads * nk


def pack_codes_device(rd):
    """(R, 150) codes on the device -> FMSI_GPU_TEXT_PACKED2 words of the concatenated text (device int64 tensor)."""
    import torch
    flat = rd.reshape(-1)
    n = flat.numel()
    nw = (n + 31) // 32
    pad = torch.zeros(nw * 32, dtype=torch.int64, device=flat.device)
    pad[:n] = flat.long()
    sh = (62 - 2 * torch.arange(32, device=flat.device, dtype=torch.int64))
    return (pad.view(nw, 32) << sh[None, :]).sum(1)  # fields do not overlap: sum == or (wraps into the sign bit as bits)


# ================================================================================================ instruments
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "100", "-f", self.path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def stored_traffic(kernel: str) -> dict | None:
    """DRAM bytes per k-mer of a kernel from the committed ncu captures (profiles/traffic.json; per-kernel, with the
    commit and workload they were taken on). Only a capture of the same kernel on the same workload shape is used."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get("kernels", {}).get(kernel)


def roofline_for(kernel: str, kmers_per_launch: int, launch_ms: float, in_bytes: float, out_bytes: float, probes_per_kmer: float | None,
                 peak: float, peak_src: str, ref_alg_bytes: float | None = None, traffic_key: str | None = None) -> dict:
    """Roofline of the timed kernel from ITS OWN algorithmic bytes: query in + result out + 32 B per dependent probe."""
    if probes_per_kmer is None:
        return dict(bound="hbm", kernel=kernel, achieved=None, peak=peak, unit="GB/s", frac=None, traffic=None, note="probe accounting unavailable")
    per = in_bytes + out_bytes + 32.0 * probes_per_kmer
    secs = launch_ms / 1e3
    achieved = per * kmers_per_launch / secs / 1e9
    r = dict(bound="hbm", kernel=kernel, achieved=round(achieved, 1), peak=peak, unit="GB/s", frac=round(achieved / peak, 4), traffic=None,
             launch_ms=round(launch_ms, 4), kmers_per_launch=int(kmers_per_launch), peak_source=peak_src,
             algorithmic=dict(bytes_per_kmer=round(per, 2), query_in=in_bytes, result_out=out_bytes, probes_per_kmer=round(probes_per_kmer, 3),
                              bytes_per_probe=32, how="probes counted live by the kernel (fmsi_gpu_count_probes) on this run's batch"),
             request_rate=dict(gprobes_s=round(probes_per_kmer * kmers_per_launch / secs / 1e9, 2),
                               random_request_ceiling_gprobes_s=RANDOM_REQUEST_CEILING / 1e9,
                               frac_of_ceiling=round(probes_per_kmer * kmers_per_launch / secs / RANDOM_REQUEST_CEILING, 3),
                               note="ceiling = dependent random 32-B reads that miss L2, measured with tools/randbw2.cu on this part (DRAM "
                                    "activate rate; the same whether a request returns 8 or 128 bytes); L2 hits let a kernel exceed it"))
    tj = stored_traffic(traffic_key or kernel)
    if tj:
        traffic = tj["dram_bytes_per_kmer"] * kmers_per_launch
        r["traffic"] = traffic
        r["traffic_source"] = tj.get("source")
        r["wasted"] = round(tj["dram_bytes_per_kmer"] / per, 3)
        r["dram_gbs"] = round(traffic / secs / 1e9, 1)
        r["dram_frac"] = round(traffic / secs / 1e9 / peak, 4)
    if ref_alg_bytes:
        r["vs_reference_algorithm"] = dict(bytes_per_kmer=round(ref_alg_bytes, 1), frac=round(ref_alg_bytes * kmers_per_launch / secs / 1e9 / peak, 4),
                                           note="SURVEY 8(d): the REFERENCE algorithm's dependent sector probes per k-mer (32 B per rank / mask probe "
                                                "+ 8 B query + 1 B result), counted by the instrumented oracle; > 1 means the kernel does not do that work")
    return r


# ================================================================================================ CPU arm
def host_kmer_sample(wl: dict, n: int, seed: int) -> np.ndarray:
    """n packed query k-mers of the workload's distribution, in host memory."""
    if wl.get("codes") is not None:
        q = device_kmer_queries(wl["codes"], wl["k"], n, seed, wl["codes"].device)
        return q.cpu().numpy().view(np.uint64)
    if wl.get("_gk") is None:
        wl["_gk"] = synth.pack_kmers(wl["genome"], wl["k"])
    return synth.packed_kmer_queries(wl["_gk"], wl["k"], n, seed)


def host_read_sample(wl: dict, n_reads: int, seed: int) -> np.ndarray:
    if wl.get("codes") is not None:
        return device_reads(wl["codes"], n_reads, seed, wl["codes"].device).cpu().numpy()
    return synth.read_queries(wl["genome"], READ_LEN, n_reads, seed)


def reads_to_fasta(reads: np.ndarray) -> bytes:
    n, L = reads.shape
    rec = np.empty((n, L + 4), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    rec[:, 2] = ord("\n")
    rec[:, 3:3 + L] = synth.ACGT[reads]
    rec[:, L + 3] = ord("\n")
    return rec.tobytes()


def reference_cpu_rate(prefix: str, k: int, ref_args: list[str], shards: list[bytes], units_per_shard: int, what: str) -> dict:
    """The reference's own CPU query path: P independent `fmsi <ref_args>` processes over FASTA shards (the reference has
    no threads). Wall time from first start to last exit; the per-process index load is measured separately with a
    one-record query file, reported, and subtracted for `value` (the GPU arm's numbers exclude its index set-up too)."""
    if not os.path.exists(REF_FMSI):
        raise RuntimeError("oracle/_ref/fmsi missing")
    P = len(shards)
    tmp = tempfile.mkdtemp(prefix="fmsi_ref_")
    try:
        files = []
        for p, blob in enumerate(shards):
            fn = os.path.join(tmp, f"q{p}.fa")
            with open(fn, "wb") as f:
                f.write(blob)
            files.append(fn)
        one = os.path.join(tmp, "one.fa")
        with open(one, "wb") as f:
            f.write(b">q\n" + b"A" * max(k, 1) + b"\n")
        t0 = time.perf_counter()
        run([REF_FMSI] + ref_args + ["-q", one, prefix])
        load_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ps = [subprocess.Popen([REF_FMSI] + ref_args + ["-q", fn, prefix], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for fn in files]
        for p_ in ps:
            if p_.wait() != 0:
                raise RuntimeError("reference fmsi failed")
        wall = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    total = P * units_per_shard
    query_wall = max(wall - load_s, 1e-9)
    return dict(value=total / query_wall, unit=UNIT, cores=P, kind="reference", wall_s=round(wall, 3), index_load_s=round(load_s, 3),
                query_wall_s=round(query_wall, 3), value_including_load=total / wall, kmers=total,
                sample=f"{P} concurrent `fmsi {' '.join(ref_args)}` processes (unmodified reference binary, oracle/_ref) x {what}, same index "
                       f"files; wall {wall:.2f}s of which per-process index load {load_s:.2f}s (subtracted)")


def cpu_kmers(wl: dict, ref_args: list[str], per_proc: int, seed: int, procs: int | None = None) -> dict:
    P = procs or os.cpu_count() or 1
    allq = host_kmer_sample(wl, P * per_proc, seed)
    shards = [synth.packed_to_fasta(allq[p * per_proc:(p + 1) * per_proc], wl["k"]) for p in range(P)]
    return reference_cpu_rate(wl["prefix"], wl["k"], ref_args, shards, per_proc, f"{per_proc} single {wl['k']}-mer FASTA records (50% present)")


def cpu_reads(wl: dict, ref_args: list[str], reads_per_proc: int, seed: int, procs: int | None = None) -> dict:
    P = procs or os.cpu_count() or 1
    rd = host_read_sample(wl, P * reads_per_proc, seed)
    shards = [reads_to_fasta(rd[p * reads_per_proc:(p + 1) * reads_per_proc]) for p in range(P)]
    nk = READ_LEN - wl["k"] + 1
    return reference_cpu_rate(wl["prefix"], wl["k"], ref_args, shards, reads_per_proc * nk, f"{reads_per_proc} reads of {READ_LEN} bp, 1% substitutions ({nk} k-mers each)")


def load_oracle(wl: dict, use_klcp: bool):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ffi import OracleIndex
    t0 = time.time()
    oi = OracleIndex.load(wl["prefix"], use_klcp=use_klcp)
    log(f"[bench] {wl['name']}: oracle loaded in {time.time() - t0:.1f}s")
    return oi


def reference_algorithm_bytes(oi, wl: dict, q: np.ndarray) -> tuple[dict, np.ndarray]:
    """32 B x (rank sectors + mask sectors) per k-mer of the REFERENCE's backward search, counted by the
    instrumented oracle on a sample (forward strand first, neutral predictor), + 8 B packed query in + 1 B result out."""
    oi.counters_reset()
    want = oi.query_packed(q, wl["k"], 1, False)
    c = oi.counters()
    n = c["kmers"]
    per = dict(lf_steps=c["lf_steps"] / n, rank_sectors=c["rank_sectors"] / n, mask_sectors=c["mask_sectors"] / n)
    per["bytes"] = 32.0 * (per["rank_sectors"] + per["mask_sectors"]) + 8 + 1
    return per, want


# ================================================================================================ timing
class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            self.dev = torch.device("cuda", self.local_rank)
            torch.cuda.set_device(self.dev)
            self.stream = torch.cuda.current_stream(self.dev)
        else:  # the bookkeeping (process group, gathers) is testable without a device: tests/test_shard.py
            self.dev, self.stream = torch.device("cpu"), None
        self.dist = None
        self.peak, self.peak_src = measured_peak_gbs()

    def init_dist(self):
        if self.world == 1:
            return
        import torch.distributed as dist
        # Bookkeeping only (barriers, max of the elapsed time, gathering the parity samples): the query path has no
        # collective, so the group runs over gloo; FMSI_BENCH_DIST_BACKEND=nccl selects NCCL.
        backend = os.environ.get("FMSI_BENCH_DIST_BACKEND", "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=self.dev)
        else:
            if os.environ.get("MASTER_ADDR", "127.0.0.1") in ("127.0.0.1", "localhost", "::1"):
                os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            try:
                dist.init_process_group(backend)
            except Exception as ex:
                log(f"[bench] {backend} process group failed ({ex}); retrying over the loopback interface")
                os.environ["GLOO_SOCKET_IFNAME"] = "lo"
                try:
                    dist.init_process_group(backend)
                except Exception as ex2:
                    log(f"[bench] {backend} failed again ({ex2}); falling back to nccl")
                    dist.init_process_group("nccl", device_id=self.dev)
        self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def sync_all(self):
        if self.stream is None:
            return self.barrier()
        self.torch.cuda.synchronize(self.dev)
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev if self.dist.get_backend() == "nccl" else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_to_rank0(self, arr: np.ndarray) -> list[np.ndarray] | None:
        """Parity samples of every rank on rank 0 (bookkeeping, outside every timed region)."""
        if not self.dist:
            return [arr]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(arr, out, dst=0)
        return out

    def time_device(self, step, steps: int, warmup: int) -> float:
        """ms per step: CUDA events on the launching stream around `steps` calls, barrier + synchronize on both
        sides, max over ranks. One event pair: a per-step event cost 0.12 ms per step on its own."""
        torch = self.torch
        for s_ in range(warmup):
            step(s_)
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for s_ in range(steps):
            step(s_)
        e1.record(self.stream)
        self.sync_all()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    def time_host(self, step, steps: int, warmup: int = 3) -> float:
        """seconds per step of a blocking host-buffer call (wall clock, device idle on both sides), max over ranks."""
        for s_ in range(warmup):
            step(s_)
        self.sync_all()
        t0 = time.perf_counter()
        for s_ in range(steps):
            step(s_)
        self.torch.cuda.synchronize(self.dev)
        return self.max_over_ranks(time.perf_counter() - t0) / steps


def probes_per_kmer(gi, fn, n_kmers: int) -> float | None:
    """One untimed, counted pass of `fn` (the timed step): dependent requests per k-mer."""
    try:
        before = gi.count_probes(True)
        fn()
        after = gi.count_probes(False)
        return (after - before) / n_kmers
    except Exception as ex:  # pragma: no cover
        log(f"[bench] probe accounting failed: {ex}")
        return None


def kernel_name(gi, reads: bool, streaming: bool, suffix: str) -> str:
    if gi.dict_kind == 2:
        return f"fold_query_kernel<{suffix}>"
    if gi.dict_kind == 1:
        return f"dict_query_kernel<{suffix}>"
    if reads and streaming:
        return f"stream_kernel<{suffix}>"
    return f"query_kmers_kernel<{suffix}>"


def bench_kmers(cx: Ctx, gi, wl: dict, mode: int, output: int, label: str, batch: int, steps: int, warmup: int, seed: int, e2e_bits: bool,
                ref_alg_bytes: float | None = None) -> dict:
    """Single packed k-mers: device-resident value (two alternating batches larger than L2), e2e from pinned host
    buffers, roofline of the timed kernel."""
    import fmsi_b200 as fg
    torch = cx.torch
    k = wl["k"]
    nbuf = 2
    if wl.get("codes") is not None:
        d_in = [device_kmer_queries(wl["codes"], k, batch, seed + 17 * cx.rank + b, cx.dev) for b in range(nbuf)]
        pinned_in = [t.cpu().pin_memory() for t in d_in]
    else:
        if wl.get("_gk") is None:
            wl["_gk"] = synth.pack_kmers(wl["genome"], k)
        host = [synth.packed_kmer_queries(wl["_gk"], k, batch, seed=seed + 17 * cx.rank + b) for b in range(nbuf)]
        pinned_in = [torch.from_numpy(h.view(np.int64)).pin_memory() for h in host]
        d_in = [p.to(cx.dev) for p in pinned_in]
    rb = 1 if output == fg.OUT_PRESENCE else 8
    d_out = torch.empty(batch * rb, dtype=torch.uint8, device=cx.dev)
    out_e2e = fg.OUT_PRESENCE_BITS if (e2e_bits and output == fg.OUT_PRESENCE) else output
    e2e_out_bytes = (batch + 7) // 8 if out_e2e == fg.OUT_PRESENCE_BITS else batch * rb
    pinned_out = torch.empty(e2e_out_bytes, dtype=torch.uint8).pin_memory()

    def step_device(s_):
        gi.query_kmers_ptr(d_in[s_ % nbuf].data_ptr(), batch, d_out.data_ptr(), k, mode, output, fg.STRANDS_LAZY, fg.MEM_DEVICE, cx.stream.cuda_stream)

    def step_host(s_):
        gi.query_kmers_ptr(pinned_in[s_ % nbuf].data_ptr(), batch, pinned_out.data_ptr(), k, mode, out_e2e, fg.STRANDS_LAZY, fg.MEM_HOST, 0)

    ms = cx.time_device(step_device, steps, warmup)
    e2e_s = cx.time_host(step_host, steps)
    # the device-resident and the host-buffer path give the same answers
    launches0 = fg.launch_count()
    step_device(steps - 1)
    launches = (fg.launch_count() - launches0) * steps  # kernels of this library inside the timed region
    torch.cuda.synchronize(cx.dev)
    dev_res = d_out.cpu().numpy()
    if out_e2e == fg.OUT_PRESENCE_BITS:
        same = bool(np.array_equal(np.packbits(dev_res, bitorder="little"), pinned_out.numpy()))
    else:
        same = bool(np.array_equal(dev_res, pinned_out.numpy()))
    frac_present = float((dev_res != 0).mean()) if output == fg.OUT_PRESENCE else float((dev_res.view(np.int64) >= 0).mean())
    ppk = probes_per_kmer(gi, lambda: step_device(0), batch)
    suffix = {(fg.MODE_ALL, fg.OUT_PRESENCE): "ALL,PRESENCE,LAZY", (fg.MODE_OR, fg.OUT_PRESENCE): "OR,PRESENCE,LAZY",
              (fg.MODE_OR, fg.OUT_ORDERS): "OR,ORDERS,LAZY"}[(mode, output)]
    kern = kernel_name(gi, False, False, suffix)
    res = dict(mode=label, value=cx.world * batch / (ms / 1e3), unit=UNIT, ms_per_step=ms, kmers_per_step_per_gpu=batch,
               tier={2: "strand-folded dictionary", 1: "SA-ordered dictionary", 0: "backward search"}[gi.dict_kind], multistep=gi.multistep,
               prefix_t=gi.prefix_t, index_hbm_bytes=gi.hbm_bytes, frac_present=round(frac_present, 4), e2e_equals_device=same,
               e2e=dict(value=cx.world * batch / e2e_s, unit=UNIT, h2d_bytes_per_step=batch * 8, d2h_bytes_per_step=e2e_out_bytes,
                        h2d_gbs_per_gpu=round(batch * 8 / e2e_s / 1e9, 1),
                        note="packed k-mers cross PCIe at 8 B each: the host link, not the kernel, bounds this path"),
               roofline=roofline_for(kern, batch, ms, 8, rb, ppk, cx.peak, cx.peak_src, ref_alg_bytes,
                                     traffic_key=f"{kern}@{wl['name']}" + (f",ms{gi.multistep}" if gi.dict_kind == 0 else "")),
               gpu_launches=int(launches))
    del d_in, d_out, pinned_in, pinned_out
    torch.cuda.empty_cache()
    return res


def bench_reads(cx: Ctx, gi, wl: dict, mode: int, output: int, streaming: bool, label: str, n_reads: int, steps: int, warmup: int, seed: int) -> tuple[dict, dict]:
    """150-bp reads. Device-resident: 2-bit packed text + read offsets in HBM (two alternating read sets). e2e:
    fmsi_gpu_query_reads_packed from pinned host memory, presence bits back (ids for lookup)."""
    import fmsi_b200 as fg
    torch = cx.torch
    k = wl["k"]
    L_ = fg.lib()
    n_res = n_reads * (READ_LEN - k + 1)
    nbuf = 2
    codes = wl["codes"] if wl.get("codes") is not None else torch.from_numpy(wl["genome"]).to(cx.dev)
    rds = [device_reads(codes, n_reads, seed + 31 * cx.rank + b, cx.dev) for b in range(nbuf)]
    d_text = [pack_codes_device(r) for r in rds]
    n_bases = n_reads * READ_LEN
    rb = 1 if output == fg.OUT_PRESENCE else 8
    d_out = torch.empty(n_res * rb, dtype=torch.uint8, device=cx.dev)
    p_text = [t.cpu().pin_memory() for t in d_text]
    out_e2e = fg.OUT_PRESENCE_BITS if output == fg.OUT_PRESENCE else output
    e2e_out_bytes = (n_res + 7) // 8 if out_e2e == fg.OUT_PRESENCE_BITS else n_res * rb
    p_out = torch.empty(e2e_out_bytes, dtype=torch.uint8).pin_memory()

    # whole reads in (fmsi_gpu_query_reads_packed): 2-bit text + one offset per read; the device cuts them into chunks
    read_off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    d_read_off = torch.from_numpy(read_off.view(np.int64)).to(cx.dev)
    p_read_off = torch.from_numpy(read_off.view(np.int64)).pin_memory()

    def call(text_ptr, roff_ptr, out_ptr, outk, mem, stream):
        rc = L_.fmsi_gpu_query_reads_packed(gi._h, mode, outk, fg.STRANDS_LAZY, int(streaming), text_ptr, n_bases, roff_ptr, n_reads, n_res, k, out_ptr, mem, stream)
        if rc != 0:
            raise RuntimeError(L_.fmsi_gpu_last_error().decode())

    def step_device(s_):
        call(d_text[s_ % nbuf].data_ptr(), d_read_off.data_ptr(), d_out.data_ptr(), output, fg.MEM_DEVICE, cx.stream.cuda_stream)

    def step_host(s_):
        call(p_text[s_ % nbuf].data_ptr(), p_read_off.data_ptr(), p_out.data_ptr(), out_e2e, fg.MEM_HOST, None)

    ms = cx.time_device(step_device, steps, warmup)
    e2e_s = cx.time_host(step_host, steps)
    launches0 = fg.launch_count()
    step_device(steps - 1)
    launches = (fg.launch_count() - launches0) * steps  # kernels of this library inside the timed region
    torch.cuda.synchronize(cx.dev)
    dev_res = d_out.cpu().numpy()
    if out_e2e == fg.OUT_PRESENCE_BITS:
        same = bool(np.array_equal(np.packbits(dev_res, bitorder="little"), p_out.numpy()))
    else:
        same = bool(np.array_equal(dev_res, p_out.numpy()))
    frac_present = float((dev_res != 0).mean()) if output == fg.OUT_PRESENCE else float((dev_res.view(np.int64) >= 0).mean())
    ppk = probes_per_kmer(gi, lambda: step_device(0), n_res)
    suffix = {(fg.MODE_ALL, fg.OUT_PRESENCE): "ALL,PRESENCE,LAZY", (fg.MODE_OR, fg.OUT_PRESENCE): "OR,PRESENCE,LAZY",
              (fg.MODE_OR, fg.OUT_ORDERS): "OR,ORDERS,LAZY"}[(mode, output)]
    kern = kernel_name(gi, True, streaming, suffix)
    in_bytes_per_kmer = n_bases / 4 / n_res  # 2 bits per base
    h2d = int(p_text[0].numel() * 8 + read_off.nbytes)
    # the timed step = the query kernel plus, when a dictionary tier answers the chunks, the slot -> k-mer extraction
    res = dict(mode=label, value=cx.world * n_res / (ms / 1e3), unit=UNIT, ms_per_step=ms, kmers_per_step_per_gpu=n_res, reads_per_step_per_gpu=n_reads,
               tier={2: "strand-folded dictionary", 1: "SA-ordered dictionary", 0: "backward search"}[gi.dict_kind], multistep=gi.multistep,
               prefix_t=gi.prefix_t, index_hbm_bytes=gi.hbm_bytes, frac_present=round(frac_present, 4), e2e_equals_device=same,
               e2e=dict(value=cx.world * n_res / e2e_s, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=e2e_out_bytes,
                        h2d_gbs_per_gpu=round(h2d / e2e_s / 1e9, 2), bytes_per_kmer=round((h2d + e2e_out_bytes) / n_res, 3),
                        note="fmsi_gpu_query_reads_packed: 2-bit packed text + one offset per read in, " + ("presence bits" if out_e2e == fg.OUT_PRESENCE_BITS else "int64 ids") + " out"),
               roofline=roofline_for(kern, n_res, ms, in_bytes_per_kmer, rb, ppk, cx.peak, cx.peak_src,
                                     traffic_key=f"{kern}@{wl['name']},reads" + (f",ms{gi.multistep}" if gi.dict_kind == 0 else "")),
               gpu_launches=int(launches))
    res["roofline"]["note"] = ("launch_ms is the whole step on the stream (text copy, slot -> k-mer extraction where a dictionary tier answers the "
                               "chunks, query kernel); the algorithmic bytes are the query kernel's")
    sample = dict(reads=rds[(steps - 1) % nbuf][:2000].cpu().numpy(), results=dev_res.view(np.int64 if rb == 8 else np.uint8)[:2000 * (READ_LEN - k + 1)].copy())
    del d_text, d_out, rds, p_text, p_out
    torch.cuda.empty_cache()
    return res, sample


def parity_kmers(cx: Ctx, gi, oi, wl: dict, n: int, seed: int, mode: int, output: int, omode: int, oord: bool) -> bool | None:
    """A sample of EVERY rank's answers (through the host-buffer C-ABI path) against the oracle on rank 0."""
    q = host_kmer_sample(wl, n, seed + 1000 * cx.rank)
    got = gi.query_kmers(q, wl["k"], mode, output).astype(np.int64)
    gathered = cx.gather_to_rank0(np.stack([q.view(np.int64), got]))
    if cx.rank != 0 or oi is None:
        return None
    ok = True
    for g in gathered:
        want = oi.query_packed(g[0].view(np.uint64), wl["k"], omode, oord)
        ok = ok and bool(np.array_equal(g[1], want))
    return ok


def parity_reads(cx: Ctx, oi, wl: dict, sample: dict, omode: int, oord: bool) -> bool | None:
    gathered = cx.gather_to_rank0(np.concatenate([sample["reads"].reshape(-1).astype(np.int64), sample["results"].astype(np.int64)]))
    if cx.rank != 0 or oi is None:
        return None
    k = wl["k"]
    nk = READ_LEN - k + 1
    ok = True
    for g in gathered:
        nr = g.size // (READ_LEN + nk)
        reads = g[:nr * READ_LEN].reshape(nr, READ_LEN).astype(np.uint8)
        res = g[nr * READ_LEN:]
        allk = np.concatenate([synth.pack_kmers(r, k) for r in reads])
        want = oi.query_packed(allk, k, omode, oord)
        ok = ok and bool(np.array_equal(res, want if oord else (want == 1).astype(np.int64)))
    return ok


# ================================================================================================ CLI arm
def sha_file(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def cli_like_for_like(wl: dict, n_records: int, seed: int, ref_args: list[str], device: int, reads: bool = False) -> dict:
    """The drop-in claim: this repo's `fmsi` binary against the reference binary, whole process (index load included),
    same FASTA (single-k-mer records, or 150-bp reads with 1 % substitutions), stdout compared byte for byte. The
    reference is single-threaded, so its run is P concurrent processes over record shards (each loads the index) and
    the concatenated outputs are compared."""
    if not (os.path.exists(OUR_FMSI) and os.path.exists(REF_FMSI)):
        return dict(error="fmsi binaries missing")
    P = os.cpu_count() or 1
    per = (n_records + P - 1) // P
    q = host_read_sample(wl, per * P, seed) if reads else host_kmer_sample(wl, per * P, seed)
    tmp = tempfile.mkdtemp(prefix="fmsi_cli_")
    try:
        whole = os.path.join(tmp, "all.fa")
        shards = []
        with open(whole, "wb") as fw:
            for p in range(P):
                blob = reads_to_fasta(q[p * per:(p + 1) * per]) if reads else synth.packed_to_fasta(q[p * per:(p + 1) * per], wl["k"])
                fw.write(blob)
                fn = os.path.join(tmp, f"s{p}.fa")
                with open(fn, "wb") as f:
                    f.write(blob)
                shards.append(fn)
        env = dict(os.environ, FMSI_GPU_DEVICE=str(device), FMSI_GPU_TIMING="1")
        ours_out = os.path.join(tmp, "ours.txt")
        t0 = time.perf_counter()
        with open(ours_out, "wb") as fo:
            r = subprocess.run([OUR_FMSI] + ref_args + ["-q", whole, wl["prefix"]], stdout=fo, stderr=subprocess.PIPE, env=env)
        ours_s = time.perf_counter() - t0
        if r.returncode != 0:
            return dict(error="our fmsi failed: " + r.stderr.decode(errors="replace")[-500:])
        t0 = time.perf_counter()
        outs = [open(os.path.join(tmp, f"r{p}.txt"), "wb") for p in range(P)]
        ps = [subprocess.Popen([REF_FMSI] + ref_args + ["-q", fn, wl["prefix"]], stdout=o, stderr=subprocess.DEVNULL) for fn, o in zip(shards, outs)]
        rcs = [p_.wait() for p_ in ps]
        ref_s = time.perf_counter() - t0
        for o in outs:
            o.close()
        if any(rcs):
            return dict(error="reference fmsi failed")
        h = hashlib.sha256()
        size = 0
        for p in range(P):
            with open(os.path.join(tmp, f"r{p}.txt"), "rb") as f:
                for blk in iter(lambda: f.read(1 << 24), b""):
                    h.update(blk)
                    size += len(blk)
        whole_equals_concat = h.hexdigest() == sha_file(ours_out) and size == os.path.getsize(ours_out)
        # Identity proper: the same input through one process of either binary (shard 0). The whole-file run equals the
        # concatenation of P independent reference processes only when no answer depends on the strand predictor's
        # history — `lookup` ids >= 2^31 do: the reference narrows the id to `int` for its predictor
        # (fms_index.h:282), takes the wrapped value for "absent" and asks the other strand, so what it prints for
        # such a k-mer depends on the queries before it. The CLI replays that state machine, one history per process.
        shard_out = os.path.join(tmp, "ours_s0.txt")
        with open(shard_out, "wb") as fo:
            r0 = subprocess.run([OUR_FMSI] + ref_args + ["-q", shards[0], wl["prefix"]], stdout=fo, stderr=subprocess.PIPE, env=env)
        if r0.returncode != 0:
            return dict(error="our fmsi failed on shard 0: " + r0.stderr.decode(errors="replace")[-500:])
        identical = sha_file(shard_out) == sha_file(os.path.join(tmp, "r0.txt"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    n = per * P
    kmers_total = n * (READ_LEN - wl["k"] + 1) if reads else n
    stages = {}
    for line in r.stderr.decode(errors="replace").splitlines():  # the CLI's own stage clock ($FMSI_GPU_TIMING)
        if line.startswith("[fmsi timing] index load + replicas:"):
            stages["index_load_s"] = float(line.split(":")[1].split()[0])
        elif line.startswith("[fmsi timing] total:"):
            stages["total_s"] = float(line.split(":")[1].split()[0])
    if len(stages) == 2:
        stages["kmers_s_after_load"] = kmers_total / max(stages["total_s"] - stages["index_load_s"], 1e-9)
        stages["process_start_and_exit_s"] = round(ours_s - stages["total_s"], 3)
    return dict(command="fmsi " + " ".join(ref_args), records=n, kmers=kmers_total, input="150-bp reads, 1% substitutions" if reads else "single k-mer records",
                ours_wall_s=round(ours_s, 3), ours_kmers_s=kmers_total / ours_s, ours_stages=stages,
                reference_wall_s=round(ref_s, 3), reference_kmers_s=kmers_total / ref_s, reference_processes=P, speedup=round(ref_s / ours_s, 2),
                outputs_byte_identical=bool(identical), identity_checked_on=f"shard 0 ({per} records), one process of either binary",
                whole_run_equals_concatenated_shards=bool(whole_equals_concat),
                note="whole-process wall time, index load and CUDA context creation included on our side, one index load per process "
                     "on the reference's; one GPU against P host cores")


# ================================================================================================ arms
def reference_arm(args) -> int:
    """`--impl reference`: the reference's CPU path on the box's host cores, same metric and config. For the device-built
    shapes the index files come from this repo's GPU builder, run in a CHILD process so that the timed process holds
    nothing of ours (the files are byte-identical to `fmsi index` output: tests/test_gpu_build.py)."""
    w = WORKLOADS[args.workload]
    if w.get("device_built"):
        prefix = os.path.join(DATA, args.workload, "ms.fa")
        if not os.path.exists(prefix + ".fmsi.misc") or not os.path.exists(os.path.join(DATA, args.workload, "cpu_sample_kmers.npy")):
            t0 = time.time()
            run([sys.executable, os.path.abspath(__file__), "--prepare-only", "--workload", args.workload])
            log(f"[bench] reference arm: index files prepared by a child process in {time.time() - t0:.1f}s")
        wl = dict(name=args.workload, prefix=prefix, k=w["k"], genome=None, codes=None, desc=w["desc"])
        # queries for the CPU arm without a GPU in this process: k-mers from the 2-bit text cannot be regenerated here,
        # so the child wrote a query sample next to the index (same generator as the GPU arm)
        qfile = os.path.join(DATA, args.workload, "cpu_sample_kmers.npy")
        allq = np.load(qfile)
    else:
        wl = prepare_file_index(args.workload)
        allq = None
    P = os.cpu_count() or 1
    per = args.cpu_sample or (200_000 if w.get("device_built") else 1_000_000)
    if allq is not None:
        per = min(per, len(allq) // (2 * P))
    vals = []
    for s_ in range(args.warmup + args.steps):
        if allq is not None:
            need = P * per
            start = (s_ * need) % max(1, len(allq) - need)
            qs = allq[start:start + need]
            shards = [synth.packed_to_fasta(qs[p * per:(p + 1) * per], wl["k"]) for p in range(P)]
            r = reference_cpu_rate(wl["prefix"], wl["k"], ["query", "-O"], shards, per, f"{per} single {wl['k']}-mer FASTA records (50% present)")
        else:
            r = cpu_kmers(wl, ["query", "-O"], per, seed=5000 + 100 * s_)
        if s_ >= args.warmup:
            vals.append(r)
    wall = sum(v["query_wall_s"] for v in vals)
    total = sum(v["kmers"] for v in vals)
    value = total / wall
    cb = dict(vals[-1])
    cb["value"] = value
    cb["index_files"] = "written by this repo's GPU builder in a child process" if w.get("device_built") else "reference `fmsi index`"
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * wall / len(vals), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64", data="synthetic",
                config=static_config(args.workload), step_kmers=P * per, cpu_baseline=cb,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))
    return 0


def prepare_only(args) -> int:
    """Child of the reference arm: build the index on the GPU, save the reference-format files and a query sample."""
    cx = Ctx(args)
    w = WORKLOADS[args.workload]
    prefix = os.path.join(DATA, args.workload, "ms.fa")
    if os.path.exists(prefix + ".fmsi.misc"):  # files already there: only the query sample (same generator, same sequence)
        if w["superstring"] == "variants":
            codes, _ = device_pangenome(w["genome_len"], w["variants"], w["seed"], w["k"], cx.dev)
        else:
            codes, _ = device_genome(w["genome_len"], w["seed"], w["k"], cx.dev)
        wl = dict(name=args.workload, k=w["k"], codes=codes)
    else:
        gi, wl = prepare_device_built(args.workload, cx.dev, cx.local_rank, with_klcp=True, dict=0, multistep=0, prefix_t=0)
        ensure_files(gi, wl)
        gi.close()
    q = host_kmer_sample(wl, 1 << 24, 424242)
    os.makedirs(os.path.join(DATA, args.workload), exist_ok=True)
    np.save(os.path.join(DATA, args.workload, "cpu_sample_kmers.npy"), q)
    return 0


def safe(what: str, fn):
    try:
        t0 = time.time()
        r = fn()
        log(f"[bench] {what}: {time.time() - t0:.1f}s")
        return r
    except Exception as ex:
        log(f"[bench] {what} FAILED: {ex}\n{traceback.format_exc()}")
        return dict(error=f"{type(ex).__name__}: {ex}"[:500])


def run_workload_modes(cx: Ctx, name: str, which: list[str], steps: int, warmup: int, want_cpu: bool, out: dict, tiers_out: dict | None = None):
    """All requested modes of one workload on one index replica per rank; fills out[<name>_<mode>]."""
    import fmsi_b200 as fg
    torch = cx.torch
    w = WORKLOADS[name]
    t_setup = time.time()
    if w.get("device_built"):
        gi, wl = prepare_device_built(name, cx.dev, cx.local_rank, with_klcp=True)
        wl.pop("ascii", None)
        torch.cuda.empty_cache()
    else:
        if cx.rank == 0:
            wl = prepare_file_index(name)
        cx.barrier()
        if cx.rank != 0:
            wl = prepare_file_index(name)
        gi = fg.Index.load(wl["prefix"], use_klcp=True, device=cx.local_rank)
    setup_s = time.time() - t_setup
    oi = None
    if cx.rank == 0 and w.get("device_built") and (want_cpu or not cx.args.no_parity):
        safe(f"{name}: index files", lambda: ensure_files(gi, wl))
    if cx.rank == 0 and not cx.args.no_parity:
        oi = safe(f"{name}: oracle load", lambda: load_oracle(wl, True))
        if isinstance(oi, dict):
            oi = None
    P = os.cpu_count() or 1
    big = w["genome_len"] >= 100_000_000 or w.get("variants", 0) >= 10_000_000
    batch = cx.args.batch or wl["batch"]
    n_reads = wl["reads"]

    def finish(key, res, parity, cpu):
        res["index_setup_s"] = round(setup_s, 2)
        res["parity_vs_oracle_sample"] = parity
        res["cpu_baseline"] = cpu
        res["config"] = dict(workload=name, desc=w["desc"], k=w["k"], n_bwt=gi.n, l2="two alternating input batches, each larger than L2")
        out[key] = res

    for m in which:
        key = f"{name}_{m}"
        if m == "query_O":
            res = safe(key, lambda: bench_kmers(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O, single k-mers", batch, steps, warmup, 1000, True))
            if "error" in res:
                out[key] = res
                continue
            par = safe(key + " parity", lambda: parity_kmers(cx, gi, oi, wl, 100_000, 77, fg.MODE_ALL, fg.OUT_PRESENCE, 1, False))
            cpu = safe(key + " cpu", lambda: cpu_kmers(wl, ["query", "-O"], 100_000 if big else 300_000, 9000)) if want_cpu and cx.rank == 0 else None
            finish(key, res, par, cpu)
        elif m == "lookup":
            res = safe(key, lambda: bench_kmers(cx, gi, wl, fg.MODE_OR, fg.OUT_ORDERS, "lookup, single k-mers", batch, steps, warmup, 2000, False))
            if "error" in res:
                out[key] = res
                continue
            par = safe(key + " parity", lambda: parity_kmers(cx, gi, oi, wl, 100_000, 78, fg.MODE_OR, fg.OUT_ORDERS, 0, True))
            cpu = safe(key + " cpu", lambda: cpu_kmers(wl, ["lookup"], 100_000 if big else 300_000, 9100)) if want_cpu and cx.rank == 0 else None
            finish(key, res, par, cpu)
        elif m in ("reads_S", "reads_lookup_S"):
            lk = m == "reads_lookup_S"
            mode, outp = (fg.MODE_OR, fg.OUT_ORDERS) if lk else (fg.MODE_ALL, fg.OUT_PRESENCE)
            r = safe(key, lambda: bench_reads(cx, gi, wl, mode, outp, True, ("lookup -S" if lk else "query -O -S") + ", 150 bp reads, 1% substitutions",
                                              n_reads, steps, warmup, 3000))
            if isinstance(r, dict):
                out[key] = r
                continue
            res, sample = r
            par = safe(key + " parity", lambda: parity_reads(cx, oi, wl, sample, 0 if lk else 1, lk))
            cpu = safe(key + " cpu", lambda: cpu_reads(wl, ["lookup", "-S"] if lk else ["query", "-O", "-S"], 2000 if big else 4000, 9200)) if want_cpu and cx.rank == 0 else None
            finish(key, res, par, cpu)
    # the backward-search tier of the same workload (dict = 0): the kernels the tier-less CLI path and wide indexes run
    if tiers_out is not None:
        def backward():
            gb = fg.Index.load(wl["prefix"], use_klcp=True, device=cx.local_rank, dict=0)
            t = dict(index_hbm_bytes=gb.hbm_bytes, prefix_t=gb.prefix_t, multistep=gb.multistep)
            res = bench_kmers(cx, gb, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O, single k-mers", batch, max(3, steps // 2), warmup, 1000, True,
                              ref_alg_bytes=cx.ref_alg.get(name))
            res["parity_vs_oracle_sample"] = parity_kmers(cx, gb, oi, wl, 100_000, 79, fg.MODE_ALL, fg.OUT_PRESENCE, 1, False)
            res["parity_lookup_vs_oracle_sample"] = parity_kmers(cx, gb, oi, wl, 50_000, 80, fg.MODE_OR, fg.OUT_ORDERS, 0, True)
            res["parity_or_vs_oracle_sample"] = parity_kmers(cx, gb, oi, wl, 50_000, 81, fg.MODE_OR, fg.OUT_PRESENCE, 0, False)
            t["query_O"] = res
            r2, sample = bench_reads(cx, gb, wl, fg.MODE_ALL, fg.OUT_PRESENCE, True, "query -O -S, 150 bp reads, 1% substitutions", n_reads, max(3, steps // 2), warmup, 3000)
            r2["parity_vs_oracle_sample"] = parity_reads(cx, oi, wl, sample, 1, False)
            t["reads_S"] = r2
            gb.close()
            return t
        tiers_out[name] = safe(f"{name}: backward-search tier", backward)
    if oi is not None:
        oi.close()
    wl.pop("codes", None)
    gi.close()
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FMSI_BENCH_WORKLOAD", "human"))
    ap.add_argument("--batch", type=int, default=0, help="k-mers per step per GPU (default: workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="k-mers per reference CPU process (default by workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--modes", default=os.environ.get("FMSI_BENCH_MODES", "auto"),
                    help="auto = all BASELINE configs at N = 1, the streaming-read configs at N > 1; none; or a comma list of workloads")
    ap.add_argument("--cli-records", type=int, default=int(os.environ.get("FMSI_BENCH_CLI_RECORDS", 30_000_000)),
                    help="records of the like-for-like CLI runs (fixed costs - CUDA context 0.6-1.4 s, index load - weigh less the longer the run)")
    ap.add_argument("--prepare-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: W < 3 violates the timing rules; using W = 3")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return 0 if rank != 0 else reference_arm(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    if args.prepare_only:
        return prepare_only(args)
    import fmsi_b200 as fg

    cx = Ctx(args)
    cx.init_dist()
    cx.ref_alg = {}
    world = cx.world
    w = WORKLOADS[args.workload]
    t_start = time.time()

    # ---- headline: the workload's index replica on this GPU, single k-mers, `query -O` ----------------------------
    t0 = time.time()
    if w.get("device_built"):
        gi, wl = prepare_device_built(args.workload, cx.dev, cx.local_rank, with_klcp=True)
        wl.pop("ascii", None)
        torch.cuda.empty_cache()
    else:
        if rank == 0:
            wl = prepare_file_index(args.workload)
        cx.barrier()
        if rank != 0:
            wl = prepare_file_index(args.workload)
        gi = fg.Index.load(wl["prefix"], use_klcp=True, device=cx.local_rank)
    load_s = time.time() - t0
    k = wl["k"]
    batch = args.batch or wl["batch"]
    info = gi.info

    oi, alg = None, None
    if rank == 0 and w.get("device_built") and not (args.no_parity and args.no_cpu_baseline and args.modes == "none"):
        safe("index files", lambda: ensure_files(gi, wl))
    if rank == 0 and not args.no_parity:
        oi = safe("oracle load", lambda: load_oracle(wl, True))
        if isinstance(oi, dict):
            oi = None
        if oi is not None:
            # cpu_baseline leg: the oracle counts the reference algorithm's sector probes on a sample
            alg, _ = reference_algorithm_bytes(oi, wl, host_kmer_sample(wl, 50_000, 76))
            cx.ref_alg[args.workload] = alg["bytes"]

    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
    head = bench_kmers(cx, gi, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O, single k-mers", batch, args.steps, args.warmup, 1000, True,
                       ref_alg_bytes=alg["bytes"] if alg else None)
    clocks = sampler.stop() if rank == 0 else {}
    parity = dict(query_O=parity_kmers(cx, gi, oi, wl, 100_000, 77, fg.MODE_ALL, fg.OUT_PRESENCE, 1, False),
                  query_or=parity_kmers(cx, gi, oi, wl, 50_000, 82, fg.MODE_OR, fg.OUT_PRESENCE, 0, False),
                  lookup=parity_kmers(cx, gi, oi, wl, 50_000, 83, fg.MODE_OR, fg.OUT_ORDERS, 0, True))

    modes, tiers, cli, pool = {}, {}, {}, None
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    cpu = None
    if want_cpu:
        cpu = safe("cpu_baseline", lambda: cpu_kmers(wl, ["query", "-O"], args.cpu_sample or (200_000 if w.get("device_built") else 1_000_000), 9000))

    # ---- the other modes of the headline workload on the same replica -------------------------------------------
    if args.modes != "none":
        name = args.workload

        def reads_mode(lk: bool):
            mode, outp = (fg.MODE_OR, fg.OUT_ORDERS) if lk else (fg.MODE_ALL, fg.OUT_PRESENCE)
            res, sample = bench_reads(cx, gi, wl, mode, outp, True, ("lookup -S" if lk else "query -O -S") + ", 150 bp reads, 1% substitutions",
                                      wl["reads"], args.steps, args.warmup, 3000)
            res["parity_vs_oracle_sample"] = parity_reads(cx, oi, wl, sample, 0 if lk else 1, lk)
            res["cpu_baseline"] = cpu_reads(wl, ["lookup", "-S"] if lk else ["query", "-O", "-S"], 2000, 9200) if want_cpu else None
            res["config"] = dict(workload=name, desc=w["desc"], k=k, n_bwt=gi.n, l2="two alternating read sets, each larger than L2 together with the index")
            return res

        modes[f"{name}_reads_S"] = safe(f"{name}_reads_S", lambda: reads_mode(False))
        if world == 1:
            def lookup_mode():
                res = bench_kmers(cx, gi, wl, fg.MODE_OR, fg.OUT_ORDERS, "lookup, single k-mers", batch, args.steps, args.warmup, 2000, False)
                res["parity_vs_oracle_sample"] = parity["lookup"]
                res["cpu_baseline"] = cpu_kmers(wl, ["lookup"], 100_000 if w.get("device_built") else 300_000, 9100) if want_cpu else None
                res["config"] = dict(workload=name, desc=w["desc"], k=k, n_bwt=gi.n, l2="two alternating input batches, each larger than L2")
                return res
            modes[f"{name}_lookup"] = safe(f"{name}_lookup", lookup_mode)

        # ---- the backward-search tier on the same workload ------------------------------------------------------
        def backward_tier():
            gb = fg.Index.load(wl["prefix"], use_klcp=True, device=cx.local_rank, dict=0)
            t = dict(index_hbm_bytes=gb.hbm_bytes, prefix_t=gb.prefix_t, multistep=gb.multistep)
            res = bench_kmers(cx, gb, wl, fg.MODE_ALL, fg.OUT_PRESENCE, "query -O, single k-mers", batch, max(3, args.steps // 2), args.warmup, 1000, True,
                              ref_alg_bytes=alg["bytes"] if alg else None)
            res["parity_vs_oracle_sample"] = dict(query_O=parity_kmers(cx, gb, oi, wl, 100_000, 79, fg.MODE_ALL, fg.OUT_PRESENCE, 1, False),
                                                  lookup=parity_kmers(cx, gb, oi, wl, 50_000, 80, fg.MODE_OR, fg.OUT_ORDERS, 0, True),
                                                  query_or=parity_kmers(cx, gb, oi, wl, 50_000, 81, fg.MODE_OR, fg.OUT_PRESENCE, 0, False))
            t["query_O"] = res
            r2, sample = bench_reads(cx, gb, wl, fg.MODE_ALL, fg.OUT_PRESENCE, True, "query -O -S, 150 bp reads, 1% substitutions", wl["reads"],
                                     max(3, args.steps // 2), args.warmup, 3000)
            r2["parity_vs_oracle_sample"] = parity_reads(cx, oi, wl, sample, 1, False)
            t["reads_S"] = r2
            gb.close()
            return t
        if world == 1 and os.path.exists(wl["prefix"] + ".fmsi.misc"):
            tiers["backward"] = safe("backward-search tier", backward_tier)

        # ---- like-for-like CLI ---------------------------------------------------------------------------------------
        if want_cpu and args.cli_records > 0:
            cli[name] = safe("cli like-for-like", lambda: cli_like_for_like(wl, args.cli_records, 31337, ["query", "-O"], cx.local_rank))
            cli[name + "_reads_S"] = safe("cli like-for-like, reads -S", lambda: cli_like_for_like(wl, max(1, args.cli_records // 60), 31339, ["query", "-O", "-S"],
                                                                                                     cx.local_rank, reads=True))
            cli[name + "_lookup"] = safe("cli like-for-like, lookup", lambda: cli_like_for_like(wl, max(1, args.cli_records // 3), 31340, ["lookup"], cx.local_rank))

    if oi is not None:
        oi.close()
        oi = None

    # ---- one COMMON batch sharded over the ranks (fmsi_b200/shard.py: contiguous ranges, gather to rank 0 in query order) --
    sharded = None
    if world > 1:
        def sharded_run():
            from fmsi_b200 import shard
            q = host_kmer_sample(wl, 1 << 22, 4242)  # same seed on every rank: the same batch
            times = {}

            def mine(part):
                t0 = time.perf_counter()
                r = gi.query_kmers(part, k, fg.MODE_ALL)
                times["query"] = time.perf_counter() - t0
                return r
            cx.barrier()  # rank 0 comes from checking every rank's parity sample against the oracle
            t0 = time.perf_counter()
            full = shard.sharded_query_kmers(q, mine, rank, world)
            total = cx.max_over_ranks(time.perf_counter() - t0)
            query = cx.max_over_ranks(times["query"])
            if rank != 0:
                return None
            whole = gi.query_kmers(q, k, fg.MODE_ALL)
            return dict(kmers=int(q.size), ranks=world, equals_single_gpu=bool(np.array_equal(full, whole)), shard_query_seconds=round(query, 4),
                        gather_seconds=round(total - query, 4),
                        note="plan_kmers ranges answered by each rank's replica (pageable host buffers), shards gathered to rank 0 in query order "
                             "over the gloo bookkeeping group")
        sharded = safe("sharded common batch", sharded_run)

    # ---- in-process multi-GPU scheduler (fmsi_gpu_pool_*): replicas by device-to-device copy --------------------------
    if world > 1 and args.modes != "none":
        def pool_run():
            gb = fg.Index.load(wl["prefix"], use_klcp=False, device=cx.local_rank, dict=0)
            t0 = time.time()
            pl = fg.Pool(gb, list(range(world)))
            rep_s = time.time() - t0
            n = world * (1 << 24)
            q = torch.from_numpy(host_kmer_sample(wl, n, 555).view(np.int64)).pin_memory().numpy().view(np.uint64)  # pinned, like the e2e arm
            out = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
            pl.query_kmers(q, k, fg.MODE_ALL, out=out)  # warm-up (buffers)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                pl.query_kmers(q, k, fg.MODE_ALL, out=out)
            dt = (time.perf_counter() - t0) / reps
            single = gb.query_kmers(q[:1 << 20], k, fg.MODE_ALL)
            ok = bool(np.array_equal(single, out[:1 << 20]))
            pl.close()
            gb.close()
            return dict(members=world, replicate_s=round(rep_s, 2), index_bytes=int(gb.hbm_bytes), value=n / dt, unit=UNIT, kmers_per_call=n,
                        h2d_gbs_aggregate=round(n * 8 / dt / 1e9, 1), equals_single_index=ok,
                        note="one process, replicas of the backward-search index copied device to device (cudaMemcpyPeer) instead of N index "
                             "builds, pinned host buffers split into contiguous ranges, one host thread per member; 8 B per k-mer over the host links")
        cx.barrier()
        if rank == 0 and os.path.exists(wl["prefix"] + ".fmsi.misc"):
            pool = safe("pool", pool_run)
        cx.barrier()

    wl.pop("codes", None)
    gi.close()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs ---------------------------------------------------------------------------------
    if args.modes != "none":
        if args.modes == "auto":
            plan = [("ecoli", ["query_O", "reads_S", "lookup"]), ("pangenome_k31", ["reads_S"]), ("pangenome_k23", ["reads_S"])] if world == 1 \
                else [("pangenome_k31", ["reads_S"]), ("pangenome_k23", ["reads_S"])]
            if args.workload != "human":
                plan = []
        else:
            plan = [(nm, ["query_O", "reads_S", "lookup"]) for nm in args.modes.split(",") if nm in WORKLOADS and nm != args.workload]
        for nm, which in plan:
            tl = tiers if (world == 1 and nm == "ecoli") else None
            sub = {}
            safe(f"modes: {nm}", lambda: run_workload_modes(cx, nm, which, args.steps, args.warmup, want_cpu, modes, sub if tl is not None else None))
            if tl is not None and sub:
                tiers[f"backward_{nm}"] = sub.get(nm)
        if want_cpu and args.cli_records > 0 and args.workload == "human":
            def cli_ecoli():
                wle = prepare_file_index("ecoli")
                return cli_like_for_like(wle, args.cli_records, 31338, ["query", "-O"], cx.local_rank)
            cli["ecoli"] = safe("cli like-for-like (ecoli)", cli_ecoli)

    if rank != 0:
        if cx.dist:
            cx.dist.barrier()
            cx.dist.destroy_process_group()
        return 0

    cfg = static_config(args.workload)
    details = dict(kmers_per_step_per_gpu=batch, parallelism=f"replicas x{world}, queries sharded, no collective",
                   l2="inputs larger than L2: 2 alternating batches of %d MiB" % (batch * 8 >> 20), n_bwt=int(info.n_bwt), prefix_t=int(info.prefix_t),
                   dictionary_tier=int(info.dict), dictionary_depth=int(info.dict_t), multistep=int(info.multistep), index_hbm_bytes=int(info.hbm_bytes),
                   index_setup_s=round(load_s, 2), frac_present=head["frac_present"], e2e_equals_device=head["e2e_equals_device"],
                   parity_vs_oracle_sample=parity, parity_ranks=world,
                   reference_algorithm=dict(bytes_per_kmer=round(alg["bytes"], 1), lf_steps_per_kmer=round(alg["lf_steps"], 2),
                                            rank_sectors_per_kmer=round(alg["rank_sectors"], 2), mask_sectors_per_kmer=round(alg["mask_sectors"], 2)) if alg else None,
                   bench_wall_s=round(time.time() - t_start, 1))
    line = dict(metric=METRIC, value=head["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=head["ms_per_step"],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64", data="synthetic", config=cfg, details=details,
                roofline=head["roofline"], cpu_baseline=cpu, e2e=head["e2e"], gpu_launches=head["gpu_launches"], clocks=clocks,
                tiers=tiers, modes=modes, cli=cli, pool=pool, sharded=sharded)
    print(json.dumps(line))
    if cx.dist:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
