#!/usr/bin/env python3
"""bench.py — 31-mer queries/s of the FMS-index query path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path (`fmsi query -O` semantics: fmsi_gpu_query_kmers, mode ALL,
strands LAZY) over one batch of synthetic packed 31-mers, 50 % present. `value` is measured with
the batch resident in HBM (CUDA events on the launching stream); `e2e` is the same metric through
the C-ABI with pinned HOST buffers, host<->device copies inside the timed region. With N > 1
(torchrun) every rank owns one GPU, a full replica of the index and its own batch (weak scaling, no
data-path collective); the time is the max over ranks.

`--impl reference` times the unmodified reference CPU binary (oracle/_ref/fmsi, one process per
host core over FASTA shards of a bounded sample) on the same index and query distribution.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from fmsi_b200 import synth  # noqa: E402

DATA = os.path.join(ROOT, "data")
REF_FMSI = os.path.join(ROOT, "oracle", "_ref", "fmsi")
REF_KMERCAMEL = os.path.join(ROOT, "oracle", "_ref", "kmercamel")
METRIC = "31-mer queries/sec"
UNIT = "kmers/s"

WORKLOADS = {
    # BASELINE.json configs[0]: E. coli-sized synthetic, k=31, masked superstring via bundled
    # kmercamel + `fmsi index`; single 31-mer `fmsi query -O`, 50 % present.
    "ecoli": dict(genome_len=5_000_000, k=31, seed=1, superstring="kmercamel", batch=1 << 26,
                  desc="5 Mbp random genome, kmercamel -c + optimize -a ones, fmsi index -k 31; single 31-mers, 50% present"),
    # BASELINE.json configs[3]: human-scale synthetic: 3.1 Gbp i.i.d. sequence (itself a valid max-ones
    # masked superstring: upper case except the last k-1 letters), k=31, index replicated per GPU.
    # The reference's `fmsi index` needs hours and ~53 GB for this, so the index is built on the GPU
    # (fmsi_gpu_index_build, byte-identical files, tests/test_gpu_build.py) and saved for the CPU arm.
    "human": dict(genome_len=3_100_000_000, k=31, seed=4, superstring="genome", batch=1 << 26, device_built=True,
                  desc="3.1 Gbp i.i.d. sequence as max-ones masked superstring, k=31, index built on GPU; single 31-mers, 50% present"),
    "human_small": dict(genome_len=400_000_000, k=31, seed=4, superstring="genome", batch=1 << 26, device_built=True,
                        desc="400 Mbp i.i.d. sequence (reduced human-scale shape, index >> L2), k=31"),
    # small variant for quick checks
    "tiny": dict(genome_len=200_000, k=31, seed=3, superstring="contigs", batch=1 << 22,
                 desc="200 kbp random genome (debug)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {r.stderr.decode(errors='replace')[-2000:]}")
    return r


# ------------------------------------------------------------------------------------------------
def prepare_index(name: str) -> dict:
    """Genome -> masked superstring -> reference `fmsi index`; cached under data/<name>/."""
    w = WORKLOADS[name]
    d = os.path.join(DATA, name)
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "ms.fa")
    k = w["k"]
    genome = synth.random_codes(w["genome_len"], w["seed"])
    if not os.path.exists(prefix + ".fmsi.misc"):
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
            how = "kmercamel"
        else:
            synth.write_fasta_single(prefix, "ms", synth.contig_superstring(genome, k, 64, w["seed"] + 100, "max"))
            how = "contigs"
        run([REF_FMSI, "index", "-k", str(k), prefix])
        with open(os.path.join(d, "how.txt"), "w") as f:
            f.write(how)
        log(f"[bench] built index {name} ({how}) in {time.time() - t0:.1f}s")
    how = open(os.path.join(d, "how.txt")).read().strip() if os.path.exists(os.path.join(d, "how.txt")) else "?"
    return dict(prefix=prefix, k=k, genome=genome, how=how, **{kk: w[kk] for kk in ("batch", "desc", "genome_len")})


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


def device_queries(codes, k: int, batch: int, seed: int, dev, frac_present: float = 0.5):
    """Packed k-mers on the device: present ones from uniform genome positions on a random strand,
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


def prepare_device_built(name: str, dev, local_rank: int, save_files: bool):
    """Human-scale shapes: genome and index built on the GPU; files saved once for the CPU arm."""
    import torch
    import fmsi_b200 as fg
    w = WORKLOADS[name]
    n, k = w["genome_len"], w["k"]
    d = os.path.join(DATA, name)
    prefix = os.path.join(d, "ms.fa")
    t0 = time.time()
    codes, ascii_ = device_genome(n, w["seed"], k, dev)
    torch.cuda.synchronize(dev)
    t1 = time.time()
    gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=False, device=local_rank, n=n, mem=fg.MEM_DEVICE)
    del ascii_
    torch.cuda.empty_cache()
    t2 = time.time()
    saved = None
    if save_files and not os.path.exists(prefix + ".fmsi.misc"):
        os.makedirs(d, exist_ok=True)
        gi.save(prefix)
        saved = time.time() - t2
    log(f"[bench] {name}: genome {t1 - t0:.1f}s, GPU index build {t2 - t1:.1f}s" + (f", save {saved:.1f}s" if saved else ""))
    wl = dict(prefix=prefix, k=k, genome=None, codes=codes, how="genome (GPU-built index)", batch=w["batch"], desc=w["desc"],
              genome_len=n, build_s=round(t2 - t1, 2))
    return gi, wl


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
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def sample_queries(wl: dict, n: int, seed: int) -> np.ndarray:
    """n packed query k-mers of the workload's distribution, in host memory."""
    if wl.get("codes") is not None:
        import torch
        q = device_queries(wl["codes"], wl["k"], n, seed, wl["codes"].device)
        return q.cpu().numpy().view(np.uint64)
    if wl.get("_gk") is None:
        wl["_gk"] = synth.pack_kmers(wl["genome"], wl["k"])
    return synth.packed_kmer_queries(wl["_gk"], wl["k"], n, seed)


def reference_cpu_rate(wl: dict, per_proc: int, seed: int, procs: int | None = None) -> dict:
    """The reference's own CPU query path: P independent `fmsi query -O` processes over FASTA shards
    (the reference has no threads). Wall time from first start to last exit; the per-process index
    load is measured separately with a one-record query file and reported."""
    if not os.path.exists(REF_FMSI):
        raise RuntimeError("oracle/_ref/fmsi missing")
    P = procs or os.cpu_count() or 1
    k = wl["k"]
    tmp = tempfile.mkdtemp(prefix="fmsi_ref_")
    try:
        files = []
        allq = sample_queries(wl, P * per_proc, seed)
        for p in range(P):
            fn = os.path.join(tmp, f"q{p}.fa")
            with open(fn, "wb") as f:
                f.write(synth.packed_to_fasta(allq[p * per_proc:(p + 1) * per_proc], k))
            files.append(fn)
        one = os.path.join(tmp, "one.fa")
        with open(one, "wb") as f:
            f.write(b">q\n" + b"A" * k + b"\n")
        t0 = time.perf_counter()
        run([REF_FMSI, "query", "-O", "-q", one, wl["prefix"]])
        load_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        ps = [subprocess.Popen([REF_FMSI, "query", "-O", "-q", fn, wl["prefix"]], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for fn in files]
        for p_ in ps:
            if p_.wait() != 0:
                raise RuntimeError("reference fmsi query failed")
        wall = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    total = P * per_proc
    query_wall = max(wall - load_s, 1e-9)
    # `value` is the steady-state query rate: the per-process index load is subtracted, as the GPU arm's
    # numbers exclude its index set-up too; the rate over the whole wall time is reported beside it.
    return dict(value=total / query_wall, unit=UNIT, cores=P, kind="reference", wall_s=round(wall, 3), index_load_s=round(load_s, 3),
                query_wall_s=round(query_wall, 3), value_including_load=total / wall,
                sample=f"{P} concurrent `fmsi query -O` processes x {per_proc} single 31-mer FASTA records (50% present), "
                       f"same index; wall {wall:.2f}s of which per-process index load {load_s:.2f}s (subtracted)")


def algorithmic_bytes_per_kmer(wl: dict, sample: int, seed: int) -> dict:
    """32 B x (rank sectors + mask sectors) per k-mer, counted by the instrumented
    oracle on a sample of the same query distribution (forward strand first, neutral predictor),
    + 8 B packed query in + 1 B result out. Part of the cpu_baseline leg: the oracle is only a counter here."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_ffi import MODE_ALL, OracleIndex
    oi = OracleIndex.load(wl["prefix"], use_klcp=False)
    q = sample_queries(wl, sample, seed)
    oi.counters_reset()
    want = oi.query_packed(q, wl["k"], MODE_ALL, False)
    c = oi.counters()
    oi.close()
    n = c["kmers"]
    per = dict(lf_steps=c["lf_steps"] / n, rank_sectors=c["rank_sectors"] / n, mask_sectors=c["mask_sectors"] / n)
    per["bytes"] = 32.0 * (per["rank_sectors"] + per["mask_sectors"]) + 8 + 1
    per["_queries"] = q
    per["_expected"] = want
    return per


# ------------------------------------------------------------------------------------------------
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
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: W < 3 violates the timing rules; using W = 3")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    device_built = bool(w.get("device_built"))
    cpu_sample = args.cpu_sample or (200_000 if device_built else 1_000_000)

    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import fmsi_b200 as fg

    if not torch.cuda.is_available():
        if args.impl == "reference" and not device_built:
            dev = None
        else:
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    else:
        dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(dev)

    if args.impl == "reference":
        # The reference's CPU path on the box's host cores. For the human-scale shapes the index files
        # come from the GPU builder (byte-identical to `fmsi index`, which would take hours here).
        if device_built:
            gi, wl = prepare_device_built(args.workload, dev, local_rank, save_files=True)
            gi.close()
        else:
            wl = prepare_index(args.workload)
        vals = []
        for s_ in range(args.warmup + args.steps):
            r = reference_cpu_rate(wl, cpu_sample, seed=5000 + 100 * s_)
            if s_ >= args.warmup:
                vals.append(r)
        wall = sum(v["query_wall_s"] for v in vals)
        total = sum(v["cores"] * cpu_sample for v in vals)
        value = total / wall
        cb = dict(vals[-1])
        cb["value"] = value
        line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1000.0 * wall / len(vals), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64",
                    data="synthetic", config=dict(workload=args.workload, desc=wl["desc"], k=wl["k"], superstring=wl["how"],
                                                  kmers_per_step=cb["cores"] * cpu_sample, mode="query -O"),
                    cpu_baseline=cb, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return 0

    if world > 1:
        import torch.distributed as dist
        # Bookkeeping only (barriers, max of the elapsed time): the query path has no collective, so the group runs
        # over gloo. Merely initialising NCCL (no collective in flight) slowed the random-probe kernel
        # by 5.5 % on every rank (1.6875 vs 1.600 ms per step at N = 2, profiles/r01v_* vs r01w_*);
        # FMSI_BENCH_DIST_BACKEND=nccl selects it anyway.
        backend = os.environ.get("FMSI_BENCH_DIST_BACKEND", "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            # one node: gloo over the loopback interface (its default picks the interface the host name resolves to,
            # and a container's host name may not resolve)
            if backend == "gloo" and os.environ.get("MASTER_ADDR", "127.0.0.1") in ("127.0.0.1", "localhost", "::1"):
                os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            try:
                dist.init_process_group(backend)
            except Exception as ex:  # e.g. a host name that does not resolve: loopback, then NCCL (same on every rank of the box)
                log(f"[bench] {backend} process group failed ({ex}); retrying over the loopback interface")
                os.environ["GLOO_SOCKET_IFNAME"] = "lo"
                try:
                    dist.init_process_group(backend)
                except Exception as ex2:
                    log(f"[bench] {backend} failed again ({ex2}); falling back to nccl")
                    dist.init_process_group("nccl", device_id=dev)

    # ---- workload: index replica on this GPU ----------------------------------------------------
    t0 = time.time()
    if device_built:
        gi, wl = prepare_device_built(args.workload, dev, local_rank, save_files=(rank == 0 and world == 1 and not args.no_cpu_baseline))
    else:
        if world > 1:  # rank 0 prepares the cached index first so that ranks do not race on the files
            if rank == 0:
                wl = prepare_index(args.workload)
            dist.barrier()
            if rank != 0:
                wl = prepare_index(args.workload)
        else:
            wl = prepare_index(args.workload)
        gi = fg.Index.load(wl["prefix"], use_klcp=False, device=local_rank)
    load_s = time.time() - t0
    k = wl["k"]
    batch = args.batch or wl["batch"]

    nbuf = 2  # alternate between distinct batches; each is larger than L2
    if device_built:
        d_in = [device_queries(wl["codes"], k, batch, 1000 + 17 * rank + b, dev) for b in range(nbuf)]
        pinned_in = [t.cpu().pin_memory() for t in d_in]
    else:
        gk = synth.pack_kmers(wl["genome"], k)
        wl["_gk"] = gk
        host_batches = [synth.packed_kmer_queries(gk, k, batch, seed=1000 + 17 * rank + b) for b in range(nbuf)]
        pinned_in = [torch.from_numpy(h.view(np.int64)).pin_memory() for h in host_batches]
        d_in = [p.to(dev, non_blocking=False) for p in pinned_in]
    pinned_out = torch.empty(batch, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(batch, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device(s_):
        gi.query_kmers_ptr(d_in[s_ % nbuf].data_ptr(), batch, d_out.data_ptr(), k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY,
                           fg.MEM_DEVICE, stream.cuda_stream)

    def step_host(s_):
        gi.query_kmers_ptr(pinned_in[s_ % nbuf].data_ptr(), batch, pinned_out.data_ptr(), k, fg.MODE_ALL, fg.OUT_PRESENCE,
                           fg.STRANDS_LAZY, fg.MEM_HOST, 0)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-resident timing -----------------------------------------------------------------
    for s_ in range(args.warmup):
        step_device(s_)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = fg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s_ in range(args.steps):
        step_device(s_)
    e1.record(stream)
    sync_all()
    launches = fg.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * batch / (ms_per_step / 1e3)
    frac_present = float(d_out.float().mean().item())  # the timed kernel did the work (~50 % present)

    # ---- end-to-end through the C-ABI with host buffers --------------------------------------
    for s_ in range(2):
        step_host(s_)
    sync_all()
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        step_host(s_)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else {}
    e2e_value = world * batch * args.steps / e2e_s
    step_device(args.steps - 1)
    torch.cuda.synchronize(dev)
    same = bool(torch.equal(d_out.cpu(), pinned_out))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline + CPU baseline (rank 0) -----------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    roofline, cpu, parity = None, None, None
    have_files = os.path.exists(wl["prefix"] + ".fmsi.misc")
    alg = None
    if have_files and not args.no_cpu_baseline and world == 1:
        # cpu_baseline leg: the oracle counts the reference algorithm's sector probes on a sample and
        # checks the GPU's answers on it; the reference binary is timed on the host cores.
        alg = algorithmic_bytes_per_kmer(wl, 100_000 if device_built else 200_000, seed=77)
        got = gi.query_kmers(alg["_queries"], k, fg.MODE_ALL)
        parity = bool(np.array_equal(got.astype(np.int64), alg["_expected"]))
        with open(os.path.join(ROOT, "profiles", f"algorithmic_{args.workload}.json"), "w") as f:
            json.dump({kk: vv for kk, vv in alg.items() if not kk.startswith("_")}, f)
        try:
            cpu = reference_cpu_rate(wl, cpu_sample, seed=9000)
        except Exception as ex:  # keep the bench line even if the reference binary did not travel
            cpu = dict(value=None, unit=UNIT, cores=0, kind="reference", sample=f"unavailable: {ex}")
    else:
        apath = os.path.join(ROOT, "profiles", f"algorithmic_{args.workload}.json")
        if os.path.exists(apath):
            alg = json.load(open(apath))
    if alg:
        # one fold_query_kernel (dict_query_kernel / query_kmers_kernel for the other tiers) launch per step
        # is the step; the 32-byte cursor memset in front of it is negligible
        launch_ms = ms_per_step
        achieved = alg["bytes"] * batch / (launch_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(args.workload)
            if tj and tj.get("kernel", "").split("<")[0] != {2: "fold_query_kernel", 1: "dict_query_kernel", 0: "query_kmers_kernel"}[gi.dict_kind]:
                tj = None  # the stored ncu capture is of another tier's kernel
            if tj and tj.get("batch"):
                traffic = tj["dram_bytes_per_launch"] * (batch / tj["batch"])
        kernel_name = {2: "fold_query_kernel<ALL,PRESENCE,LAZY>", 1: "dict_query_kernel<ALL,PRESENCE,LAZY>",
                       0: "query_kmers_kernel<ALL,PRESENCE,LAZY>"}[gi.dict_kind]
        hw = None
        if traffic:
            # what the hardware actually did (ncu, profiles/traffic.json): DRAM bytes moved per second
            # against the same measured peak, and L2-miss requests per second against the measured
            # random-request ceiling of this part (profiles/r01b_randbw2_b200.jsonl)
            hw = dict(dram_gbs=round(traffic / (launch_ms / 1e3) / 1e9, 1), dram_frac=round(traffic / (launch_ms / 1e3) / 1e9 / peak, 4))
            if tj.get("l1_sectors_per_launch"):
                req = tj["l1_sectors_per_launch"] * (batch / tj["batch"]) / (launch_ms / 1e3) / 1e9
                hw.update(requests_per_kmer=round(tj["l1_sectors_per_launch"] / tj["batch"], 3), grequests_s=round(req, 1),
                          random_request_ceiling_grequests_s=tj.get("random_request_ceiling_grequests_s"))
                if tj.get("query_stream_sectors_per_launch"):
                    # requested sectors minus the coalesced read of the queries themselves = dependent random probes
                    # (bucket / rows sectors); the ceiling is what tools/randbw2 measured for pure L2 misses on this part
                    probes = (tj["l1_sectors_per_launch"] - tj["query_stream_sectors_per_launch"]) / tj["batch"]
                    hw.update(random_probes_per_kmer=round(probes, 3), grandom_probes_s=round(probes * batch / (launch_ms / 1e3) / 1e9, 1),
                              l2_hit_rate=tj.get("l2_hit_rate"))
        roofline = dict(bound="hbm", achieved=round(achieved, 1), peak=peak, unit="GB/s", frac=round(achieved / peak, 4), traffic=traffic,
                        kernel=kernel_name, launch_ms=round(launch_ms, 4), peak_source=peak_src, hardware=hw,
                        algorithmic=dict(bytes_per_kmer=round(alg["bytes"], 1), lf_steps_per_kmer=round(alg["lf_steps"], 2),
                                         rank_sectors_per_kmer=round(alg["rank_sectors"], 2), mask_sectors_per_kmer=round(alg["mask_sectors"], 2)),
                        note="algorithmic bytes = the REFERENCE algorithm's dependent sector probes per k-mer (SURVEY 8d: 32 B per "
                             "rank/mask probe + 8 B query + 1 B result), counted by the instrumented oracle on the same query "
                             "distribution; frac > 1 because the strand-folded dictionary answers a k-mer (both strands) in ~1 request "
                             "instead of 2 x (k-t) LF-steps - `hardware` says how close the kernel runs to the memory system's own limits")

    info = gi.info
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64", data="synthetic",
                config=dict(workload=args.workload, desc=wl["desc"], k=k, superstring=wl["how"], kmers_per_step_per_gpu=batch,
                            mode="query -O (MODE_ALL, STRANDS_LAZY)", parallelism=f"replicas x{world}, queries sharded, no collective",
                            l2="inputs larger than L2: 2 alternating batches of %d MiB" % (batch * 8 >> 20), n_bwt=int(info.n_bwt),
                            prefix_t=int(info.prefix_t), dictionary_tier=int(info.dict), dictionary_depth=int(info.dict_t), index_hbm_bytes=int(info.hbm_bytes), index_setup_s=round(load_s, 2),
                            frac_present=round(frac_present, 4), e2e_equals_device=same, parity_vs_oracle_sample=parity),
                roofline=roofline, cpu_baseline=cpu,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=batch * 8, d2h_bytes_per_step=batch * 1),
                gpu_launches=int(launches), clocks=clocks)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
