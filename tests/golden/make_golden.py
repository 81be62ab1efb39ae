#!/usr/bin/env python3
"""Generate tests/golden/ fixtures with the UNMODIFIED reference binary (oracle/_ref/fmsi, built
from /root/reference by oracle/Makefile). Run in the authoring container only; the outputs are
committed so the GPU box (which has no /root/reference) can test against them.

Each case directory holds
    ms.fa                     masked superstring (input of `fmsi index`)
    ms.fa.fmsi.*              index files written by the reference's `fmsi index`
    q.fa                      query file
    exp_<cmd>.txt             stdout of the reference: query, query_O, query_S, query_OS, lookup, lookup_S
    meta.json                 k, flags, provenance

Cases cover the reference's own goldens (tests/testfiles/*, the hand-built unit-test fixtures
re-created as real index files, tests/data/* superstrings) plus seeded synthetic superstrings with
max-ones and min-ones masks, both-strand occurrences, invalid characters, lower case, FASTQ,
multi-line FASTA, CRLF, empty and short records.
"""
from __future__ import annotations

import json
import os
import random
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fmsi_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "fmsi")
REFSRC = os.environ.get("FMSI_REFERENCE", "/root/reference")

CMDS = {
    "query": ["query"],
    "query_O": ["query", "-O"],
    "query_S": ["query", "-S"],
    "query_OS": ["query", "-O", "-S"],
    "lookup": ["lookup"],
    "lookup_S": ["lookup", "-S"],
}


def run_ref(args, cwd):
    r = subprocess.run([REF] + args, cwd=cwd, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed: {args}\n{r.stderr.decode()}")
    return r.stdout


ONLY = set(sys.argv[1:])  # `make_golden.py CASE...` regenerates just those cases


def make_case(name: str, ms: bytes, k: int, queries: bytes, klcp: bool = True, note: str = "", header: bytes = b"ms"):
    d = os.path.join(HERE, name)
    if ONLY and name not in ONLY:
        return
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    with open(os.path.join(d, "ms.fa"), "wb") as f:
        f.write(b">" + header + b"\n" + ms + b"\n")
    with open(os.path.join(d, "q.fa"), "wb") as f:
        f.write(queries)
    run_ref(["index", "-k", str(k)] + ([] if klcp else ["-x"]) + ["ms.fa"], d)
    cmds = []
    for tag, args in CMDS.items():
        if "-S" in args and not klcp:
            continue
        out = run_ref(args + ["-q", "q.fa", "ms.fa"], d)
        with open(os.path.join(d, f"exp_{tag}.txt"), "wb") as f:
            f.write(out)
        cmds.append(tag)
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump({"k": k, "klcp": klcp, "cmds": cmds, "note": note, "reference": "OndrejSladky/fmsi v0.4.0 (39c71a1)"}, f, indent=1)
    print(f"{name}: ms={len(ms)} k={k} queries={len(queries)}B cmds={cmds}")


def fasta(records) -> bytes:
    return b"".join(b">" + n + b"\n" + s + b"\n" for n, s in records)


def mixed_queries(g: np.ndarray, k: int, n_kmers: int, n_reads: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    q = bytearray(synth.rows_to_fasta(synth.kmer_queries(g, k, n_kmers, seed + 1)))
    reads = synth.read_queries(g, 150, n_reads, seed + 2, 0.02)
    for i, rd in enumerate(reads):
        s = bytearray(synth.ACGT[rd].tobytes())
        if i % 7 == 0:
            s[int(rng.integers(0, len(s)))] = ord("N")
        if i % 11 == 0:
            s[-2] = ord("n")
        if i % 13 == 0:
            s = bytearray(bytes(s).lower())
        if i % 17 == 0:
            s[3:3] = b"RY"
        if i % 5 == 0:
            q += b"@r%d some comment\n" % i + bytes(s) + b"\n+\n" + b"I" * len(s) + b"\n"
        else:
            q += b">r%d\n" % i + bytes(s[:70]) + b"\n" + bytes(s[70:]) + b"\n"
    # one long record (chunking: max chunk = k + clamp(2*floor(sqrt(len)), 10, 400))
    long_codes = np.concatenate([g[:3000], synth.revcomp_codes(g[1000:2500]), g[5000:7000]])
    q += b">long\n" + synth.codes_to_ascii(long_codes) + b"\n"
    q += b">short\nACG\n>empty\n\n>\nACGTNNACGT\n>tail desc\tmore\nACGTACGTAC\n"
    return bytes(q)


def main():
    if not os.path.exists(REF):
        sys.exit(f"{REF} missing: run `make -C oracle` first (needs /root/reference)")
    tf = os.path.join(REFSRC, "tests", "testfiles")
    # --- the reference's CLI goldens (tests/integration_test.sh) --------------------------------
    queries = open(os.path.join(tf, "queries.txt"), "rb").read()
    for tag in ("a", "b"):
        lines = open(os.path.join(tf, f"integration_{tag}.fa"), "rb").read().split(b"\n")
        make_case(f"integration_{tag}", lines[1].strip(), 3, queries, note=f"tests/testfiles/integration_{tag}.fa + queries.txt",
                  header=lines[0][1:].strip())
        # the reference's own committed expectations must equal what its binary prints here
        want = open(os.path.join(tf, f"result_{tag}_complements.txt"), "rb").read()
        got = open(os.path.join(HERE, f"integration_{tag}", "exp_query.txt"), "rb").read()
        assert want == got, f"integration_{tag}: reference binary disagrees with its own golden"
        shutil.copy(os.path.join(tf, f"result_{tag}_complements.txt"), os.path.join(HERE, f"integration_{tag}", "ref_golden_query.txt"))
    want = open(os.path.join(tf, "result_a_complements_hash.txt"), "rb").read()
    assert want == open(os.path.join(HERE, "integration_a", "exp_lookup.txt"), "rb").read()
    shutil.copy(os.path.join(tf, "result_a_complements_hash.txt"), os.path.join(HERE, "integration_a", "ref_golden_lookup.txt"))

    # --- unit-test fixtures as real indexes (tests/fms_index_test.h:10-69) -----------------------
    make_case("fixture1_CaGGTag_k2", b"CaGGTag", 2,
              fasta([(b"AG", b"AG"), (b"CA", b"CA"), (b"AC", b"AC"), (b"TA", b"TA"), (b"GA", b"GA"), (b"CC", b"CC"), (b"CCAG", b"CCAG"), (b"all", b"CAGGTAGCTACCTG")]),
              note="get_dummy_index(): QUERY_ORDERS / QUERY cases with k=2")
    make_case("fixture3_CACaCat_k3", b"CACaCat", 3,
              fasta([(b"s1", b"CACATACA"), (b"s2", b"TGTATGTG"), (b"s3", b"CACATTGT"), (b"s4", b"CACATACA")]),
              note="get_dummy_index3(): QUERY_KMERS_STREAMING(_ORDERS) cases")

    # --- tests/data superstrings used by verify.py ----------------------------------------------
    random.seed(42)
    for fn, k in (("GCF_009858895.2_ASM985889v3_genomic.fna.ms.k13", 13), ("GCF_test2_ones.fna.ms.k31", 31)):
        lines = open(os.path.join(REFSRC, "tests", "data", fn), "rb").read().split(b"\n")
        ms = b"".join(l.strip() for l in lines[1:])
        up = ms.upper()
        recs = []
        for i in range(60):  # verify.py: positive windows of consecutive k-mers + random sequences
            L = 50 + k - 1
            p = random.randrange(0, len(up) - L)
            recs.append((b"pos%d" % i, up[p:p + L]))
            recs.append((b"rnd%d" % i, bytes(random.choice(b"ACGT") for _ in range(L))))
        make_case(f"data_k{k}", ms, k, fasta(recs), note=f"tests/data/{fn}")

    # --- seeded synthetic ---------------------------------------------------------------------
    g = synth.random_codes(60000, 11)
    make_case("syn_k31_max", synth.contig_superstring(g, 31, 40, 12, "max"), 31, mixed_queries(g, 31, 400, 40, 13),
              note="60 kbp random genome, 40 shuffled/RC contigs, max-ones mask")
    g = synth.random_codes(60000, 21)
    make_case("syn_k31_min", synth.contig_superstring(g, 31, 40, 22, "min"), 31, mixed_queries(g, 31, 400, 40, 23),
              note="min-ones mask (one ON occurrence per canonical k-mer)")
    g = synth.random_codes(20000, 31)
    make_case("syn_k9_max", synth.contig_superstring(g, 9, 30, 32, "max"), 9, mixed_queries(g, 9, 400, 40, 33),
              note="k=9: k-mers occur on both strands, strand predictor changes lookup output")
    g = synth.random_codes(20000, 41)
    make_case("syn_k9_min", synth.contig_superstring(g, 9, 30, 42, "min"), 9, mixed_queries(g, 9, 400, 40, 43),
              note="k=9 min-ones: -O output depends on the strand predictor")
    g = synth.random_codes(3000, 51)
    make_case("syn_k5_min", synth.contig_superstring(g, 5, 10, 52, "min"), 5, mixed_queries(g, 5, 300, 30, 53),
              note="k=5: dense both-strand / OFF occurrences")
    g = synth.random_codes(40000, 61)
    make_case("syn_k31_noklcp", synth.genome_superstring(g, 31), 31, mixed_queries(g, 31, 300, 20, 63), klcp=False,
              note="index built with -x (no kLCP)")
    g = synth.random_codes(5000, 71)
    make_case("syn_k32", synth.genome_superstring(g, 32), 32, mixed_queries(g, 32, 200, 20, 73), note="k=32 (widest packed k-mer)")

    # --- k > 32: the k-mer no longer fits one packed word (get_range_with_pattern takes any k) -------
    g = synth.random_codes(20000, 81)
    make_case("syn_k47_max", synth.contig_superstring(g, 47, 12, 82, "max"), 47, mixed_queries(g, 47, 300, 40, 83),
              note="k=47 (two pattern windows per strand), max-ones mask")
    g = synth.random_codes(20000, 91)
    half = synth.random_codes(32, 92)
    pal = np.concatenate([half, synth.revcomp_codes(half)])  # a self-complementary 64-mer (general mode counts it once)
    g[4000:4064] = pal
    g[9000:9064] = pal
    make_case("syn_k64_min", synth.contig_superstring(g, 64, 12, 93, "min"), 64,
              mixed_queries(g, 64, 300, 40, 94) + b">pal\n" + synth.codes_to_ascii(pal) + b"\n>palctx\n" + synth.codes_to_ascii(g[3990:4080]) + b"\n",
              note="k=64 (largest k with kLCP), min-ones mask, a self-complementary 64-mer occurring twice")
    g = synth.random_codes(20000, 101)
    make_case("syn_k97_noklcp", synth.contig_superstring(g, 97, 8, 102, "max"), 97, mixed_queries(g, 97, 300, 40, 103), klcp=False,
              note="k=97 > 64: `fmsi index` builds no kLCP array (main.cpp:225-228); four pattern windows per strand")

    # --- parser / record-loop quirks ----------------------------------------
    quirks = (b">q1\nACGTN\n>q2\nNNACG\n>q3\nACGNTAC\nGTA\n>q4 lower\nacgtacgt\n@fq1 c\nACGTACG\n+\nIIIIIII\n"
              b">crlf\r\nACGTAC\r\nGTAC\r\n>\nACGT\n>onlyN\nNNNN\n>x\nAC\n>y\n\n>mixed\nACGTXACGTACGT-ACG\n;comment line\n>z\nTTTT\n"
              b">ambig1\nAAC\n>a\nTGA\n>b\nTGA\n>c\nTGA\n>d\nTGA\n>ambig2\nAAC\n")
    make_case("quirks_k3", b"AACGTTCAtt", 3, quirks, note="probe cases of SURVEY §8a rows P and 10")
    make_case("quirks_k3_nonmax", b"AACgTTCAtt", 3, quirks, note="non-max-ones mask: -O depends on the predictor")


if __name__ == "__main__":
    main()
