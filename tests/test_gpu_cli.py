"""The `fmsi` front end (fmsi_b200/bin/fmsi) must print exactly what the reference prints:
byte-for-byte comparison with the reference binary's committed outputs under tests/golden/,
for every command/flag combination, including the cases whose output depends on the reference's
stateful strand predictor."""
import gzip
import json
from concurrent.futures import ThreadPoolExecutor
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, golden_cases

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "fmsi_b200", "bin", "fmsi")
ARGS = {"query": ["query"], "query_O": ["query", "-O"], "query_S": ["query", "-S"], "query_OS": ["query", "-O", "-S"],
        "lookup": ["lookup"], "lookup_S": ["lookup", "-S"]}
# cases where no k-mer is decided differently by strand order (max-ones masks, k=31 random genomes)
LAZY_EXACT = {"syn_k31_max": list(ARGS), "data_k31": list(ARGS), "syn_k31_noklcp": ["query", "query_O", "lookup"],
              "syn_k31_min": ["query", "query_S"], "syn_k9_min": ["query", "query_S"], "syn_k5_min": ["query", "query_S"]}


def run_cli(args, env=None, stdin=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([CLI] + args, capture_output=True, env=e, input=stdin)


def run_many(arg_lists, env=None):
    """Several CLI processes at once: creating a CUDA context costs 1-3 s per process on the box."""
    with ThreadPoolExecutor(max_workers=6) as ex:
        return list(ex.map(lambda a: run_cli(a, env=env), arg_lists))


@pytest.mark.parametrize("case", golden_cases())
def test_cli_matches_reference_outputs(case):
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    runs = run_many([ARGS[cmd] + ["-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")] for cmd in meta["cmds"]])
    for cmd, r in zip(meta["cmds"], runs):
        assert r.returncode == 0, r.stderr.decode()
        want = open(os.path.join(d, f"exp_{cmd}.txt"), "rb").read()
        assert r.stdout == want, f"{case}/{cmd}"


F_ARGS = {"query_f_xor": ["query", "-f", "xor"], "query_f_and": ["query", "-f", "and"], "query_f_1-1": ["query", "-f", "1-1"],
          "query_f_2-9": ["query", "-f", "2-9"], "query_S_f_xor": ["query", "-S", "-f", "xor"]}


@pytest.mark.parametrize("case", golden_cases())
def test_cli_general_demasking_functions(case):
    """`query -f xor|and|INT-INT` (f-MS framework, reference fms_index.h:317-327 + functions.h) against
    the reference binary's outputs, incl. the reference's own golden result_b_complements_xor.txt."""
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    cmds = meta.get("cmds_f", [])
    runs = run_many([F_ARGS[cmd] + ["-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")] for cmd in cmds])
    for cmd, r in zip(cmds, runs):
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == open(os.path.join(d, f"exp_{cmd}.txt"), "rb").read(), f"{case}/{cmd}"
    if case == "integration_b":
        assert runs[cmds.index("query_f_xor")].stdout == open(os.path.join(d, "ref_golden_query_f_xor.txt"), "rb").read()


@pytest.mark.parametrize("case", sorted(LAZY_EXACT))
def test_cli_lazy_mode_where_order_independent(case):
    d = os.path.join(GOLDEN, case)
    runs = run_many([ARGS[cmd] + ["-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")] for cmd in LAZY_EXACT[case]],
                    env={"FMSI_GPU_STRANDS": "lazy"})
    for cmd, r in zip(LAZY_EXACT[case], runs):
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == open(os.path.join(d, f"exp_{cmd}.txt"), "rb").read(), f"{case}/{cmd}"


@pytest.mark.parametrize("case", ["quirks_k3", "quirks_k3_nonmax", "syn_k9_min", "syn_k31_min"])
def test_cli_or_mode_needs_no_replay_but_can_do_it(case):
    """Plain `fmsi query` [-S] prints the OR over both strands whatever the predictor says, so the CLI asks for one value
    per k-mer by default (test_cli_matches_reference_outputs covers that on every case); with FMSI_GPU_STRANDS=both it takes
    both strands and replays the predictor as for `-O` / `lookup` — the output must be the same reference bytes."""
    d = os.path.join(GOLDEN, case)
    cmds = [c for c in json.load(open(os.path.join(d, "meta.json")))["cmds"] if c in ("query", "query_S")]
    runs = run_many([ARGS[cmd] + ["-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")] for cmd in cmds], env={"FMSI_GPU_STRANDS": "both"})
    for cmd, r in zip(cmds, runs):
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == open(os.path.join(d, f"exp_{cmd}.txt"), "rb").read(), f"{case}/{cmd}"


def test_cli_small_batches_stdin_gzip_and_k_flag(tmp_path):
    d = os.path.join(GOLDEN, "syn_k9_max")
    want = open(os.path.join(d, "exp_lookup_S.txt"), "rb").read()
    q = open(os.path.join(d, "q.fa"), "rb").read()
    # tiny batches: predictor state must carry across GPU batches
    r = run_cli(["lookup", "-S", "-k", "9", "-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")], env={"FMSI_GPU_BATCH_BASES": "500"})
    assert r.returncode == 0 and r.stdout == want
    # stdin (default and "-"), gzip input
    r = run_cli(["lookup", "-S", os.path.join(d, "ms.fa")], stdin=q)
    assert r.stdout == want
    gz = tmp_path / "q.fa.gz"
    gz.write_bytes(gzip.compress(q))
    r = run_cli(["lookup", "-S", "-q", str(gz), os.path.join(d, "ms.fa")])
    assert r.stdout == want


@pytest.mark.parametrize("case", ["syn_k31_max", "quirks_k3_nonmax", "integration_a"])
def test_cli_long_records_are_cut_into_pieces(case):
    """Blocks beyond $FMSI_GPU_GIANT_BLOCK bytes (64 MB by default: a chromosome-sized record) are laid out by the reader
    in pieces of $FMSI_GPU_PIECE_RESULTS results that run through the pipeline one after the other. With both limits
    forced down to almost nothing every record of the goldens is cut, mid-run, and the output — predictor replay
    across the pieces included — must still be the reference's, byte for byte."""
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    for piece in ("1", "37"):
        env = {"FMSI_GPU_GIANT_BLOCK": "1", "FMSI_GPU_PIECE_RESULTS": piece}
        runs = run_many([ARGS[cmd] + ["-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")] for cmd in meta["cmds"]], env=env)
        for cmd, r in zip(meta["cmds"], runs):
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout == open(os.path.join(d, f"exp_{cmd}.txt"), "rb").read(), f"{case}/{cmd}/piece={piece}"


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fmsi")), reason="oracle/_ref/fmsi not shipped")
@pytest.mark.parametrize("k", [4, 6])
def test_cli_general_mode_mixed_case_palindromes(k, tmp_path):
    """`query -f xor|INT-INT` on soft-masked queries: the reference compares the ASCII k-mer with its reverse complement
    (AreStringsEqual, fms_index.h:319), so a palindromic k-mer with an asymmetric case pattern (`acGT`) counts twice
    there. Live differential against the reference binary, even k, every demasking function."""
    import numpy as np
    ref = os.path.join(ROOT, "oracle", "_ref", "fmsi")
    rng = np.random.default_rng(k)
    half = ["ACGT"[i] for i in rng.integers(0, 4, size=k // 2)]
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    pal = "".join(half) + "".join(comp[c] for c in reversed(half))
    body = "".join("ACGT"[i] for i in rng.integers(0, 4, size=400))
    ms_text = (pal + body[:50] + pal + body[50:200] + pal.lower() + body[200:]).upper()
    mask = rng.random(len(ms_text)) < 0.6
    ms = "".join(c if m else c.lower() for c, m in zip(ms_text, mask))
    ms = ms[:len(ms) - (k - 1)] + ms[len(ms) - (k - 1):].lower()
    fa = tmp_path / "ms.fa"
    fa.write_text(">ms\n" + ms + "\n")
    subprocess.run([ref, "index", "-k", str(k), str(fa)], check=True, capture_output=True)
    variants = [pal, pal.lower(), pal[:k // 2].lower() + pal[k // 2:], pal[:1].lower() + pal[1:], pal[:1].lower() + pal[1:-1] + pal[-1:].lower(),
                pal[:-1] + pal[-1:].lower()]
    recs = []
    for i, v in enumerate(variants):
        recs.append(f">p{i}\n{v}\n")
        recs.append(f">ctx{i}\n{body[:7].lower()}{v}{body[7:13]}{v.swapcase()}N{v}\n")
    recs.append(">rand\n" + "".join(c.lower() if r < 0.5 else c for c, r in zip(body, rng.random(len(body)))) + "\n")
    q = tmp_path / "q.fa"
    q.write_text("".join(recs))
    for f in ("xor", "and", "1-1", "2-2", "1-2", "3-4", "0-0"):
        want = subprocess.run([ref, "query", "-f", f, "-q", str(q), str(fa)], capture_output=True, check=True).stdout
        r = run_cli(["query", "-f", f, "-q", str(q), str(fa)])
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == want, (k, f)


def test_cli_multi_gpu_scheduler_is_byte_identical():
    """$FMSI_GPU_DEVICES shards every batch over index replicas (repeated ordinals share one GPU)."""
    for case, flags, exp in (("syn_k31_max", ["query", "-O", "-S"], "exp_query_OS.txt"), ("syn_k9_min", ["lookup"], "exp_lookup.txt"),
                             ("syn_k31_min", ["query"], "exp_query.txt")):
        d = os.path.join(GOLDEN, case)
        want = open(os.path.join(d, exp), "rb").read()
        for devs in ("0,0,0", "all"):
            r = run_cli([*flags, "-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")], env={"FMSI_GPU_DEVICES": devs})
            assert r.returncode == 0 and r.stdout == want, (case, devs, r.stderr[-300:])


def test_cli_errors_match_reference_behaviour():
    d = os.path.join(GOLDEN, "syn_k31_noklcp")
    r = run_cli(["query", "-S", "-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")])
    assert r.returncode == 1 and b"kLCP array was not constructed" in r.stderr and r.stdout == b""
    r = run_cli(["query", "-k", "30", "-q", os.path.join(d, "q.fa"), os.path.join(d, "ms.fa")])
    assert r.returncode == 1 and b"Mismatch. Provided k (30) does not match the k of the index (31)." in r.stderr
    r = run_cli(["query", "-q", os.path.join(d, "q.fa"), "/nonexistent/prefix"])
    assert r.returncode == 1 and b"index not correctly loaded" in r.stderr
    r = run_cli(["query", "-f", "bogus", os.path.join(d, "ms.fa")])
    assert r.returncode == 1 and b"Function 'bogus' not recognized." in r.stderr
    r = run_cli(["lookup", "-f", "xor", os.path.join(d, "ms.fa")])
    assert r.returncode == 1 and b"Minimum Perfect Hash Function is not allowed" in r.stderr
    r = run_cli(["query", "-h"])
    assert r.returncode == 0 and b"Usage:   fmsi query" in r.stderr
    r = run_cli(["query"])
    assert r.returncode == 1 and b"Path to the fasta file is a required argument" in r.stderr
    r = run_cli(["query", "-q", "/nonexistent/q.fa", os.path.join(d, "ms.fa")])
    assert r.returncode != 0  # the reference aborts on an uncaught std::invalid_argument


@pytest.mark.parametrize("case", golden_cases())
def test_cli_index_writes_the_reference_files(case, tmp_path):
    """`fmsi index [-k K] [-x] MS.fa` (ms_index, reference src/main.cpp:172-236) on the GPU builder: the six
    `.fmsi.*` files next to the input are byte-identical to those the reference's `fmsi index` wrote."""
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    ms = tmp_path / "ms.fa"
    ms.write_bytes(open(os.path.join(d, "ms.fa"), "rb").read())
    flags = ["-k", str(meta["k"])] + ([] if meta["klcp"] else ["-x"])
    r = run_cli(["index", *flags, str(ms)])
    assert r.returncode == 0, r.stderr.decode()
    assert b"Constructed index" in r.stderr and b"Written index" in r.stderr
    for ext in ("ac_gt", "ac", "gt", "mask", "klcp", "misc"):
        want = os.path.join(d, f"ms.fa.fmsi.{ext}")
        got = f"{ms}.fmsi.{ext}"
        if not os.path.exists(want):
            assert not os.path.exists(got), ext
            continue
        assert open(got, "rb").read() == open(want, "rb").read(), f"{case}: {ext}"


def test_cli_index_infers_k_reads_gzip_and_reports_errors(tmp_path):
    d = os.path.join(GOLDEN, "syn_k9_max")
    text = open(os.path.join(d, "ms.fa"), "rb").read()
    gz = tmp_path / "ms.fa.gz"
    gz.write_bytes(gzip.compress(text + b">second\nACGT\n"))
    r = run_cli(["index", str(gz)])
    assert r.returncode == 0, r.stderr.decode()
    assert b"Inferred k from the masked case convention: 9" in r.stderr and b"more than one entry" in r.stderr
    for ext in ("ac_gt", "ac", "gt", "mask", "klcp", "misc"):
        assert open(f"{gz}.fmsi.{ext}", "rb").read() == open(os.path.join(d, f"ms.fa.fmsi.{ext}"), "rb").read(), ext
    r = run_cli(["index", "-k", "7", str(gz)])
    assert r.returncode == 0 and b"does not match the k inferred from the mask convention (9)" in r.stderr
    # multi-line FASTA with CRLF line ends and a FASTQ record hold the same superstring (kseq semantics, parser.h:41-55)
    name, seq = text.split(b"\n")[:2]
    for fn, body in (("multi.fa", name + b" comment\r\n" + b"\r\n".join(seq[i:i + 61] for i in range(0, len(seq), 61)) + b"\r\n"),
                     ("one.fq", b"@" + name[1:] + b"\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n")):
        alt = tmp_path / fn
        alt.write_bytes(body)
        r = run_cli(["index", str(alt)])
        assert r.returncode == 0, r.stderr.decode()
        for ext in ("ac_gt", "ac", "gt", "mask", "klcp", "misc"):
            assert open(f"{alt}.fmsi.{ext}", "rb").read() == open(os.path.join(d, f"ms.fa.fmsi.{ext}"), "rb").read(), (fn, ext)
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b">x\n\n")
    r = run_cli(["index", str(empty)])
    assert r.returncode == 1 and b"is in incorrect format" in r.stderr
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">x\nACGTNACGTacg\n")
    r = run_cli(["index", str(bad)], env={"FMSI_REFERENCE_BIN": ""})
    assert r.returncode != 0 and b"other than ACGTacgt" in r.stderr
    r = run_cli(["index"])
    assert r.returncode == 1 and b"Path to the masked superstring is a required argument" in r.stderr
    r = run_cli(["index", "-h"])
    assert r.returncode == 0 and b"Usage:   fmsi index" in r.stderr
