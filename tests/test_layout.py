"""Host logic of the CLI without a GPU: how many results each query record prints. The record layout
(fmsi_cli.cpp: layout_record) restates ms_query's record loop (reference src/main.cpp:328-373), whose handling of
invalid characters over-emits in known ways (`ACGTN`, k = 3 -> 5 outputs). fmsi_b200/bin/layout_dump prints the
layout's per-record counts; they must equal the line lengths of the reference binary's own output — committed for
the fuzz corpus under tests/golden/layout_fuzz (made by tests/golden/make_layout_fuzz.py with oracle/_ref/fmsi), and
compared live on fresh corpora where the reference binary exists. The tool also checks that GPU chunks and
reference chunks each cover exactly the k-mer results."""
import importlib.util
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from oracle_ffi import REF_EXE

TOOL = os.path.join(ROOT, "fmsi_b200", "bin", "layout_dump")
FUZZ = os.path.join(GOLDEN, "layout_fuzz")


def layout_counts(k, streaming, qfile, piece_limit=None):
    r = subprocess.run([TOOL, str(k), "1" if streaming else "0", qfile] + ([str(piece_limit)] if piece_limit else []), capture_output=True)
    assert r.returncode == 0, (r.returncode, r.stderr.decode())
    if piece_limit:
        layout_counts.pieces = int(r.stderr.decode().split()[-1])
    rows = []
    for line in r.stdout.decode().split("\n")[:-1]:
        name, total, kmers = line.split("\t")
        rows.append((name, int(total), int(kmers)))
    return rows


@pytest.mark.parametrize("k", [3, 9, 31])
@pytest.mark.parametrize("streaming", [False, True])
def test_result_counts_match_reference_output_lengths(k, streaming):
    want = [tuple(l.split("\t")) for l in open(os.path.join(FUZZ, f"exp_k{k}.tsv")).read().split("\n")[:-1]]
    got = layout_counts(k, streaming, os.path.join(FUZZ, "q.fa"))
    assert len(got) == len(want) == 300
    assert [(n, t) for n, t, _ in got] == [(n, int(t)) for n, t in want]
    assert any(t > m for _, t, m in got) and any(m > 400 for _, _, m in got)  # fillers and multi-chunk records occur


@pytest.mark.parametrize("k", [3, 31])
@pytest.mark.parametrize("limit", [1, 50, 1000])
def test_records_cut_into_pieces_print_the_same(k, limit):
    """A record too long for one batch is laid out in pieces (fmsi_cli.cpp: layout_record with a limit; the CLI does
    this for blocks beyond 64 MB). Cut at the reference's chunk boundaries into pieces of >= `limit` results, the
    pieces of every record must add up to the reference's output lengths, chain up, and cover their k-mers exactly."""
    want = [tuple(l.split("\t")) for l in open(os.path.join(FUZZ, f"exp_k{k}.tsv")).read().split("\n")[:-1]]
    whole = layout_counts(k, True, os.path.join(FUZZ, "q.fa"))
    got = layout_counts(k, True, os.path.join(FUZZ, "q.fa"), piece_limit=limit)
    assert got == whole and [(n, t) for n, t, _ in got] == [(n, int(t)) for n, t in want]
    assert layout_counts.pieces > (300 if limit <= 50 else 3)  # records were in fact cut


def test_known_quirks(tmp_path):
    q = tmp_path / "q.fa"
    q.write_bytes(b">a\nACGTN\n>b\nNNACG\n>c\nACGNTAC\nGTA\n>d\nAC\n>\nACGTACGT\n@fq\nACGTAC\n+\nIIIIII\n")
    assert layout_counts(3, False, str(q)) == [("a", 5, 2), ("b", 3, 1), ("c", 8, 5), ("d", 0, 0), ("", 6, 6), ("fq", 4, 4)]


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/fmsi not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_differential_against_reference_binary(seed, tmp_path):
    spec = importlib.util.spec_from_file_location("make_layout_fuzz", os.path.join(GOLDEN, "make_layout_fuzz.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    q = tmp_path / "q.fa"
    q.write_bytes(mk.corpus(seed, 250))
    for k in (3, 31):
        want = mk.reference_counts(str(q), k)
        got = layout_counts(k, seed % 2 == 0, str(q))
        assert [(n, t) for n, t, _ in got] == want, (seed, k)
