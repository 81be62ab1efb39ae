"""The input side of the host pipeline (fmsi_b200/csrc/fasta_blocks.hpp: serial record-boundary scan
+ per-block parsing on worker threads) must yield exactly the records of the reference's reader
(kseq.h via parser.h) for ANY byte stream, whatever the block size: FASTA, FASTQ, multi-line records,
CR/LF, empty lines, garbage in front of a header, truncated last records.

Committed vectors (tests/golden/fasta_blocks/vectors.json, produced by the reference's own kseq through
oracle/_ref/kseq_dump by make_vectors.py) are always checked; where oracle/_ref/kseq_dump exists a
fresh batch of random inputs is compared live as well."""
import base64
import gzip
import json
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, ROOT

SRC = os.path.join(ROOT, "fmsi_b200", "csrc", "tools", "fasta_dump.cpp")
DUMP = os.path.join(ROOT, "fmsi_b200", "bin", "fasta_dump")
KSEQ_DUMP = os.path.join(ROOT, "oracle", "_ref", "kseq_dump")
VECTORS = os.path.join(GOLDEN, "fasta_blocks", "vectors.json")
BLOCKS = (1, 2, 3, 7, 16, 61, 1 << 20)


@pytest.fixture(scope="module")
def dump_tool():
    hdr = os.path.join(ROOT, "fmsi_b200", "csrc", "fasta_blocks.hpp")
    if not os.path.exists(DUMP) or os.path.getmtime(DUMP) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(DUMP), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", DUMP, SRC, "-lz"], check=True)
    return DUMP


def ours(tool, data: bytes, block: int) -> bytes:
    r = subprocess.run([tool, str(block), "-"], input=data, capture_output=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def random_input(rng: random.Random) -> bytes:
    """Record-like text with every separator the reader gives meaning to sprinkled in at random."""
    style = rng.random()
    out = bytearray()
    if style < 0.25:  # pure noise over the critical alphabet
        alpha = b">@+\n\r \tACGTN"
        return bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 80)))
    for _ in range(rng.randrange(0, 9)):
        if rng.random() < 0.15:
            out += bytes(rng.choice(b"xy \n\r") for _ in range(rng.randrange(0, 5)))  # garbage between records
        fastq = rng.random() < 0.4
        out += b"@" if fastq else b">"
        out += bytes(rng.choice(b"abc1") for _ in range(rng.randrange(0, 4)))
        if rng.random() < 0.3:
            out += rng.choice([b" comment", b"\tc", b"\r", b" "])
        out += rng.choice([b"\n", b"\r\n", b"\n", b""])
        seq_len = 0
        for _ in range(rng.randrange(0, 4)):
            line = bytes(rng.choice(b"ACGTNacgt") for _ in range(rng.randrange(0, 12)))
            if rng.random() < 0.1:
                line += rng.choice([b"\r", b"\r\r", b">", b"@", b"+"])
            out += line + rng.choice([b"\n", b"\r\n", b"\n\n", b"\n"])
            seq_len += len(line)
        if fastq:
            out += b"+" + rng.choice([b"", b"name", b"\r"]) + rng.choice([b"\n", b"\r\n", b""])
            q = seq_len + rng.choice([0, 0, 0, 0, -1, 1, 3])
            qual = bytes(rng.choice(b"I#@>+5") for _ in range(max(q, 0)))
            while qual:
                cut = rng.randrange(1, len(qual) + 1) if rng.random() < 0.3 else len(qual)
                out += qual[:cut] + rng.choice([b"\n", b"\r\n", b"\n"])
                qual = qual[cut:]
            if rng.random() < 0.2:
                out += b"\n"
    if rng.random() < 0.3 and out:
        del out[-rng.randrange(1, min(len(out), 6) + 1):]  # cut the input short
    return bytes(out)


def test_committed_vectors(dump_tool):
    vectors = json.load(open(VECTORS))
    assert len(vectors) >= 300
    for v in vectors:
        data, want = base64.b64decode(v["input"]), base64.b64decode(v["records"])
        for block in BLOCKS:
            assert ours(dump_tool, data, block) == want, (data, block)


def test_gzip_and_large_blocks(dump_tool, tmp_path):
    rng = random.Random(5)
    recs = []
    for r in range(3000):
        recs.append(b">r%d\n%s\n" % (r, bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 300)))))
    data = b"".join(recs)
    want = b"".join(b"%d %d %s%s\n" % (len(b"r%d" % i), len(x.split(b"\n")[1]), b"r%d" % i, x.split(b"\n")[1]) for i, x in enumerate(recs))
    p = tmp_path / "q.fa.gz"
    p.write_bytes(gzip.compress(data))
    for block in (1000, 1 << 16, 1 << 24):
        r = subprocess.run([dump_tool, str(block), str(p)], capture_output=True)
        assert r.returncode == 0 and r.stdout == want


@pytest.mark.skipif(not os.path.exists(KSEQ_DUMP), reason="oracle/_ref/kseq_dump (the reference's own kseq) not built here")
def test_live_against_reference_kseq(dump_tool):
    rng = random.Random(20261017)
    for _ in range(400):
        data = random_input(rng)
        want = subprocess.run([KSEQ_DUMP, "-"], input=data, capture_output=True).stdout
        for block in (1, 5, 1 << 20):
            assert ours(dump_tool, data, block) == want, (data, block)
