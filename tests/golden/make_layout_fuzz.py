#!/usr/bin/env python3
"""Writes tests/golden/layout_fuzz/: a corpus of awkward query records (invalid characters at every position class,
lower case, IUPAC codes, empty and short sequences, multi-line FASTA, FASTQ, long records) and, for k = 3 / 9 / 31,
the number of results the REFERENCE binary prints per record (oracle/_ref/fmsi query on golden indexes). Run in the
build container (needs oracle/_ref/fmsi); tests/test_layout.py compares the product's record layout with these."""
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "fmsi")
INDEX = {3: "quirks_k3", 9: "syn_k9_max", 31: "syn_k31_max"}


def corpus(seed: int, n: int) -> bytes:
    rng = random.Random(seed)
    out = []
    for r in range(n):
        kind = rng.randrange(10)
        length = rng.choice([0, 1, 2, 3, 4, 8, 9, 10, 30, 31, 32, 33, 61, 62, 63, 100, 150, 400, 1000]) if kind else rng.choice([2000, 5000])
        s = [rng.choice("ACGT") for _ in range(length)]
        if kind in (1, 2, 3) and length:  # invalid characters: anywhere, at the ends, in runs
            for _ in range(rng.randrange(1, 6)):
                p = rng.choice([0, length - 1, rng.randrange(length), max(0, length - rng.randrange(1, 40)), min(length - 1, rng.randrange(0, 40))])
                for q in range(p, min(length, p + rng.choice([1, 1, 1, 2, 5]))):
                    s[q] = rng.choice("NnRYKM-.*")
        if kind == 4:
            s = [c.lower() if rng.random() < 0.5 else c for c in s]
        seq = "".join(s)
        name = f"r{r}" + (" some comment" if rng.random() < 0.2 else "")
        if kind == 5 and length:  # FASTQ
            out.append(f"@{name}\n{seq}\n+\n{'I' * len(seq)}\n")
        elif kind == 6 and length > 10:  # multi-line FASTA, CRLF
            w = rng.choice([7, 60, 80])
            out.append(f">{name}\r\n" + "\r\n".join(seq[i:i + w] for i in range(0, len(seq), w)) + "\r\n")
        else:
            out.append(f">{name}\n{seq}\n")
    return "".join(out).encode()


def reference_counts(qfile: str, k: int) -> list[tuple[str, int]]:
    prefix = os.path.join(HERE, INDEX[k], "ms.fa")
    r = subprocess.run([REF, "query", "-q", qfile, prefix], capture_output=True, check=True)
    rows = []
    for line in r.stdout.decode().split("\n")[:-1]:
        name, bits = line.split("\t")
        rows.append((name, len(bits)))
    return rows


def main():
    d = os.path.join(HERE, "layout_fuzz")
    os.makedirs(d, exist_ok=True)
    q = os.path.join(d, "q.fa")
    open(q, "wb").write(corpus(20261017, 300))
    for k in INDEX:
        with open(os.path.join(d, f"exp_k{k}.tsv"), "w") as f:
            for name, n in reference_counts(q, k):
                f.write(f"{name}\t{n}\n")
    print("wrote", d)


if __name__ == "__main__":
    sys.exit(main())
