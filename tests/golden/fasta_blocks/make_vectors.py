#!/usr/bin/env python3
"""Writes vectors.json: random adversarial FASTA/FASTQ byte streams and the records the REFERENCE's
reader (oracle/_ref/kseq_dump = kseq.h + parser.h compiled from /root/reference) yields for them.
Run in the authoring container: python tests/golden/fasta_blocks/make_vectors.py"""
import base64
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_fasta_blocks import KSEQ_DUMP, random_input  # noqa: E402

HAND = [b"", b">", b"@", b">a", b">a\n", b">a\nACGT", b">a\nACGT\n", b"ACGT\n>a\nAC\n", b">a b c\nAC\nGT\n>b\n\n\nA\n",
        b"@q\nACGT\n+\nIIII\n", b"@q\nACGT\n+\nIII\n", b"@q\nACGT\n+\nIIIII\n@r\nAC\n+\nII\n", b"@q\nACGT\n+", b"@q\nACGT\n+\n",
        b"@q\nAC\nGT\n+q\nII\nII\n@r\nA\n+\nI\n", b">a\r\nAC\r\nGT\r\n>b\r\nA\r\n", b">a\nA\r\n\r\n", b"@q\nAC\r\n+\r\nII\r\n@r\nA\n+\nI\n",
        b">a\n>b\n>c\nA\n", b">a\nAC+GT\n+\nxx\n", b"@q\n\n+\n\n@r\nA\n+\nI\n", b"@q\nACGT\n+\n@III\n@r\nA\n+\nI\n", b"x>a\nAC\n", b">a\tb\nAC\n",
        b"@q\nA\r\r\n+\nI\r\r\n\n@r\nC\n+\nI\n", b">a\n\r", b">a\nAC\n\r"]


def main():
    rng = random.Random(7)
    inputs = list(HAND) + [random_input(rng) for _ in range(400)]
    out = []
    for data in inputs:
        r = subprocess.run([KSEQ_DUMP, "-"], input=data, capture_output=True)
        assert r.returncode == 0
        out.append({"input": base64.b64encode(data).decode(), "records": base64.b64encode(r.stdout).decode()})
    json.dump(out, open(os.path.join(HERE, "vectors.json"), "w"), indent=0)
    print(len(out), "vectors")


if __name__ == "__main__":
    main()
