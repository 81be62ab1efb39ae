#!/usr/bin/env python3
"""Golden hashes of the REFERENCE's `fmsi index` output on large deterministic inputs.

Run where oracle/_ref/fmsi exists (built from /root/reference by oracle/Makefile):

    python tests/golden/make_ref_index_hashes.py

For every case below the seeded masked superstring of fmsi_b200.synth is written as FASTA, the unmodified
reference indexes it (QSufSort; ~110 s and 1.6 GB at 100 Mbp), and the SHA-256 of each `.fmsi.*` file is
stored in tests/golden/ref_index_hashes.json. tests/test_gpu_build.py rebuilds the same inputs with the GPU
builder on the box and compares the hashes — a byte-for-byte check at sizes where shipping the files
themselves (50-150 MB) would not be reasonable.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fmsi_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "fmsi")
OUT = os.path.join(ROOT, "tests", "golden", "ref_index_hashes.json")

CASES = {
    # name: genome length, seed, k, fraction of OFF positions in the mask, kLCP (`fmsi index` without / with -x)
    "iid_100m_k31": dict(n=100_000_000, seed=314, k=31, off=0.05, klcp=True),
    "iid_20m_k23_noklcp": dict(n=20_000_000, seed=2718, k=23, off=0.5, klcp=False),
}


def masked_superstring(case: dict) -> bytes:
    return synth.random_masked_superstring(case["n"], case["seed"], case["k"], case["off"])


def sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name, case in CASES.items():
        if name in res and "--force" not in sys.argv:
            continue
        with tempfile.TemporaryDirectory(prefix="fmsi_gold_") as d:
            fa = os.path.join(d, "ms.fa")
            synth.write_fasta_single(fa, "ms", masked_superstring(case))
            t0 = time.time()
            subprocess.run([REF, "index", "-k", str(case["k"])] + ([] if case["klcp"] else ["-x"]) + [fa], check=True, capture_output=True)
            exts = ["ac_gt", "ac", "gt", "mask", "misc"] + (["klcp"] if case["klcp"] else [])
            res[name] = dict(case, reference_index_s=round(time.time() - t0, 1), input_sha256=sha(fa),
                             files={e: sha(fa + ".fmsi." + e) for e in exts},
                             sizes={e: os.path.getsize(fa + ".fmsi." + e) for e in exts})
            print(name, res[name]["reference_index_s"], "s", flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
