#!/usr/bin/env python3
"""Add f-MS (general demasking function) goldens to the existing cases: stdout of the UNMODIFIED
reference binary for `query -f xor|and|1-1|2-9` (and `-S -f xor`, which the reference answers without
kLCP). Run in the authoring container only (needs oracle/_ref/fmsi); outputs are committed.
Also copies the reference's own committed golden tests/testfiles/result_b_complements_xor.txt."""
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "fmsi")
REFSRC = os.environ.get("FMSI_REFERENCE", "/root/reference")
F_CMDS = {"query_f_xor": ["query", "-f", "xor"], "query_f_and": ["query", "-f", "and"], "query_f_1-1": ["query", "-f", "1-1"],
          "query_f_2-9": ["query", "-f", "2-9"], "query_S_f_xor": ["query", "-S", "-f", "xor"]}

for name in sorted(os.listdir(HERE)):
    d = os.path.join(HERE, name)
    mp = os.path.join(d, "meta.json")
    if not os.path.exists(mp):
        continue
    meta = json.load(open(mp))
    done = []
    for tag, args in F_CMDS.items():
        if "-S" in args and not meta["klcp"]:
            continue
        r = subprocess.run([REF] + args + ["-q", "q.fa", "ms.fa"], cwd=d, capture_output=True)
        assert r.returncode == 0, (name, tag, r.stderr.decode())
        with open(os.path.join(d, f"exp_{tag}.txt"), "wb") as f:
            f.write(r.stdout)
        done.append(tag)
    meta["cmds_f"] = done
    json.dump(meta, open(mp, "w"), indent=1)
    print(name, done)
shutil.copy(os.path.join(REFSRC, "tests", "testfiles", "result_b_complements_xor.txt"), os.path.join(HERE, "integration_b", "ref_golden_query_f_xor.txt"))
