"""Host logic of the CLI's exact `-S` mode, without a GPU: the strand-predictor replay is split into per-chunk
summaries (any thread) and an O(1)-per-chunk in-order pass; fmsi_b200/bin/predictor_check compares that split with
the plain sequential replay on 2000 random cases (presence / ids, or / -O, consistent and conflicting strands)."""
import os
import subprocess

from conftest import ROOT


def test_replay_split_equals_sequential_replay():
    tool = os.path.join(ROOT, "fmsi_b200", "bin", "predictor_check")
    r = subprocess.run([tool], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("all ok"), r.stdout[-500:] + r.stderr[-500:]
