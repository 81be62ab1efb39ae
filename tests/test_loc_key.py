"""The key function of the minimizer-bucketed dictionary (fmsi_b200/csrc/loc.cuh), checked on the CPU: the functions are
`__host__ __device__`, and fmsi_b200/bin/loc_key_check (host code only) runs them over every k-mer of small spaces and
over seeded samples at the benchmark geometries. A k-mer and its reverse complement must map to the same (bucket, R) with
opposite strand flags, distinct strand pairs to distinct (bucket, R) — so a row match is exact —, the pick must be a
minimum over both strands' m-mers, and every field must stay inside its bit budget. The kernels built on it face the
oracle on the GPU (tests/test_gpu_parity.py, tiers loc / loc_backward)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

TOOL = os.path.join(ROOT, "fmsi_b200", "bin", "loc_key_check")


def run(*args):
    p = subprocess.run([TOOL, *map(str, args)], capture_output=True, text=True)
    return p.returncode, json.loads(p.stdout)


@pytest.mark.skipif(not os.path.exists(TOOL), reason="fmsi_b200/bin/loc_key_check not built")
@pytest.mark.parametrize("k,m,t", [(6, 3, 2), (7, 4, 3), (5, 5, 5), (3, 2, 1), (2, 2, 1), (8, 3, 3), (9, 8, 7), (1, 1, 1), (4, 1, 1)])
def test_exhaustive_small_spaces(k, m, t):
    rc, out = run(k, m, t)
    assert rc == 0 and out["fits"] and out["violations"] == 0
    assert out["kmers"] == 4 ** k
    assert out["distinct_rows"] == (4 ** k + out["self_complementary"]) // 2  # one row per strand pair


@pytest.mark.skipif(not os.path.exists(TOOL), reason="fmsi_b200/bin/loc_key_check not built")
def test_exhaustive_every_geometry_up_to_k6():
    """Every (k, m, t) with t <= m <= k <= 6: windows of every width (1 .. 6 places), even and odd m (palindromic m-mers
    exist only for even m), bucket depths from one base to the whole minimizer."""
    n = 0
    for k in range(1, 7):
        for m in range(1, k + 1):
            for t in range(1, m + 1):
                rc, out = run(k, m, t)
                assert rc == 0 and out["fits"] and out["violations"] == 0, (k, m, t, out)
                assert out["distinct_rows"] == (4 ** k + out["self_complementary"]) // 2, (k, m, t)
                n += 1
    assert n == 56


@pytest.mark.skipif(not os.path.exists(TOOL), reason="fmsi_b200/bin/loc_key_check not built")
@pytest.mark.parametrize("k,m,t,rbits", [(31, 16, 15, 36), (32, 16, 15, 39), (23, 16, 15, 19), (31, 16, 7, 52), (13, 8, 7, 15), (32, 16, 5, 59)])
def test_sampled_benchmark_geometries(k, m, t, rbits):
    rc, out = run(k, m, t, 100000, 7)
    assert rc == 0 and out["fits"] and out["violations"] == 0 and out["rbits"] == rbits and out["rbits"] + 4 <= 64


@pytest.mark.skipif(not os.path.exists(TOOL), reason="fmsi_b200/bin/loc_key_check not built")
def test_geometries_that_do_not_fit_are_refused():
    for k, m, t in ((31, 16, 2), (32, 16, 4), (31, 17, 15), (31, 16, 17), (5, 6, 3)):
        rc, out = run(k, m, t)
        assert rc == 0 and out["fits"] is False
