"""The identity behind the multi-step rank arrays (fmsi_b200/csrc/multistep.cuh), checked on the CPU with the oracle:
m applications of update_range (reference src/fms_index.h:98-103) with the bases c_1, ..., c_m equal ONE lookup
    LF_m(i, x) = C_m[x] + #{rows r < i whose suffix is preceded by the m-mer x},   x = c_1 | c_2 << 2 | ...
where c_1 = BWT[r] is the base right before the suffix of row r, c_2 = BWT[LF(r)] the one before that, rows reached
through the '$' slot are preceded by no m-mer, and C_m[x] is where the m steps take position 0. The device kernels are
tested against the oracle on the GPU (tests/test_gpu_parity.py, variants ms*); this pins the arithmetic itself."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle_ffi import OracleIndex

pytestmark = pytest.mark.usefixtures("oracle_built")


def preceding_codes(oi, m):
    """code[r] = c_1 | c_2 << 2 | ... of the m characters before the suffix of row r, or -1 (fewer than m exist)."""
    n, counts, dollar = oi.n, oi.counts(), oi.dollar()
    rank = np.zeros((4, n + 1), dtype=np.int64)
    for c in range(4):
        rank[c] = [oi.rank(i, c) for i in range(n + 1)]
    sym = np.full(n, -1, dtype=np.int64)
    for c in range(4):
        sym[np.flatnonzero(np.diff(rank[c]) == 1)] = c
    assert sym[dollar] == -1 and (sym >= 0).sum() == n - 1
    code = np.zeros(n, dtype=np.int64)
    cur = np.arange(n)
    ok = np.ones(n, dtype=bool)
    for s in range(m):
        c = np.where(ok, sym[np.where(ok, cur, 0)], -1)
        ok &= c >= 0
        code |= np.where(ok, c, 0) << (2 * s)
        cur = np.where(ok, np.asarray(counts)[np.where(ok, c, 0)] + rank[np.where(ok, c, 0), np.where(ok, cur, 0)], 0)
    return np.where(ok, code, -1)


@pytest.mark.parametrize("case", ["quirks_k3", "quirks_k3_nonmax", "fixture3_CACaCat_k3", "syn_k5_min"])
@pytest.mark.parametrize("m", [2, 3])
def test_m_single_steps_equal_one_multi_step(case, m):
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    oi = OracleIndex.load(os.path.join(d, "ms.fa"), use_klcp=meta["klcp"])
    n = oi.n
    if n > 6000:
        pytest.skip("pure-Python rank tables: small indexes only")
    code = preceding_codes(oi, m)
    rng = np.random.default_rng(m)
    for x in range(4 ** m):
        bases = [(x >> (2 * s)) & 3 for s in range(m)]
        i0 = 0
        for c in bases:  # C_m[x]: where the m steps take position 0
            i0 = oi.counts()[c] + oi.rank(i0, c)
        prefix = np.concatenate([[0], np.cumsum(code == x)])
        for i in np.unique(np.concatenate([rng.integers(0, n + 1, size=40), [0, n, oi.dollar(), oi.dollar() + 1]])).tolist():
            want = i
            for c in bases:  # m single LF-steps of one interval end: counts[c] + rank(., c)  (fms_index.h:100-101)
                want = oi.counts()[c] + oi.rank(want, c)
            assert i0 + int(prefix[i]) == want, (case, m, x, i)
    oi.close()
