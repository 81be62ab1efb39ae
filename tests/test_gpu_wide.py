"""A REAL wide index: more than 2^32 BWT rows, where the device layout carries 64-bit positions, per-superblock counter
bases and 16-byte table entries (the reference indexes any length: construct<uint64_t|__uint128_t>, src/main.cpp:230-231).
Neither the GPU builder (n + 1 < 2^32) nor the reference's own `fmsi index` (hours, 70 GB) can produce such an index in
test time, so fmsi_b200/bin/wide_synth writes the six `.fmsi.*` files of a SYNTHETIC one: a seeded pseudo-random symbol
sequence with one '$' slot, a 63/64-dense mask and a random kLCP vector. rank / update_range / get_range_with_pattern /
infer_presence / kmer_order / extend_range_with_klcp are functions of those bit vectors alone, so the oracle (the
reference's algorithm, loading the same files) and the GPU must agree on them value for value — with rows, interval ends
and lookup ids beyond 2^32. Patterns that occur are obtained by walking the LF-mapping from random rows."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle_ffi import MODE_ALL, MODE_OR, OracleIndex

import fmsi_b200 as fg
from fmsi_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("oracle_built")]

TOOL = os.path.join(ROOT, "fmsi_b200", "bin", "wide_synth")
N_ROWS = (1 << 32) + (1 << 27) + 12345
K = 31


_PREFIX = {}


def wide_prefix(gi):
    return _PREFIX[id(gi)]


def lf_walk(oi, row, steps):
    """Characters preceding the suffix of `row`, nearest first, by walking the LF-mapping with the oracle's rank();
    stops at the '$' row. Returns (codes, last row)."""
    counts = oi.counts()
    out = []
    for _ in range(steps):
        c = next((c for c in range(4) if oi.rank(row + 1, c) - oi.rank(row, c) == 1), None)
        if c is None:  # the '$' slot
            break
        out.append(c)
        row = counts[c] + oi.rank(row, c)
    return out, row


def occurring_sequences(oi, rng, count, length):
    seqs = []
    rows = np.concatenate([rng.integers(0, oi.n, size=count - 4), [oi.n - 1, oi.n - 2, (1 << 32) - 1, 1 << 32]])
    for r in rows.tolist():
        codes, _ = lf_walk(oi, int(r), length)
        if len(codes) == length:
            seqs.append(np.array(codes[::-1], dtype=np.uint8))  # text order: the farthest character first
    return seqs


@pytest.fixture(scope="module")
def wide(tmp_path_factory):
    if not os.path.exists(TOOL):
        pytest.skip("fmsi_b200/bin/wide_synth not built")
    import shutil
    d = tmp_path_factory.mktemp("wide")
    if shutil.disk_usage(str(d)).free < 4 << 30:
        pytest.skip("not enough scratch space for a 4.4 G-row index")
    prefix = str(d / "w")
    subprocess.run([TOOL, prefix, str(N_ROWS), str(K), "7"], check=True, capture_output=True)
    oi = OracleIndex.load(prefix, use_klcp=True)
    gi = fg.Index.load(prefix, use_klcp=True)
    _PREFIX[id(gi)] = prefix
    yield oi, gi
    gi.close()
    oi.close()


def test_wide_index_shape_and_building_blocks(wide):
    oi, gi = wide
    assert gi.n == oi.n == N_ROWS and gi.wide and not gi.dict and gi.multistep == 2 and gi.has_klcp  # 64-bit multi-step counters
    assert gi.counts == oi.counts() and gi.dollar_position == oi.dollar()
    assert gi.mask_ones == oi.mask_rank(oi.n) > (1 << 32)
    rng = np.random.default_rng(1)
    N = oi.n
    ii = np.concatenate([rng.integers(0, N + 1, size=400), rng.integers(1 << 32, N + 1, size=200),
                         [0, N, oi.dollar(), oi.dollar() + 1, (1 << 32) - 1, 1 << 32, (1 << 32) + 1]]).astype(np.uint64)
    cc = rng.integers(0, 4, size=len(ii)).astype(np.uint8)
    assert gi.rank(ii, cc).tolist() == [oi.rank(i, c) for i, c in zip(ii, cc)]
    a, b = rng.integers(0, N + 1, size=400), rng.integers(0, N + 1, size=400)
    lo, hi = np.minimum(a, b).astype(np.uint64), np.maximum(a, b).astype(np.uint64)
    c3 = rng.integers(0, 4, size=400).astype(np.uint8)
    g_lo, g_hi = gi.update_range(lo, hi, c3)
    assert list(zip(g_lo.tolist(), g_hi.tolist())) == [oi.update_range(x, y, c) for x, y, c in zip(lo, hi, c3)]
    tiny = np.minimum(lo + rng.integers(0, 3, size=400).astype(np.uint64), N)
    for s, e in ((lo, hi), (lo, tiny)):
        for mo in (False, True):
            assert gi.infer_presence(s, e, mo).tolist() == [oi.infer_presence(x, y, mo) for x, y in zip(s, e)]
        assert gi.kmer_order_if_present(s, e).tolist() == [oi.kmer_order_if_present(x, y) for x, y in zip(s, e)]


def test_wide_index_queries_match_oracle(wide):
    oi, gi = wide
    rng = np.random.default_rng(2)
    L = "ACGT"
    present = occurring_sequences(oi, rng, 300, K)
    assert len(present) > 250
    kmers = np.concatenate([synth.pack_rows(np.stack(present)), synth.revcomp_packed(synth.pack_rows(np.stack(present[:100])), K),
                            rng.integers(0, 1 << 62, size=300, dtype=np.uint64)])
    strs = ["".join(L[(int(v) >> (2 * (K - 1 - t))) & 3] for t in range(K)) for v in kmers]
    beyond = 0
    for use_table in (False, True):
        s, e = gi.get_range_with_pattern(kmers, K, use_table)
        for q, st in enumerate(strs):
            ws, we = oi.get_range_with_pattern(st)
            if ws == we:
                assert s[q] == e[q]
            else:
                assert (s[q], e[q]) == (ws, we)
                beyond += ws >= (1 << 32)
    assert beyond >= 4  # intervals beyond 2^32 were in fact compared
    s, e = gi.get_range_with_pattern(kmers, K, False)
    ne = s < e
    xs, xe = gi.extend_range_with_klcp(s[ne], e[ne])
    assert list(zip(xs.tolist(), xe.tolist())) == [oi.extend_range_with_klcp(x, y) for x, y in zip(s[ne], e[ne])]
    for mode, out, omode, oord in ((fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False),
                                   (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
        got = gi.query_kmers(kmers, K, mode, out, fg.STRANDS_LAZY)
        assert got.astype(np.int64).tolist() == oi.query_packed(kmers, K, omode, oord).tolist()
        both = gi.query_kmers(kmers, K, mode, out, fg.STRANDS_BOTH)
        want = [oi.kmer_both_strands(st, omode, oord) for st in strs]
        if out == fg.OUT_PRESENCE:
            assert [((int(v) & 3) - 1, ((int(v) >> 2) & 3) - 1) for v in both] == want
        else:
            assert [tuple(r) for r in both.tolist()] == want
            assert max(max(f, r) for f, r in want) > (1 << 32)  # ids beyond 2^32 came back intact
    # streamed reads: 150-base walks (every k-mer of them occurs), some reverse-complemented, some with substitutions
    reads = occurring_sequences(oi, rng, 60, 150)
    for r in range(0, len(reads), 3):
        reads[r] = synth.revcomp_codes(reads[r])
    for r in range(1, len(reads), 3):
        reads[r] = reads[r].copy()
        reads[r][int(rng.integers(0, 150))] ^= 1
    texts = [synth.codes_to_ascii(r) for r in reads]
    allk = np.concatenate([synth.pack_kmers(r, K) for r in reads])
    # the synthetic kLCP vector is random, not the kLCP array of these rows, so a streamed answer is NOT the single-query
    # answer here (on a real index they coincide) — it is whatever query_kmers_streaming (fms_index.h:181-254) computes
    # from these bits. That makes the comparison sharper: the kernel must follow the reference's sequence of
    # extend_range_with_klcp / update_range / restarts step for step, on the same chunks (<= 64 k-mers, a fresh
    # predictor per chunk = forward strand first), to print the same characters.
    bases = b"".join(texts)
    offs, lens, chunks = [], [], []
    for r, t in enumerate(texts):
        nk, p = len(t) - K + 1, 0
        while p < nk:
            m = min(64, nk - p)
            offs.append(150 * r + p)
            lens.append(m + K - 1)
            chunks.append(t[p:p + m + K - 1].decode())
            p += m
    for mode, out, omode, oord in ((fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False), (fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False),
                                   (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
        single = oi.query_packed(allk, K, omode, oord)
        assert np.array_equal(gi.query_reads(texts, K, mode, out, fg.STRANDS_LAZY, False).astype(np.int64), single), (mode, out)
        assert np.array_equal(gi.query_chunks(bases, offs, lens, K, mode, out, fg.STRANDS_LAZY, False).astype(np.int64), single), (mode, out)
        want = []
        for c in chunks:
            oi.reset_predictor()
            txt = oi.query_kmers(c, K, omode, True, oord)
            want += [int(x) for x in txt.split(",")] if oord else [int(ch) for ch in txt]
        got = gi.query_chunks(bases, offs, lens, K, mode, out, fg.STRANDS_LAZY, True).astype(np.int64)
        assert got.tolist() == want, (mode, out, "streamed")
        assert np.array_equal(gi.query_reads(texts, K, mode, out, fg.STRANDS_LAZY, True).astype(np.int64), got), (mode, out, "reads")
    assert gi.query_reads(texts, K, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, False).mean() > 0.5
    # general (f-MS) mode on the wide layout: 1..inf == or
    assert np.array_equal(gi.query_kmers_general(kmers, "1-1000000", K), gi.query_kmers(kmers, K, fg.MODE_OR))


@pytest.mark.parametrize("ms", [0, 3])
def test_wide_index_single_and_three_steps_per_probe(wide, ms):
    """The same index with the multi-step arrays off (one LF-step per probe) and with 3 bases per probe (28 GB of
    192-row sectors at this size): k-mer and streamed answers must equal those of the default load, which the tests above
    hold against the oracle."""
    oi, gi = wide
    rng = np.random.default_rng(3)
    present = occurring_sequences(oi, rng, 200, K)
    kmers = np.concatenate([synth.pack_rows(np.stack(present)), synth.revcomp_packed(synth.pack_rows(np.stack(present[:80])), K),
                            rng.integers(0, 1 << 62, size=200, dtype=np.uint64)])
    reads = occurring_sequences(oi, rng, 30, 150)
    for r in range(0, len(reads), 2):
        reads[r] = synth.revcomp_codes(reads[r])
    texts = [synth.codes_to_ascii(r) for r in reads]
    other = fg.Index.load(wide_prefix(gi), use_klcp=True, multistep=ms)
    try:
        assert other.wide and other.multistep == ms
        for mode, out in ((fg.MODE_OR, fg.OUT_PRESENCE), (fg.MODE_ALL, fg.OUT_PRESENCE), (fg.MODE_OR, fg.OUT_ORDERS)):
            for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
                assert np.array_equal(other.query_kmers(kmers, K, mode, out, strands), gi.query_kmers(kmers, K, mode, out, strands)), (ms, mode, out, strands)
            for streaming in (False, True):
                assert np.array_equal(other.query_reads(texts, K, mode, out, fg.STRANDS_LAZY, streaming),
                                      gi.query_reads(texts, K, mode, out, fg.STRANDS_LAZY, streaming)), (ms, mode, out, streaming)
        want = oi.query_packed(kmers, K, MODE_OR, True)
        assert other.query_kmers(kmers, K, fg.MODE_OR, fg.OUT_ORDERS).astype(np.int64).tolist() == want.tolist()
    finally:
        other.close()
