"""GPU parity tests: every call goes through the C-ABI of libfmsi_gpu.so (ctypes) and is compared
bit for bit with the oracle on the same inputs, with the reference's unit goldens, and with the
reference binary's committed outputs. Integer work: the tolerance is zero."""
import json
import os
import subprocess
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle_ffi import FIXTURES, MODE_ALL, MODE_OR, REF_EXE, OracleIndex

import fmsi_b200 as fg
from fmsi_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("oracle_built")]

L = "ACGT"
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | L.index(ch.upper())
    return v


def gpu_fixture(n, **kw):
    f = FIXTURES[n]
    return fg.Index.from_bits(f["ac_gt"], f["ac"], f["gt"], f["mask"], f["counts"], f["dollar"], f["klcp"], k=3, **kw)


def test_library_is_native_and_device_present():
    assert os.path.exists(fg.lib_path())
    assert fg.device_count() >= 1
    assert fg.lib().fmsi_gpu_abi_version() == 1


# ---- the reference's unit goldens against the device building blocks ------------------------------
@pytest.mark.parametrize("t", [0, 1, -1])
def test_unit_goldens_on_device(t):
    idx = gpu_fixture(1, prefix_t=t)
    cases = [(4, 3, 1), (1, 3, 0), (1, 2, 1), (5, 0, 1), (7, 1, 1), (7, 2, 2), (8, 2, 3), (0, 2, 0)]  # RANK :71-94
    got = idx.rank([c[0] for c in cases], [c[1] for c in cases])
    assert got.tolist() == [c[2] for c in cases]
    ur = [(0, 8, 0, 1, 3), (0, 5, 0, 1, 2), (4, 5, 0, 1, 2), (5, 6, 0, 2, 3), (0, 8, 1, 3, 4), (0, 8, 2, 4, 7), (0, 8, 3, 7, 8), (0, 2, 0, 1, 1)]
    gi, gj = idx.update_range([c[0] for c in ur], [c[1] for c in ur], [c[2] for c in ur])  # UPDATE_RANGE :116-142
    assert gi.tolist() == [c[3] for c in ur] and gj.tolist() == [c[4] for c in ur]
    idx2 = gpu_fixture(2, prefix_t=t)
    assert idx2.rank([4, 5, 6], [0, 0, 0]).tolist() == [2, 3, 3]  # RANK2 :96-114

    idx3 = gpu_fixture(3, prefix_t=t)
    ex = [(4, 6, 4, 7), (4, 5, 4, 7), (5, 6, 4, 7), (2, 3, 1, 3), (1, 2, 1, 3), (3, 4, 3, 4)]  # EXTEND_RANGE_WITH_KLCP :144-167
    gi, gj = idx3.extend_range_with_klcp([c[0] for c in ex], [c[1] for c in ex])
    assert gi.tolist() == [c[2] for c in ex] and gj.tolist() == [c[3] for c in ex]
    for pat, wi, wj in [("ACA", 1, 3), ("CAC", 4, 6), ("CAT", 6, 7), ("AAA", 1, 1), ("TAC", 8, 8), ("A", 1, 4), ("CA", 4, 7), ("T", 7, 8)]:
        for use_table in (False, True):  # GET_RANGE_WITH_PATTERN :169-196
            gi, gj = idx3.get_range_with_pattern([pack(pat)], len(pat), use_table)
            if wi == wj:
                assert gi[0] == gj[0]  # empty (through the table any i == j stands for empty)
                if not use_table:
                    assert (gi[0], gj[0]) == (wi, wj)
            else:
                assert (gi[0], gj[0]) == (wi, wj)
    ko = [(1, 2, 0), (2, 3, -1), (1, 1, -1), (3, 4, -1), (4, 6, 1), (5, 6, 2), (1, 6, 0)]  # KMER_ORDER_IF_PRESENT :198-220
    assert idx3.kmer_order_if_present([c[0] for c in ko], [c[1] for c in ko]).tolist() == [c[2] for c in ko]


TIER_KW = {"auto": {}, "backward": {"dict": 0}, "backward_t0": {"dict": 0, "prefix_t": 0}, "backward_ms0": {"dict": 0, "multistep": 0},
           "backward_ms3": {"dict": 0, "multistep": 3}, "backward_t0_ms2": {"dict": 0, "prefix_t": 0, "multistep": 2},
           # the wide layout (N >= 2^32) forced on small indexes: 64-bit positions, multi-step sectors of 192 rows
           "wide": {"sb_shift_log2": 1, "prefix_t": 1}, "wide_ms3": {"sb_shift_log2": 2, "prefix_t": 0, "multistep": 3},
           # the minimizer-bucketed dictionary (loc.cuh) answers chunks / reads with presence outputs, next to either tier
           "loc": {"locality": 1}, "loc_backward": {"dict": 0, "locality": 1}}


@pytest.mark.parametrize("tier", list(TIER_KW))
def test_streaming_and_query_goldens_on_device(tier):
    """The reference's streaming / query unit goldens on every tier: with `auto` a dictionary tier answers the
    chunks (fmsi_gpu.cu: via_kmers), with dict = 0 stream_kernel / query_kmers_kernel do — with and without the
    suffix table, with single and multi-step probes."""
    kw = TIER_KW[tier]
    idx3 = gpu_fixture(3, **kw)
    assert bool(idx3.dict) == (tier in ("auto", "loc")) and bool(idx3.locality) == tier.startswith("loc")
    # QUERY_KMERS_STREAMING :223-249 and _ORDERS :251-276 (results here do not depend on the predictor)
    for q, mo, want in [("CACATACA", False, "111001"), ("TGTATGTG", False, "100111"), ("CACATTGT", False, "111001"), ("CACATACA", True, "111001")]:
        got = idx3.query_chunks(q.encode(), [0], [len(q)], k=3, mode=fg.MODE_ALL if mo else fg.MODE_OR, streaming=True)
        assert "".join(map(str, got.tolist())) == want
        got = idx3.query_chunks(q.encode(), [0], [len(q)], k=3, mode=fg.MODE_ALL if mo else fg.MODE_OR, streaming=False)
        assert "".join(map(str, got.tolist())) == want
    for q, want in [("CACATACA", [1, 0, 3, -1, -1, 0]), ("TGTATGTG", [0, -1, -1, 3, 0, 1]), ("CACATTGT", [1, 0, 3, -1, -1, 0])]:
        for streaming in (True, False):
            got = idx3.query_chunks(q.encode(), [0], [len(q)], k=3, output=fg.OUT_ORDERS, streaming=streaming)
            assert got.tolist() == want
    idx1 = gpu_fixture(1, **kw)
    # QUERY_ORDERS :278-306 / QUERY :308-332 (k varies per case)
    for q, k, want in [("A", 1, [3]), ("AG", 2, [-1]), ("CA", 2, [0]), ("AC", 2, [2]), ("TA", 2, [3]), ("GGTA", 4, [1]), ("ATGG", 4, [-1]),
                       ("GA", 2, [-1]), ("GGG", 3, [-1]), ("CC", 2, [1]), ("CCAG", 2, [1, 0, -1])]:
        assert idx1.query_chunks(q.encode(), [0], [len(q)], k=k, output=fg.OUT_ORDERS).tolist() == want
    for q, want in [("A", 1), ("AG", 0), ("CA", 1), ("GGTA", 1), ("ATGG", 0), ("GA", 0), ("GGG", 0), ("CC", 1)]:
        assert idx1.query_kmers([pack(q)], k=len(q)).tolist() == [want]
    idx2 = gpu_fixture(2, **kw)
    for q, want in [("AAGA", 1), ("AAGAA", 0), ("GGTTAAGA", 1), ("GTTAAGA", 1)]:  # QUERY2 :334-354
        assert idx2.query_kmers([pack(q)], k=len(q)).tolist() == [want]


# ---- golden indexes: device vs oracle on seeded random inputs -----------------------------------
def _random_kmers(rng, ms_codes, k, n):
    """Half from the superstring (random strand), half random; plus a few edge k-mers."""
    pos = rng.integers(0, len(ms_codes) - k + 1, size=n // 2)
    win = ms_codes[pos[:, None] + np.arange(k)[None, :]]
    flip = rng.integers(0, 2, size=len(win)).astype(bool)
    win[flip] = 3 - win[flip][:, ::-1]
    rnd = rng.integers(0, 4, size=(n - len(win), k), dtype=np.uint8)
    rows = np.concatenate([win, rnd, np.zeros((1, k), np.uint8), np.full((1, k), 3, np.uint8)])
    return synth.pack_rows(rows.astype(np.uint8))


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("variant", ["auto", "t0", "wide", "wide_ms0", "wide_ms3_t0", "nodict", "dict_t1", "dict_t3", "dict_auto", "fold_t1",
                                     "fold_t3", "ms0", "ms2_t0", "ms2_t2", "ms3", "ms3_t1"])
def test_device_matches_oracle(case, variant):
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    k = meta["k"]
    # dict_t1 / dict_t3: shallow dictionary buckets hold far more rows than fit a sector, which drives
    # the ROWS phase and the overflow list -> backward-search fixup launch of dict.cuh. fold_t1 / fold_t3: the
    # same for the strand-folded dictionary (fold.cuh): every bucket is binary-searched in rows[].
    # auto = strand-folded dictionary at the automatic depth; dict_auto = SA-ordered dictionary there.
    # wide* = the layout of indexes with 2^32 rows and more (64-bit positions, superblock counter bases), forced on small
    # indexes; its multi-step sectors carry a 64-bit counter and 192 rows (auto = 2 bases per probe, like nodict)
    kw = {"auto": {}, "t0": {"prefix_t": 0}, "wide": {"sb_shift_log2": 1, "prefix_t": 2}, "nodict": {"dict": 0},
          "wide_ms0": {"sb_shift_log2": 1, "prefix_t": 2, "multistep": 0}, "wide_ms3_t0": {"sb_shift_log2": 2, "prefix_t": 0, "multistep": 3},
          "dict_t1": {"dict": 1, "prefix_t": 1}, "dict_t3": {"dict": 1, "prefix_t": 3}, "dict_auto": {"dict": 1},
          "fold_t1": {"dict": 2, "prefix_t": 1}, "fold_t3": {"dict": 2, "prefix_t": 3},
          # backward search with the multi-step rank arrays (multistep.cuh): off / 2 / 3 bases per probe, at table depths
          # that leave k - t a multiple of m or not (the remainder takes single steps); nodict = auto = 2 per probe
          "ms0": {"dict": 0, "multistep": 0}, "ms2_t0": {"dict": 0, "multistep": 2, "prefix_t": 0},
          "ms2_t2": {"dict": 0, "multistep": 2, "prefix_t": 2}, "ms3": {"dict": 0, "multistep": 3},
          "ms3_t1": {"dict": 0, "multistep": 3, "prefix_t": 1}}[variant]
    if variant != "auto" and case not in ("syn_k31_max", "syn_k9_min", "syn_k5_min", "quirks_k3", "data_k13", "syn_k32"):
        pytest.skip("variants run on a subset")
    prefix = os.path.join(d, "ms.fa")
    gi = fg.Index.load(prefix, use_klcp=meta["klcp"], **kw)
    oi = OracleIndex.load(prefix, use_klcp=meta["klcp"])
    assert gi.n == oi.n and gi.k == k and gi.counts == oi.counts() and gi.dollar_position == oi.dollar()
    assert gi.wide == variant.startswith("wide")
    assert gi.dict == (variant not in ("t0", "nodict") and not variant.startswith(("ms", "wide")) and k <= 32)
    if variant.startswith(("ms", "wide")) or variant == "nodict":
        assert gi.multistep == {"ms0": 0, "ms3": 3, "ms3_t1": 3, "wide_ms0": 0, "wide_ms3_t0": 3}.get(variant, 2)
    if gi.dict:  # the tier asked for, unless the payload would not fit a row (k - t > 30: k = 32 at depth 1)
        fold_ok = k - gi.dict_t <= 30
        assert gi.dict_kind == (2 if variant in ("auto", "fold_t1", "fold_t3") and fold_ok else 1)
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    N = gi.n
    # rank / update_range
    ii = np.concatenate([rng.integers(0, N + 1, size=300), [0, N, oi.dollar(), oi.dollar() + 1, min(N, 64), min(N, 63)]]).astype(np.uint64)
    cc = rng.integers(0, 4, size=len(ii)).astype(np.uint8)
    assert gi.rank(ii, cc).tolist() == [oi.rank(i, c) for i, c in zip(ii, cc)]
    a = rng.integers(0, N + 1, size=300)
    b = rng.integers(0, N + 1, size=300)
    lo, hi = np.minimum(a, b).astype(np.uint64), np.maximum(a, b).astype(np.uint64)
    c3 = rng.integers(0, 4, size=300).astype(np.uint8)
    g_lo, g_hi = gi.update_range(lo, hi, c3)
    want = [oi.update_range(x, y, c) for x, y, c in zip(lo, hi, c3)]
    assert list(zip(g_lo.tolist(), g_hi.tolist())) == want
    # mask: infer_presence<>, kmer_order_if_present on random and tiny intervals
    tiny = np.minimum(lo + rng.integers(0, 3, size=300).astype(np.uint64), N)
    for s, e in ((lo, hi), (lo, tiny)):
        for mo in (False, True):
            assert gi.infer_presence(s, e, mo).tolist() == [oi.infer_presence(x, y, mo) for x, y in zip(s, e)]
        assert gi.kmer_order_if_present(s, e).tolist() == [oi.kmer_order_if_present(x, y) for x, y in zip(s, e)]
    if k > 32:  # packed k-mers stop at k = 32: the text path is test_long_k_chunks_match_oracle
        gi.close()
        oi.close()
        return
    # k-mers
    ms = open(prefix, "rb").read().split(b"\n")[1]
    ms_codes = synth.ascii_to_codes(ms)
    kmers = _random_kmers(rng, ms_codes, k, 600)
    strs = ["".join(L[(int(v) >> (2 * (k - 1 - t))) & 3] for t in range(k)) for v in kmers]
    for use_table in (False, True):
        s, e = gi.get_range_with_pattern(kmers, k, use_table)
        for q, st in enumerate(strs):
            ws, we = oi.get_range_with_pattern(st)
            if ws == we:
                assert s[q] == e[q]
            else:
                assert (s[q], e[q]) == (ws, we)
    # kLCP extension on real intervals
    if meta["klcp"]:
        s, e = gi.get_range_with_pattern(kmers, k, False)
        ne = s < e
        if ne.any():
            xs, xe = gi.extend_range_with_klcp(s[ne], e[ne])
            assert list(zip(xs.tolist(), xe.tolist())) == [oi.extend_range_with_klcp(x, y) for x, y in zip(s[ne], e[ne])]
    # the hot path, all modes, LAZY (neutral predictor) and BOTH (per-strand values)
    for mode, out, omode, oord in ((fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False),
                                   (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
        got = gi.query_kmers(kmers, k, mode, out, fg.STRANDS_LAZY)
        assert got.astype(np.int64).tolist() == oi.query_packed(kmers, k, omode, oord).tolist()
        both = gi.query_kmers(kmers, k, mode, out, fg.STRANDS_BOTH)
        want = [oi.kmer_both_strands(st, omode, oord) for st in strs]
        if out == fg.OUT_PRESENCE:
            assert [((int(v) & 3) - 1, ((int(v) >> 2) & 3) - 1) for v in both] == want
        else:
            assert [tuple(r) for r in both.tolist()] == want
    gi.close()
    oi.close()


def test_dictionary_tier_on_repetitive_index(tmp_path):
    """Repeats (mutated copies + homopolymer runs) give buckets with many rows at the automatic depth:
    the dictionary answers (bucket scan, ROWS phase, overflow -> fixup launch) must equal the
    backward-search kernel's and the oracle's, in every mode, on a non-max-ones mask."""
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/fmsi not shipped")
    rng = np.random.default_rng(99)
    unit = rng.integers(0, 4, size=3000).astype(np.uint8)
    parts = []
    for c in range(40):
        u = unit.copy()
        sub = rng.random(len(u)) < 0.01
        u[sub] = (u[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
        parts += [u, np.full(int(rng.integers(20, 200)), c & 3, np.uint8)]
    g = np.concatenate(parts)
    for k in (31, 13):
        mask = rng.random(len(g)) < 0.7
        mask[len(g) - (k - 1):] = False
        fa = str(tmp_path / f"rep{k}.fa")
        synth.write_fasta_single(fa, "ms", synth.codes_to_ascii(g, mask))
        subprocess.run([REF_EXE, "index", "-k", str(k), fa], check=True, capture_output=True)
        oi = OracleIndex.load(fa, use_klcp=False)
        # k-mers made of the text's last characters padded with A's equal the padded payload of the
        # invalid (too short) rows: they must not match through those rows
        tails = np.stack([np.concatenate([g[len(g) - m:], np.zeros(k - m, np.uint8)]) for m in range(1, k)])
        kmers = np.concatenate([_random_kmers(rng, g, k, 20000), synth.pack_kmers(g[:5000], k), synth.pack_kmers(g[-300:], k),
                                synth.pack_rows(tails), synth.revcomp_packed(synth.pack_rows(tails), k),
                                synth.pack_rows(np.stack([np.full(k, c, np.uint8) for c in range(4)]))])
        strs = None
        for kw in ({}, {"dict": 2, "prefix_t": 1}, {"dict": 2, "prefix_t": 4}, {"dict": 2, "prefix_t": 7}, {"dict": 1},
                   {"dict": 1, "prefix_t": 1}, {"dict": 1, "prefix_t": 4}, {"dict": 1, "prefix_t": 7}):
            gd = fg.Index.load(fa, use_klcp=False, **kw)
            gb = fg.Index.load(fa, use_klcp=False, dict=0)
            assert gd.dict and not gb.dict and gd.dict_kind == kw.get("dict", 2)
            for mode, out, omode, oord in ((fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False),
                                           (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
                want = oi.query_packed(kmers, k, omode, oord)
                for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
                    a = gd.query_kmers(kmers, k, mode, out, strands)
                    b = gb.query_kmers(kmers, k, mode, out, strands)
                    assert np.array_equal(a, b), (k, kw, mode, out, strands)
                    if strands == fg.STRANDS_LAZY:
                        assert np.array_equal(a.astype(np.int64), want), (k, kw, mode, out)
            gd.close()
            gb.close()
        oi.close()


def test_fold_lookup_ids_are_built_on_demand():
    """The strand-folded dictionary keeps its lookup ids (8 bytes per distinct k-mer) out of device memory until a lookup
    needs them (fmsi_gpu_options.fold_ids: 0 = on the first lookup, 1 = with the tier, -1 = never: lookups then run on
    the backward-search kernels). The answers are the oracle's under every policy."""
    d = os.path.join(GOLDEN, "syn_k31_min")
    prefix = os.path.join(d, "ms.fa")
    k = 31
    oi = OracleIndex.load(prefix, use_klcp=False)
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    kmers = _random_kmers(np.random.default_rng(3), ms_codes, k, 3000)
    want = oi.query_packed(kmers, k, MODE_OR, True)
    sizes = {}
    for policy in (0, 1, -1):
        gi = fg.Index.load(prefix, use_klcp=True, dict=2, fold_ids=policy)
        assert gi.dict_kind == 2 and gi.fold_ids == (policy == 1)
        assert np.array_equal(gi.query_kmers(kmers, k, fg.MODE_ALL).astype(np.int64), oi.query_packed(kmers, k, MODE_ALL, False))
        assert gi.refresh_info().fold_ids == (policy == 1)  # presence queries never build them
        assert np.array_equal(gi.query_kmers(kmers, k, output=fg.OUT_ORDERS), want)
        assert gi.refresh_info().fold_ids == (policy != -1)
        sizes[policy] = gi.hbm_bytes
        reads = list(synth.read_queries(ms_codes, 150, 40, 4))
        bases, offs, lens = _chunks_of(reads, k, 64)
        allk = np.concatenate([synth.pack_kmers(synth.ascii_to_codes(bases[o:o + l]), k) for o, l in zip(offs.tolist(), lens.tolist())])
        assert np.array_equal(gi.query_chunks(bases, offs, lens, k, output=fg.OUT_ORDERS, streaming=True), oi.query_packed(allk, k, MODE_OR, True))
        gi.close()
    assert sizes[0] == sizes[1] > sizes[-1]
    oi.close()


def test_auto_tier_falls_back_loudly_when_memory_is_short():
    """`dict = auto` takes the fastest tier that fits in the free device memory. When that is not the one-probe dictionary the
    caller must be able to see it: a note on stderr, and fmsi_gpu_index_info.dict (round 1 fell back silently — an 11 x cliff
    on a shared GPU). $FMSI_GPU_FREE_CAP stands in for a GPU that is mostly taken."""
    import sys
    prefix = os.path.join(GOLDEN, "syn_k31_max", "ms.fa")
    code = ("import sys; sys.path.insert(0, %r); import fmsi_b200 as fg; gi = fg.Index.load(%r, use_klcp=False); "
            "print(gi.dict_kind, gi.multistep); import numpy as np; print(int(gi.query_kmers(np.zeros(3, np.uint64), 31).sum()))" % (ROOT_DIR, prefix))
    plenty = subprocess.run([sys.executable, "-c", code], capture_output=True)
    assert plenty.returncode == 0 and plenty.stdout.split()[0] == b"2" and b"does not fit" not in plenty.stderr
    short = subprocess.run([sys.executable, "-c", code], capture_output=True, env=dict(os.environ, FMSI_GPU_FREE_CAP=str(70 << 20)))
    assert short.returncode == 0, short.stderr.decode()
    assert short.stdout.split()[0] in (b"0", b"1") and b"strand-folded dictionary does not fit" in short.stderr


def _chunks_of(seq_codes_list, k, max_kmers):
    """Concatenate sequences into one base buffer and cut each into chunks of <= max_kmers k-mers
    overlapping by k-1 (the shape ms_query produces)."""
    bases, offs, lens = bytearray(), [], []
    for codes in seq_codes_list:
        start = len(bases)
        bases += synth.codes_to_ascii(codes)
        n = len(codes)
        p = 0
        while n - p >= k:
            ln = min(n - p, max_kmers + k - 1)
            offs.append(start + p)
            lens.append(ln)
            p += ln - k + 1
    return bytes(bases), np.array(offs, np.uint64), np.array(lens, np.uint32)


@pytest.mark.parametrize("case", ["syn_k31_max", "syn_k31_min", "syn_k9_max", "syn_k9_min", "syn_k5_min", "data_k13", "data_k31", "syn_k32", "quirks_k3_nonmax"])
@pytest.mark.parametrize("tier", ["auto", "backward", "backward_t0", "backward_ms0", "backward_ms3", "wide", "wide_ms3", "loc", "loc_backward"])
def test_chunks_streaming_and_single_match_oracle(case, tier):
    """Chunks of text, streamed (-S) and single, LAZY and BOTH strands, all three outputs, against the oracle. With
    `auto` a dictionary tier answers streamed chunks too; the dict = 0 tiers put stream_kernel (query_kmers_streaming,
    fms_index.h:181-254) itself in front of the oracle — with and without the suffix table, with multi-step probes
    off / 2 / 3 bases per probe after a miss."""
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    k = meta["k"]
    prefix = os.path.join(d, "ms.fa")
    gi = fg.Index.load(prefix, use_klcp=True, **TIER_KW[tier])
    assert bool(gi.dict) == (tier in ("auto", "loc")) and bool(gi.locality) == tier.startswith("loc")
    oi = OracleIndex.load(prefix, use_klcp=True)
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    rng = np.random.default_rng(7)
    seqs = []
    for r in range(60):  # reads from the superstring with substitutions, random strand, assorted lengths
        ln = int(rng.integers(k, min(len(ms_codes), 260)))
        p = int(rng.integers(0, len(ms_codes) - ln + 1))
        s = ms_codes[p:p + ln].copy()
        sub = rng.random(ln) < 0.03
        s[sub] = (s[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
        if r % 2:
            s = synth.revcomp_codes(s)
        seqs.append(s.astype(np.uint8))
    seqs.append(rng.integers(0, 4, size=200).astype(np.uint8))
    seqs.append(ms_codes[:k].copy())
    for max_kmers in (64, 25, 1):
        bases, offs, lens = _chunks_of(seqs, k, max_kmers)
        kmers = np.concatenate([synth.pack_kmers(synth.ascii_to_codes(bases[o:o + l]), k) for o, l in zip(offs.tolist(), lens.tolist())])
        strs = ["".join(L[(int(v) >> (2 * (k - 1 - t))) & 3] for t in range(k)) for v in kmers]
        for mode, out, omode, oord in ((fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False),
                                       (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
            want = oi.query_packed(kmers, k, omode, oord).tolist()
            for streaming in (False, True):
                got = gi.query_chunks(bases, offs, lens, k, mode, out, fg.STRANDS_LAZY, streaming)
                assert got.astype(np.int64).tolist() == want, (case, max_kmers, mode, out, streaming)
            wb = [oi.kmer_both_strands(st, omode, oord) for st in strs]
            for streaming in (False, True):
                both = gi.query_chunks(bases, offs, lens, k, mode, out, fg.STRANDS_BOTH, streaming)
                if out == fg.OUT_PRESENCE:
                    assert [((int(v) & 3) - 1, ((int(v) >> 2) & 3) - 1) for v in both] == wb, (case, max_kmers, mode, streaming)
                else:
                    assert [tuple(r) for r in both.tolist()] == wb, (case, max_kmers, streaming)
    gi.close()
    oi.close()


@pytest.mark.parametrize("case", ["syn_k47_max", "syn_k64_min", "syn_k97_noklcp"])
@pytest.mark.parametrize("variant", ["auto", "t0", "wide", "wide_ms0", "wide_ms3", "ms0", "ms3"])
def test_long_k_chunks_match_oracle(case, variant):
    """k > 32 (longk_kernels.cuh): k-mers are searched from the packed text, 32 pattern characters per
    register window. Every mode, LAZY and BOTH strands, with and without `streaming`, chunks of assorted
    sizes, against the oracle's per-strand values; packed-k-mer entry points must refuse such k."""
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    k = meta["k"]
    prefix = os.path.join(d, "ms.fa")
    kw = {"auto": {}, "t0": {"prefix_t": 0}, "wide": {"sb_shift_log2": 1, "prefix_t": 2}, "ms0": {"multistep": 0}, "ms3": {"multistep": 3},
          "wide_ms0": {"sb_shift_log2": 1, "prefix_t": 2, "multistep": 0}, "wide_ms3": {"sb_shift_log2": 1, "prefix_t": 1, "multistep": 3}}[variant]
    gi = fg.Index.load(prefix, use_klcp=meta["klcp"], **kw)
    assert gi.multistep == {"wide_ms0": 0, "ms0": 0, "ms3": 3, "wide_ms3": 3}.get(variant, 2) and gi.wide == variant.startswith("wide")
    oi = OracleIndex.load(prefix, use_klcp=meta["klcp"])
    assert gi.k == k and not gi.dict
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    rng = np.random.default_rng(zlib.crc32(case.encode()) + 1)
    seqs = []
    for r in range(50):  # reads from the superstring with substitutions, random strand, assorted lengths
        ln = int(rng.integers(k, k + 200))
        p = int(rng.integers(0, len(ms_codes) - ln + 1))
        s = ms_codes[p:p + ln].copy()
        sub = rng.random(ln) < 0.01
        s[sub] = (s[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
        if r % 2:
            s = synth.revcomp_codes(s)
        seqs.append(s.astype(np.uint8))
    seqs.append(rng.integers(0, 4, size=k + 100).astype(np.uint8))
    seqs.append(ms_codes[:k].copy())
    seqs.append(ms_codes[len(ms_codes) - k:].copy())
    half = rng.integers(0, 4, size=(k + 1) // 2).astype(np.uint8)
    seqs.append(np.concatenate([half, synth.revcomp_codes(half)])[:max(k, 2 * (k // 2))])  # self-complementary when k is even
    seqs = [s for s in seqs if len(s) >= k]

    def lazy(f, r, omode, oord):  # query_kmers_single with a neutral predictor, fms_index.h:283-298
        if oord:
            return f if f >= 0 else r
        if omode == MODE_OR:
            return int((f if f == 1 else r) == 1)
        return int((f if f != -1 else r) == 1)

    for max_kmers in (64, 7, 300):
        bases, offs, lens = _chunks_of(seqs, k, max_kmers)
        strs = [bases[o + q:o + q + k].decode() for o, l in zip(offs.tolist(), lens.tolist()) for q in range(l - k + 1)]
        for mode, out, omode, oord in ((fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False),
                                       (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
            wb = [oi.kmer_both_strands(st, omode, oord) for st in strs]
            want = [lazy(f, r, omode, oord) for f, r in wb]
            for streaming in ((False, True) if meta["klcp"] else (False,)):
                got = gi.query_chunks(bases, offs, lens, k, mode, out, fg.STRANDS_LAZY, streaming)
                assert got.astype(np.int64).tolist() == want, (case, max_kmers, mode, out, streaming)
                both = gi.query_chunks(bases, offs, lens, k, mode, out, fg.STRANDS_BOTH, streaming)
                if out == fg.OUT_PRESENCE:
                    assert [((int(v) & 3) - 1, ((int(v) >> 2) & 3) - 1) for v in both] == wb, (case, max_kmers, mode, streaming)
                else:
                    assert [tuple(r) for r in both.tolist()] == wb, (case, max_kmers, streaming)
    with pytest.raises(fg.FmsiGpuError):
        gi.query_kmers(np.zeros(4, np.uint64), k)
    if not meta["klcp"]:
        with pytest.raises(fg.FmsiGpuError):
            gi.query_chunks(bases, offs, lens, k, fg.MODE_OR, fg.OUT_PRESENCE, fg.STRANDS_LAZY, True)
    gi.close()
    oi.close()


@pytest.mark.parametrize("case", ["syn_k31_min", "syn_k9_min", "syn_k5_min", "quirks_k3_nonmax", "syn_k31_max"])
def test_general_demasking_functions_properties(case):
    """fmsi_gpu_query_kmers_general (single_query_general on both strands, fms_index.h:171-179,
    :317-327): cross-checked against the oracle-verified modes through identities of the functions
    of src/functions.h — 1..inf == or; xor == (1-1) + (3-3) + ...; and on a max-ones mask == -O."""
    d = os.path.join(GOLDEN, case)
    k = json.load(open(os.path.join(d, "meta.json")))["k"]
    prefix = os.path.join(d, "ms.fa")
    rng = np.random.default_rng(5)
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    kmers = _random_kmers(rng, ms_codes, k, 4000)
    for kw in ({}, {"prefix_t": 0}, {"sb_shift_log2": 1, "prefix_t": 2}):
        gi = fg.Index.load(prefix, use_klcp=False, **kw)
        orr = gi.query_kmers(kmers, k, fg.MODE_OR)
        ge1 = gi.query_kmers_general(kmers, "1-1000000", k)
        assert np.array_equal(ge1, orr)
        odd = np.zeros(len(kmers), dtype=np.uint8)
        for c in range(1, 40, 2):
            odd |= gi.query_kmers_general(kmers, f"{c}-{c}", k)
        big = gi.query_kmers_general(kmers, "40-1000000", k)
        xor = gi.query_kmers_general(kmers, "xor", k)
        assert np.array_equal(xor[big == 0], odd[big == 0])
        land = gi.query_kmers_general(kmers, "and", k)
        assert not (land & ~orr).any()          # and => or
        if case == "syn_k31_max":                # every occurrence ON: and == or == -O
            assert np.array_equal(land, orr) and np.array_equal(land, gi.query_kmers(kmers, k, fg.MODE_ALL))
        assert np.array_equal(gi.query_kmers_general(kmers, fg.Function(kind=0), k), orr)  # f_or
        gi.close()
    with pytest.raises(fg.FmsiGpuError):
        gi = fg.Index.load(prefix, use_klcp=False)
        try:
            gi.query_kmers_general(kmers, fg.Function(kind=7), k)
        finally:
            gi.close()


@pytest.mark.parametrize("tier", ["auto", "backward", "loc"])
def test_multi_gpu_pool_matches_single_index(tier):
    """The scheduler (fmsi_gpu_pool_*): replicas made by device-to-device copy answer contiguous
    shards from their own host threads; results must equal the single-index call, in query order.
    With one GPU the replicas share it (repeated ordinals), with more they spread over the GPUs.
    auto: replicas of the strand-folded dictionary (each builds its lookup ids on its first lookup);
    backward: replicas carry the multi-step rank arrays; loc: and the minimizer-bucketed dictionary (chunk calls)."""
    d = os.path.join(GOLDEN, "syn_k31_max")
    prefix = os.path.join(d, "ms.fa")
    k = 31
    gi = fg.Index.load(prefix, use_klcp=True, **TIER_KW[tier])
    assert (gi.multistep == 2) == (tier == "backward")
    ndev = fg.device_count()
    devices = [0, 1 % ndev, 2 % ndev, 0]
    pool = fg.Pool(gi, devices)
    assert pool.size == 4
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    rng = np.random.default_rng(17)
    for n in (0, 1, 3, 4, 5003):
        kmers = _random_kmers(rng, ms_codes, k, n)[:n] if n else np.zeros(0, np.uint64)
        for mode, out in ((fg.MODE_ALL, fg.OUT_PRESENCE), (fg.MODE_OR, fg.OUT_PRESENCE), (fg.MODE_OR, fg.OUT_ORDERS)):
            for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
                assert np.array_equal(pool.query_kmers(kmers, k, mode, out, strands), gi.query_kmers(kmers, k, mode, out, strands))
    reads = [r for r in synth.read_queries(ms_codes, 150, 57, 4)] + [ms_codes[:k].copy()]
    for max_kmers in (64, 7):
        bases, offs, lens = _chunks_of(reads, k, max_kmers)
        for streaming in (False, True):
            for out in (fg.OUT_PRESENCE, fg.OUT_ORDERS):
                a = pool.query_chunks(bases, offs, lens, k, fg.MODE_ALL, out, fg.STRANDS_BOTH, streaming)
                b = gi.query_chunks(bases, offs, lens, k, fg.MODE_ALL, out, fg.STRANDS_BOTH, streaming)
                assert np.array_equal(a, b)
    with pytest.raises(fg.FmsiGpuError):
        fg.Pool(gi, [ndev + 7])
    # a replica that fails AFTER the primary and another replica joined the pool, with the primary not first in the
    # device list: the failure path must release the replica and leave the primary alive (it used to free members[1:])
    with pytest.raises(fg.FmsiGpuError):
        fg.Pool(gi, [0, 0, ndev + 7])
    with pytest.raises(fg.FmsiGpuError):
        fg.Pool(gi, [(1 % ndev), 0, ndev + 7])
    kmers = _random_kmers(rng, ms_codes, k, 500)
    assert np.array_equal(pool.query_kmers(kmers, k, fg.MODE_ALL), gi.query_kmers(kmers, k, fg.MODE_ALL))  # primary still usable
    pool.close()
    assert gi.query_kmers(kmers, k, fg.MODE_ALL).shape == (len(kmers),)
    gi.close()


@pytest.mark.parametrize("tier", ["auto", "backward", "loc"])
def test_bit_packed_results_and_packed_text(tier):
    """FMSI_GPU_OUT_PRESENCE_BITS (one bit per k-mer, packed on the device) must equal the byte results packed on the
    host, for every batch shape of the host path; fmsi_gpu_query_chunks_packed (2-bit text in) must equal the ASCII
    call, streamed and single, every output. The byte / ASCII results are the ones checked against the oracle."""
    d = os.path.join(GOLDEN, "syn_k31_min")
    k = 31
    prefix = os.path.join(d, "ms.fa")
    gi = fg.Index.load(prefix, use_klcp=True, **TIER_KW[tier])
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    rng = np.random.default_rng(23)
    for n in (1, 7, 8, 9, 63, 64, 65, 5003, 5_000_011):  # the last one spans several host batches
        kmers = _random_kmers(rng, ms_codes, k, max(n, 4))[:n]
        for mode in (fg.MODE_ALL, fg.MODE_OR):
            by = gi.query_kmers(kmers, k, mode, fg.OUT_PRESENCE)
            bi = gi.query_kmers(kmers, k, mode, fg.OUT_PRESENCE_BITS)
            assert bi.shape == ((n + 7) // 8,)
            assert np.array_equal(bi, np.packbits(by, bitorder="little")), (n, mode)
    reads = list(synth.read_queries(ms_codes, 150, 300, 4)) + [ms_codes[:k].copy(), ms_codes[5:5 + k + 1].copy()]
    for max_kmers in (64, 9):
        bases, offs, lens = _chunks_of(reads, k, max_kmers)
        words = fg.pack_text(np.frombuffer(bases, dtype=np.uint8))
        # the packed layout is the packed k-mer layout: word 0 of the text == the first 32 bases
        assert int(words[0]) == int(synth.pack_kmers(synth.ascii_to_codes(bases[:32]), 32)[0])
        for streaming in (False, True):
            for mode, out, strands in ((fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY), (fg.MODE_OR, fg.OUT_PRESENCE, fg.STRANDS_BOTH),
                                       (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_LAZY), (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_BOTH)):
                a = gi.query_chunks(bases, offs, lens, k, mode, out, strands, streaming)
                b = gi.query_chunks(bases, offs, lens, k, mode, out, strands, streaming, packed=True)
                assert np.array_equal(a, b), (max_kmers, streaming, mode, out, strands)
            by = gi.query_chunks(bases, offs, lens, k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, streaming)
            for packed in (False, True):
                bi = gi.query_chunks(bases, offs, lens, k, fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_LAZY, streaming, packed=packed)
                assert np.array_equal(bi, np.packbits(by, bitorder="little")), (max_kmers, streaming, packed)
    with pytest.raises(fg.FmsiGpuError):  # one bit cannot hold two strands
        gi.query_kmers(kmers[:10], k, fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_BOTH)
    gi.close()


@pytest.mark.parametrize("tier", ["auto", "backward", "backward_t0", "backward_ms0", "loc", "loc_backward"])
def test_reads_api_matches_chunk_calls(tier):
    """fmsi_gpu_query_reads_packed: the caller hands over whole reads (offsets into one 2-bit text), the device cuts them
    into the chunks its kernels take. Reads of every awkward length (empty, shorter than k, exactly k, 64 / 65 / 129
    k-mers, thousands of bases) must give what the chunk calls give — those are the ones checked against the oracle —
    streamed and single, every output, and through the pipelined host path (> 16 MiB of text)."""
    d = os.path.join(GOLDEN, "syn_k31_min")
    k = 31
    prefix = os.path.join(d, "ms.fa")
    gi = fg.Index.load(prefix, use_klcp=True, **TIER_KW[tier])
    ms_codes = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    rng = np.random.default_rng(31)
    lens = [0, 5, k - 1, k, k + 1, k + 63, k + 64, k + 128, 150, 150, 151, 3000, 1, 200, k + 62]
    seqs = []
    for ln in lens * 3:
        if ln <= len(ms_codes) and rng.random() < 0.7:
            p0 = int(rng.integers(0, len(ms_codes) - ln + 1))
            c = ms_codes[p0:p0 + ln].copy()
            if ln and rng.random() < 0.5:
                c = synth.revcomp_codes(c)
        else:
            c = rng.integers(0, 4, size=ln).astype(np.uint8)
        seqs.append(c.astype(np.uint8))
    reads = [synth.codes_to_ascii(c) for c in seqs]
    bases, offs, clens = _chunks_of([c for c in seqs if len(c) >= k], k, 64)
    for streaming in (False, True):
        for mode, out, strands in ((fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY), (fg.MODE_OR, fg.OUT_PRESENCE, fg.STRANDS_BOTH),
                                   (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_LAZY), (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_BOTH), (fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_LAZY)):
            a = gi.query_reads(reads, k, mode, out, strands, streaming)
            b = gi.query_chunks(bases, offs, clens, k, mode, out, strands, streaming)
            assert np.array_equal(a, b), (tier, streaming, mode, out, strands)
    # pipelined host path: 130 000 reads of 150 bp (19.5 M bases)
    big = synth.read_queries(ms_codes, 150, 130_000, 21)
    big_reads = [synth.codes_to_ascii(r) for r in big]
    bases, offs, clens = _chunks_of(list(big), k, 64)
    for streaming, out in ((True, fg.OUT_PRESENCE), (False, fg.OUT_PRESENCE_BITS), (True, fg.OUT_ORDERS)):
        assert np.array_equal(gi.query_reads(big_reads, k, fg.MODE_ALL if out != fg.OUT_ORDERS else fg.MODE_OR, out, fg.STRANDS_LAZY, streaming),
                              gi.query_chunks(bases, offs, clens, k, fg.MODE_ALL if out != fg.OUT_ORDERS else fg.MODE_OR, out, fg.STRANDS_LAZY, streaming)), (tier, "big", streaming, out)
    # pipelined again with 121 k-mers per read: the result spans of the text pieces then end in the middle of a byte of
    # the bit-packed output, which is packed and returned span by span
    odd = synth.read_queries(ms_codes, 151, 120_000, 22)
    odd_reads = [synth.codes_to_ascii(r) for r in odd]
    bases, offs, clens = _chunks_of(list(odd), k, 64)
    assert len(bases) > (16 << 20)
    by = gi.query_chunks(bases, offs, clens, k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, True)
    assert np.array_equal(gi.query_reads(odd_reads, k, fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_LAZY, True), np.packbits(by, bitorder="little")), (tier, "odd reads")
    assert np.array_equal(gi.query_chunks(bases, offs, clens, k, fg.MODE_ALL, fg.OUT_PRESENCE_BITS, fg.STRANDS_LAZY, False, packed=True),
                          np.packbits(by, bitorder="little")), (tier, "odd chunks")
    # malformed calls
    L_ = fg.lib()
    words = fg.pack_text(np.frombuffer(b"ACGT" * 50, dtype=np.uint8))
    res = np.zeros(1000, np.uint8)
    for off, n_res in (([0, 100, 50, 200], 140 - 0), ([0, 100, 300], 240), ([0, 100, 200], 7)):
        o = np.array(off, dtype=np.uint64)
        rc = L_.fmsi_gpu_query_reads_packed(gi._h, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, 0, words.ctypes.data, 200, o.ctypes.data, len(off) - 1, n_res, k,
                                            res.ctypes.data, fg.MEM_HOST, None)
        assert rc == -1, (off, n_res)
    gi.close()


def test_error_behaviour():
    with pytest.raises(fg.FmsiGpuError) as e:
        fg.Index.load("/nonexistent/prefix")
    assert e.value.code == -2  # "index not correctly loaded"
    d = os.path.join(GOLDEN, "syn_k31_noklcp")
    gi = fg.Index.load(os.path.join(d, "ms.fa"), use_klcp=True)  # no .klcp file on disk
    assert not gi.has_klcp
    with pytest.raises(fg.FmsiGpuError) as e:
        gi.query_chunks(b"A" * 40, [0], [40], streaming=True)
    assert e.value.code == -4  # kLCP mismatch (reference main.cpp:309-312)
    with pytest.raises(fg.FmsiGpuError):
        gi.query_kmers([0], k=33)
    assert gi.query_kmers(np.zeros(0, np.uint64)).size == 0  # empty batch
    gi.close()
    d = os.path.join(GOLDEN, "syn_k31_max")
    gi = fg.Index.load(os.path.join(d, "ms.fa"), use_klcp=True, dict=0)
    with pytest.raises(fg.FmsiGpuError) as e:  # the kLCP array is the index's k's: streaming with another k is refused
        gi.query_chunks(b"A" * 40, [0], [40], k=21, streaming=True)
    assert e.value.code == -5
    assert gi.query_chunks(b"A" * 40, [0], [40], k=21, streaming=False).shape == (20,)
    for streaming in (False, True):
        with pytest.raises(fg.FmsiGpuError) as e:  # a chunk that runs past the text
            gi.query_chunks(b"A" * 40, [0, 20], [40, 40], k=31, streaming=streaming)
        assert e.value.code == -1
    gi.close()


# ---- mid-size index built on the box with the reference binary -------------------------------------
@pytest.fixture(scope="module")
def midsize(tmp_path_factory):
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/fmsi not shipped")
    d = tmp_path_factory.mktemp("mid")
    g = synth.random_codes(1_000_000, 2024)
    ms = synth.contig_superstring(g, 31, 200, 2025, "max")
    fa = str(d / "ms.fa")
    synth.write_fasta_single(fa, "ms", ms)
    subprocess.run([REF_EXE, "index", "-k", "31", fa], check=True, capture_output=True)
    return g, fa


def test_midsize_parity_and_properties(midsize):
    g, fa = midsize
    k = 31
    gi = fg.Index.load(fa, use_klcp=True)
    oi = OracleIndex.load(fa, use_klcp=True)
    rows = synth.kmer_queries(g, k, 200_000, 5)
    kmers = synth.pack_rows(rows)
    for mode, out, omode, oord in ((fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False), (fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False),
                                   (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True)):
        got = gi.query_kmers(kmers, k, mode, out)
        assert np.array_equal(got.astype(np.int64), oi.query_packed(kmers, k, omode, oord))
    # size-independent properties at a larger batch (2M k-mers): strand symmetry, idempotence,
    # every genome k-mer present, lookup ids form an injection into [0, mask_ones)
    big = synth.pack_rows(synth.kmer_queries(g, k, 2_000_000, 6))
    r1 = gi.query_kmers(big, k, fg.MODE_ALL)
    r2 = gi.query_kmers(synth.revcomp_packed(big, k), k, fg.MODE_ALL)
    assert np.array_equal(r1, r2) and np.array_equal(r1, gi.query_kmers(big, k, fg.MODE_ALL))
    assert np.array_equal(r1, gi.query_kmers(big, k, fg.MODE_OR))  # max-ones mask: -O == or
    gk = synth.pack_kmers(g[:300_000], k)
    assert gi.query_kmers(gk, k, fg.MODE_ALL).all()
    ids = gi.query_kmers(gk, k, output=fg.OUT_ORDERS)
    assert (ids >= 0).all() and ids.max() < gi.mask_ones
    canon = synth.canonical_packed(gk, k)
    _, first = np.unique(canon, return_index=True)
    assert len(np.unique(ids[first])) == len(first)  # distinct canonical k-mers -> distinct ids
    # streaming reads == single queries
    reads = synth.read_queries(g, 150, 3000, 8)
    bases, offs, lens = _chunks_of(list(reads), k, 64)
    a = gi.query_chunks(bases, offs, lens, k, fg.MODE_ALL, streaming=True)
    b = gi.query_chunks(bases, offs, lens, k, fg.MODE_ALL, streaming=False)
    assert np.array_equal(a, b)
    la = gi.query_chunks(bases, offs, lens, k, output=fg.OUT_ORDERS, streaming=True)
    lb = gi.query_chunks(bases, offs, lens, k, output=fg.OUT_ORDERS, streaming=False)
    assert np.array_equal(la, lb)
    allk = np.concatenate([synth.pack_kmers(synth.ascii_to_codes(bases[o:o + l]), k) for o, l in zip(offs.tolist()[:400], lens.tolist()[:400])])
    assert np.array_equal(a[:len(allk)].astype(np.int64), oi.query_packed(allk, k, MODE_ALL, False))
    gi.close()
    oi.close()


@pytest.mark.parametrize("tier", ["backward", "fold"])
def test_large_host_chunk_calls_are_pipelined_and_identical(midsize, tier):
    """A host-mode fmsi_gpu_query_chunks call with more than 16 MiB of text is cut into text pieces whose H2D copies
    overlap the queries and result copies of the chunks already complete (query_chunks_impl). Its results must
    equal those of small (single-batch) calls over the same chunks, for streamed and single queries, presence and
    ids, both strand policies; gaps in the text between chunks and a tail after the last chunk included."""
    g, fa = midsize
    k = 31
    gi = fg.Index.load(fa, use_klcp=True, dict=2 if tier == "fold" else 0)
    reads = synth.read_queries(g, 150, 130_000, 21)  # 19.5 M bases
    bases, offs, lens = _chunks_of(list(reads), k, 64)
    bases = bases + b"ACGT" * 1000  # text after the last chunk
    assert len(bases) > (16 << 20)
    oi = OracleIndex.load(fa, use_klcp=True)
    sample = np.concatenate([synth.pack_kmers(synth.ascii_to_codes(bases[o:o + l]), k) for o, l in zip(offs.tolist()[-300:], lens.tolist()[-300:])])
    for mode, out, strands, streaming in ((fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, True), (fg.MODE_OR, fg.OUT_PRESENCE, fg.STRANDS_BOTH, False),
                                          (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_LAZY, True), (fg.MODE_OR, fg.OUT_ORDERS, fg.STRANDS_BOTH, False)):
        big = gi.query_chunks(bases, offs, lens, k, mode, out, strands, streaming)
        parts = []
        step = 40_000  # chunks per small call: ~3.4 MiB of text each, below the pipelining threshold
        for c in range(0, len(offs), step):
            o, l = offs[c:c + step], lens[c:c + step]
            lo, hi = int(o[0]), int(o[-1] + l[-1])
            parts.append(gi.query_chunks(bases[lo:hi], o - np.uint64(lo), l, k, mode, out, strands, streaming))
        assert np.array_equal(big, np.concatenate(parts)), (tier, mode, out, strands, streaming)
        assert np.array_equal(big, gi.query_chunks(bases, offs, lens, k, mode, out, strands, streaming, packed=True)), (tier, "packed text")
        if out == fg.OUT_PRESENCE and strands == fg.STRANDS_LAZY:
            assert np.array_equal(big[-len(sample):].astype(np.int64), oi.query_packed(sample, k, MODE_ALL, False))
            for packed in (False, True):
                bits = gi.query_chunks(bases, offs, lens, k, mode, fg.OUT_PRESENCE_BITS, strands, streaming, packed=packed)
                assert np.array_equal(bits, np.packbits(big, bitorder="little")), (tier, "bits", packed)
    gi.close()
    oi.close()


def test_large_host_chunk_calls_with_long_k():
    """The pipelined host path for k > 32 (queries are start positions in the packed text): 17 MiB of text against the
    k = 47 golden index, chunks that embed the superstring's own k-mers every so often; one big call == small calls."""
    d = os.path.join(GOLDEN, "syn_k47_max")
    k = json.load(open(os.path.join(d, "meta.json")))["k"]
    gi = fg.Index.load(os.path.join(d, "ms.fa"), use_klcp=True)
    ms = open(os.path.join(d, "ms.fa"), "rb").read().split(b"\n")[1].upper()
    rng = np.random.default_rng(3)
    L, R = 200, 90_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L * R, dtype=np.uint8)].copy()
    msa = np.frombuffer(ms, dtype=np.uint8)
    for r in range(0, R, 50):  # every 50th chunk carries a piece of the superstring: present k-mers
        p = int(rng.integers(0, len(msa) - 120))
        text[r * L + 30:r * L + 150] = msa[p:p + 120]
    bases = text.tobytes()
    assert len(bases) > (16 << 20)
    offs = np.arange(R, dtype=np.uint64) * L
    lens = np.full(R, L, dtype=np.uint32)
    for out, strands in ((fg.OUT_PRESENCE, fg.STRANDS_BOTH), (fg.OUT_ORDERS, fg.STRANDS_LAZY)):
        big = gi.query_chunks(bases, offs, lens, k, fg.MODE_OR, out, strands, True)
        parts = []
        step = 15_000  # 3 MB of text per call: single batch
        for c in range(0, R, step):
            o, l = offs[c:c + step], lens[c:c + step]
            lo, hi = int(o[0]), int(o[-1] + l[-1])
            parts.append(gi.query_chunks(bases[lo:hi], o - np.uint64(lo), l, k, fg.MODE_OR, out, strands, False))
        small = np.concatenate(parts)
        assert np.array_equal(big, small)
        hits = (big.reshape(-1, 2)[:, 0] if out == fg.OUT_ORDERS and strands == fg.STRANDS_BOTH else big)
        assert (hits >= 0).any() if out == fg.OUT_ORDERS else big.any()  # the embedded superstring pieces are found
    gi.close()
