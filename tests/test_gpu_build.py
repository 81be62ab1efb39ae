"""GPU index construction (fmsi_gpu_index_build / _save) must reproduce the reference's
`fmsi index` output byte for byte: the suffix array is unique, so ac_gt/ac/gt/klcp/mask/misc are."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
from oracle_ffi import REF_EXE

import fmsi_b200 as fg
from fmsi_b200 import synth

pytestmark = pytest.mark.gpu
EXTS = ["ac_gt", "ac", "gt", "mask", "klcp", "misc"]


def compare_files(got_prefix, want_prefix, klcp=True):
    for ext in EXTS:
        if ext == "klcp" and not klcp:
            assert not os.path.exists(f"{got_prefix}.fmsi.klcp")
            continue
        a = open(f"{got_prefix}.fmsi.{ext}", "rb").read()
        b = open(f"{want_prefix}.fmsi.{ext}", "rb").read()
        assert a == b, f"{ext} differs ({len(a)} vs {len(b)} bytes)"


@pytest.mark.parametrize("case", golden_cases())
def test_build_reproduces_reference_index_files(case, tmp_path):
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "meta.json")))
    ms = open(os.path.join(d, "ms.fa"), "rb").read().split(b"\n")[1]
    idx = fg.Index.build(ms, meta["k"], with_klcp=meta["klcp"])
    out = str(tmp_path / "ms.fa")
    idx.save(out)
    compare_files(out, os.path.join(d, "ms.fa"), klcp=meta["klcp"])
    # and the built index answers queries like the loaded one
    ref = fg.Index.load(os.path.join(d, "ms.fa"), use_klcp=meta["klcp"])
    codes = synth.ascii_to_codes(ms)
    if meta["k"] > 32:  # no packed form: the k-mers of the superstring and of random text, as chunks
        k = meta["k"]
        text = synth.codes_to_ascii(codes[:6000]) + synth.codes_to_ascii(np.random.default_rng(1).integers(0, 4, size=3000, dtype=np.uint8))
        offs, lens = [0, 6000], [6000, 3000]
        for out_kind in (fg.OUT_PRESENCE, fg.OUT_ORDERS):
            assert np.array_equal(idx.query_chunks(text, offs, lens, k, fg.MODE_OR, out_kind), ref.query_chunks(text, offs, lens, k, fg.MODE_OR, out_kind))
        if meta["klcp"]:
            a = idx.query_chunks(text, offs, lens, k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_BOTH, True)
            assert np.array_equal(a, ref.query_chunks(text, offs, lens, k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_BOTH, False))
    elif len(codes) >= meta["k"]:
        kmers = synth.pack_kmers(codes, meta["k"])[:5000]
        rnd = np.random.default_rng(1).integers(0, 1 << (2 * meta["k"]) - 1, size=2000, dtype=np.uint64)
        q = np.concatenate([kmers, rnd])
        for out_kind in (fg.OUT_PRESENCE, fg.OUT_ORDERS):
            assert np.array_equal(idx.query_kmers(q, meta["k"], fg.MODE_OR, out_kind), ref.query_kmers(q, meta["k"], fg.MODE_OR, out_kind))
    idx.close()
    ref.close()


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/fmsi not shipped")
@pytest.mark.parametrize("name", ["polyA", "tandem", "repeats", "random300k", "random2M", "two_letters"])
def test_build_on_repetitive_inputs(name, tmp_path):
    rng = np.random.default_rng(5)
    if name == "polyA":
        ms, k = b"A" * 700 + b"a" * 4, 5
    elif name == "tandem":
        ms, k = (b"ACGTTGCA" * 300) + b"acgt", 5
    elif name == "repeats":
        unit = synth.codes_to_ascii(rng.integers(0, 4, size=900, dtype=np.uint8))
        parts = []
        for r in range(12):
            u = bytearray(unit)
            for _ in range(3):
                u[int(rng.integers(0, len(u)))] = b"ACGT"[int(rng.integers(0, 4))]
            parts.append(bytes(u))
        s = b"".join(parts)
        ms, k = s[:-30] + s[-30:].lower(), 31
    elif name in ("random300k", "random2M"):  # several granules / thread ranges of the save path's ac / gt compaction
        g = synth.random_codes(300_000 if name == "random300k" else 2_000_000, 77)
        ms, k = synth.contig_superstring(g, 23, 50, 78, "max"), 23
    else:
        s = synth.codes_to_ascii(rng.integers(0, 2, size=5000, dtype=np.uint8) * 3)  # only A and T
        ms, k = s[:-8] + s[-8:].lower(), 9
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta_single(fa, "ms", ms)
    subprocess.run([REF_EXE, "index", "-k", str(k), fa], check=True, capture_output=True)
    idx = fg.Index.build(ms, k, with_klcp=True)
    out = str(tmp_path / "gpu.fa")
    idx.save(out)
    compare_files(out, fa)
    idx.close()


HASHES = os.path.join(GOLDEN, "ref_index_hashes.json")


@pytest.mark.skipif(not os.path.exists(HASHES), reason="tests/golden/ref_index_hashes.json not generated")
@pytest.mark.parametrize("name", sorted(json.load(open(HASHES))) if os.path.exists(HASHES) else [])
def test_build_matches_reference_hashes_at_scale(name, tmp_path):
    """Byte identity with the reference's `fmsi index` at 100 Mbp (and 20 Mbp without kLCP, half the mask OFF): the
    reference was run once on the seeded input where it compiles (tests/golden/make_ref_index_hashes.py: 110 s,
    1.6 GB) and the SHA-256 of each of its files committed; the GPU builder must reproduce every one of them."""
    import hashlib
    case = json.load(open(HASHES))[name]
    ms = synth.random_masked_superstring(case["n"], case["seed"], case["k"], case["off"])
    h = hashlib.sha256(b">ms\n" + ms + b"\n").hexdigest()
    assert h == case["input_sha256"], "the seeded input is not the one the reference indexed"
    idx = fg.Index.build(ms, case["k"], with_klcp=case["klcp"], dict=0, multistep=0, prefix_t=0)
    out = str(tmp_path / "ms.fa")
    idx.save(out)
    for ext, want in case["files"].items():
        got = hashlib.sha256(open(f"{out}.fmsi.{ext}", "rb").read()).hexdigest()
        assert got == want, f"{name}: .fmsi.{ext} differs from the reference's file"
        assert os.path.getsize(f"{out}.fmsi.{ext}") == case["sizes"][ext]
    assert os.path.exists(f"{out}.fmsi.klcp") == case["klcp"]
    idx.close()


@pytest.mark.usefixtures("oracle_built")
@pytest.mark.skipif(not os.path.exists(HASHES), reason="tests/golden/ref_index_hashes.json not generated")
def test_wide_layout_at_100mbp(tmp_path):
    """The wide (N >= 2^32) device layout — 64-bit positions, per-superblock counter bases, 16-byte table entries — on a
    100 Mbp index (the reference-hash input above), forced through the `sb_shift_log2` hook with superblocks of 2^10 and
    2^16 blocks: every mode and strand policy, single k-mers and streamed reads, must equal the narrow layout on 1 M
    queries and the oracle on a sample. (A BWT beyond 2^32 rows itself is out of the GPU builder's reach.)"""
    from oracle_ffi import MODE_ALL, MODE_OR, OracleIndex
    case = json.load(open(HASHES))["iid_100m_k31"]
    k = case["k"]
    ms = synth.random_masked_superstring(case["n"], case["seed"], k, case["off"])
    built = fg.Index.build(ms, k, with_klcp=True, dict=0, multistep=0, prefix_t=0)
    prefix = str(tmp_path / "ms.fa")
    built.save(prefix)
    built.close()
    codes = synth.ascii_to_codes(ms[:3_000_000])
    kmers = synth.pack_rows(synth.kmer_queries(codes, k, 1_000_000, 11))
    reads = list(synth.read_queries(codes, 150, 3000, 12))
    bases = b"".join(synth.codes_to_ascii(r) for r in reads)
    offs = np.repeat(np.arange(len(reads), dtype=np.uint64) * 150, 2) + np.tile(np.array([0, 64], dtype=np.uint64), len(reads))
    lens = np.tile(np.array([64 + k - 1, 150 - 64], dtype=np.uint32), len(reads))
    narrow = fg.Index.load(prefix, use_klcp=True, dict=0)
    assert not narrow.wide and narrow.multistep == 2
    oi = OracleIndex.load(prefix, use_klcp=True)
    sample = kmers[:20_000]
    combos = ((fg.MODE_ALL, fg.OUT_PRESENCE, MODE_ALL, False), (fg.MODE_OR, fg.OUT_PRESENCE, MODE_OR, False), (fg.MODE_OR, fg.OUT_ORDERS, MODE_OR, True))
    ref = {}
    for mode, out, omode, oord in combos:
        for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
            ref[(mode, out, strands)] = narrow.query_kmers(kmers, k, mode, out, strands)
            ref[(mode, out, strands, "S")] = narrow.query_chunks(bases, offs, lens, k, mode, out, strands, True)
        assert np.array_equal(ref[(mode, out, fg.STRANDS_LAZY)][:len(sample)].astype(np.int64), oi.query_packed(sample, k, omode, oord))
    narrow.close()
    for shift, ms in ((10, -1), (16, 3), (12, 0)):  # multi-step sectors of the wide layout: auto (2 bases per probe), 3, off
        wide = fg.Index.load(prefix, use_klcp=True, sb_shift_log2=shift, multistep=ms)
        assert wide.wide and not wide.dict and wide.multistep == {-1: 2}.get(ms, ms)
        for mode, out, omode, oord in combos:
            for strands in (fg.STRANDS_LAZY, fg.STRANDS_BOTH):
                assert np.array_equal(wide.query_kmers(kmers, k, mode, out, strands), ref[(mode, out, strands)]), (shift, mode, out, strands)
                assert np.array_equal(wide.query_chunks(bases, offs, lens, k, mode, out, strands, True), ref[(mode, out, strands, "S")]), (shift, "S")
        ii = np.random.default_rng(shift).integers(0, wide.n + 1, size=2000).astype(np.uint64)
        cc = np.random.default_rng(shift + 1).integers(0, 4, size=2000).astype(np.uint8)
        assert wide.rank(ii, cc).tolist() == [oi.rank(i, c) for i, c in zip(ii, cc)]
        wide.close()
    oi.close()


def test_build_rejects_bad_input():
    with pytest.raises(fg.FmsiGpuError):
        fg.Index.build(b"ACGTNACGT", 3)
    with pytest.raises(fg.FmsiGpuError):
        fg.Index.build(b"ACGTACGT", 0)
    with pytest.raises(fg.FmsiGpuError):
        fg.Index.build(b"ACGTACGT", 65537)
