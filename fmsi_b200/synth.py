"""Seeded synthetic workloads of the shapes BASELINE.json names.

Workload generators only — nothing here is on the query path. Everything is numpy so that the
same bytes are produced in the authoring container and on the GPU box.

* genomes: i.i.d. uniform ACGT
* masked superstrings (mask-cased: upper = ON, lower = OFF, last k-1 letters lower), either the
  genome itself (every k-mer ON: a valid max-ones superstring) or shuffled / reverse-complemented
  contigs glued together with the max-ones mask recomputed against the genome's canonical k-mer set
  (so OFF occurrences and both-strand occurrences exist)
* query sets: single k-mers (50 % present, random strand) and 150 bp reads with 1 % substitutions
"""
from __future__ import annotations

import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_acgt_lower = np.frombuffer(b"acgt", dtype=np.uint8)
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _ch in enumerate(b"ACGT"):
    _CODE[_ch] = _i
    _CODE[_ch + 32] = _i


def random_codes(n: int, seed: int) -> np.ndarray:
    """n i.i.d. uniform base codes (A=0 C=1 G=2 T=3)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def codes_to_ascii(codes: np.ndarray, mask: np.ndarray | None = None) -> bytes:
    if mask is None:
        return ACGT[codes].tobytes()
    return np.where(mask.astype(bool), ACGT[codes], _acgt_lower[codes]).tobytes()


def ascii_to_codes(s: bytes) -> np.ndarray:
    return _CODE[np.frombuffer(s, dtype=np.uint8)]


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    return (3 - codes[::-1]).astype(np.uint8)


def pack_kmers(codes: np.ndarray, k: int) -> np.ndarray:
    """All len-k+1 k-mers of a code array, packed 2 bits/base, first base in the highest used bits
    (the layout of include/fmsi_gpu.h). k <= 32."""
    assert 1 <= k <= 32
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    c64 = codes.astype(np.uint64)
    for t in range(k):
        out <<= np.uint64(2)
        out |= c64[t:t + n]
    return out


def revcomp_packed(kmers: np.ndarray, k: int) -> np.ndarray:
    """Reverse complement of packed k-mers."""
    x = ~kmers.astype(np.uint64)
    # reverse the 32 2-bit groups of a 64-bit word
    x = ((x >> np.uint64(2)) & np.uint64(0x3333333333333333)) | ((x & np.uint64(0x3333333333333333)) << np.uint64(2))
    x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4))
    x = x.byteswap()
    return x >> np.uint64(64 - 2 * k)


def canonical_packed(kmers: np.ndarray, k: int) -> np.ndarray:
    return np.minimum(kmers, revcomp_packed(kmers, k))


def genome_superstring(codes: np.ndarray, k: int) -> bytes:
    """The sequence itself, upper-case except the last k-1 letters: every k-mer occurrence ON."""
    mask = np.ones(len(codes), dtype=np.uint8)
    mask[len(codes) - (k - 1):] = 0
    return codes_to_ascii(codes, mask)


def random_masked_superstring(n: int, seed: int, k: int, off: float) -> bytes:
    """n i.i.d. bases (seed) with a random mask: every position OFF with probability `off` (seed + 1), the last k-1
    OFF. Platform-independent (numpy PCG64), so that hashes of the reference's index of it can be committed
    (tests/golden/make_ref_index_hashes.py) and the input regenerated anywhere."""
    codes = random_codes(n, seed)
    rng = np.random.default_rng(seed + 1)
    mask = rng.random(n) >= off
    mask[n - (k - 1):] = False
    return codes_to_ascii(codes, mask.astype(np.uint8))


def contig_superstring(codes: np.ndarray, k: int, n_pieces: int, seed: int,
                       ones: str = "max") -> bytes:
    """Cut the genome into n_pieces contigs (consecutive contigs overlap by k-1 so no k-mer is
    lost), shuffle them, reverse-complement half, concatenate, and mask.

    ones="max": ON at every position whose canonical k-mer belongs to the genome's k-mer set (what
    `kmercamel optimize -a ones -c` yields for this superstring). ones="min": exactly one ON
    occurrence per canonical k-mer (the first), a valid but non-max-ones mask.
    """
    rng = np.random.default_rng(seed)
    n = len(codes)
    cuts = np.sort(rng.choice(np.arange(k, n - k), size=n_pieces - 1, replace=False)) if n_pieces > 1 else np.array([], dtype=np.int64)
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts + (k - 1), [n]])
    order = rng.permutation(n_pieces)
    flip = rng.integers(0, 2, size=n_pieces).astype(bool)
    parts = []
    for p in order:
        piece = codes[starts[p]:min(ends[p], n)]
        parts.append(revcomp_codes(piece) if flip[p] else piece)
    s = np.concatenate(parts)
    if k <= 32:
        kset = np.unique(canonical_packed(pack_kmers(codes, k), k))
        sk = canonical_packed(pack_kmers(s, k), k)
    else:  # k-mers no longer fit a word: canonical k-mers as byte strings -> dense ids (small inputs only)
        ids: dict[bytes, int] = {}

        def canon_ids(c: np.ndarray) -> np.ndarray:
            win = np.lib.stride_tricks.sliding_window_view(c.astype(np.uint8), k)
            out = np.empty(len(win), dtype=np.uint64)
            for r in range(len(win)):
                f = win[r].tobytes()
                b = (3 - win[r][::-1]).astype(np.uint8).tobytes()
                out[r] = ids.setdefault(min(f, b), len(ids))
            return out

        kset = np.unique(canon_ids(codes))
        sk = canon_ids(s)
    pos = np.searchsorted(kset, sk)
    pos[pos >= len(kset)] = len(kset) - 1
    member = kset[pos] == sk
    mask = np.zeros(len(s), dtype=np.uint8)
    if ones == "max":
        mask[:len(sk)] = member
    else:
        first = np.zeros(len(sk), dtype=bool)
        idx = np.flatnonzero(member)
        _, first_idx = np.unique(sk[idx], return_index=True)
        first[idx[first_idx]] = True
        mask[:len(sk)] = first
    return codes_to_ascii(s, mask)


def write_fasta_single(path: str, name: str, seq: bytes) -> None:
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        f.write(seq)
        f.write(b"\n")


def kmer_queries(codes: np.ndarray, k: int, n_queries: int, seed: int, frac_present: float = 0.5) -> np.ndarray:
    """(n_queries, k) uint8 code matrix: present k-mers (uniform genome position, random strand)
    interleaved with i.i.d. random k-mers."""
    rng = np.random.default_rng(seed)
    present = rng.random(n_queries) < frac_present
    pos = rng.integers(0, len(codes) - k + 1, size=n_queries)
    win = codes[pos[:, None] + np.arange(k)[None, :]]
    strand = rng.integers(0, 2, size=n_queries).astype(bool)
    win[strand] = 3 - win[strand][:, ::-1]
    rnd = rng.integers(0, 4, size=(n_queries, k), dtype=np.uint8)
    return np.where(present[:, None], win, rnd).astype(np.uint8)


def pack_rows(rows: np.ndarray) -> np.ndarray:
    """(n, k) code matrix -> n packed k-mers."""
    n, k = rows.shape
    out = np.zeros(n, dtype=np.uint64)
    for t in range(k):
        out <<= np.uint64(2)
        out |= rows[:, t].astype(np.uint64)
    return out


def rows_to_fasta(rows: np.ndarray, prefix: str = "q") -> bytes:
    """One FASTA record per row: >q<i>\\nSEQ\\n."""
    n, k = rows.shape
    seqs = ACGT[rows]
    out = bytearray()
    for i in range(n):
        out += b">" + prefix.encode() + str(i).encode() + b"\n" + seqs[i].tobytes() + b"\n"
    return bytes(out)


def read_queries(codes: np.ndarray, read_len: int, n_reads: int, seed: int, sub_rate: float = 0.01) -> np.ndarray:
    """(n_reads, read_len) code matrix: uniform positions, random strand, each base replaced by a
    different base with probability sub_rate."""
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, len(codes) - read_len + 1, size=n_reads)
    win = codes[pos[:, None] + np.arange(read_len)[None, :]]
    strand = rng.integers(0, 2, size=n_reads).astype(bool)
    win[strand] = 3 - win[strand][:, ::-1]
    sub = rng.random(win.shape) < sub_rate
    shift = rng.integers(1, 4, size=win.shape, dtype=np.uint8)
    return np.where(sub, (win + shift) & 3, win).astype(np.uint8)


def ensure_dir(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path


def packed_kmer_queries(genome_kmers: np.ndarray, k: int, n_queries: int, seed: int, frac_present: float = 0.5) -> np.ndarray:
    """Packed form of kmer_queries() for large batches: genome_kmers = pack_kmers(genome, k).
    Present k-mers are drawn from uniform genome positions with a random strand, absent ones are
    i.i.d. uniform k-mers; the two kinds are interleaved at random."""
    rng = np.random.default_rng(seed)
    present = rng.random(n_queries) < frac_present
    pos = rng.integers(0, len(genome_kmers), size=n_queries)
    km = genome_kmers[pos]
    flip = rng.integers(0, 2, size=n_queries).astype(bool)
    km = np.where(flip, revcomp_packed(km, k), km)
    hi = np.uint64((1 << (2 * k)) - 1)
    rnd = rng.integers(0, np.iinfo(np.uint64).max, size=n_queries, dtype=np.uint64, endpoint=True) & hi
    return np.where(present, km, rnd).astype(np.uint64)


def packed_to_fasta(kmers: np.ndarray, k: int) -> bytes:
    """Fixed-width FASTA: one record `>q\\nSEQ\\n` per packed k-mer (vectorised)."""
    n = len(kmers)
    rec = np.empty((n, k + 4), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("q")
    rec[:, 2] = ord("\n")
    for t in range(k):
        rec[:, 3 + t] = ACGT[((kmers >> np.uint64(2 * (k - 1 - t))) & np.uint64(3)).astype(np.uint8)]
    rec[:, k + 3] = ord("\n")
    return rec.tobytes()
