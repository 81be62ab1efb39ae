"""Multi-GPU host logic on CPU: the shard plan and the world_size-2 gather path (gloo). The device
work is replaced by the oracle here (tests may use it as a stand-in; the product path is the GPU
replica on each rank — see bench.py and tests/test_gpu_parity.py)."""
import json
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

from fmsi_b200 import shard, synth


def test_plan_kmers_covers_everything_once():
    for n in (0, 1, 7, 64, 1000, 12345):
        for w in (1, 2, 3, 4, 8):
            p = shard.plan_kmers(n, w)
            assert len(p) == w and p[0][0] == 0 and p[-1][1] == n
            assert all(p[r][1] == p[r + 1][0] for r in range(w - 1))
            sizes = [e - b for b, e in p]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.plan_kmers(5, 0)


def test_plan_chunks_balances_kmers_and_never_splits_a_chunk():
    rng = np.random.default_rng(5)
    k = 31
    for w in (1, 2, 4, 8):
        lens = rng.integers(k, 400, size=1000)
        p = shard.plan_chunks(lens, k, w)
        assert p[0][0] == 0 and p[-1][1] == len(lens)
        assert all(p[r][1] == p[r + 1][0] for r in range(w - 1))
        per = [int((lens[b:e] - (k - 1)).sum()) for b, e in p]
        assert sum(per) == int((lens - (k - 1)).sum())
        assert max(per) - min(per) <= 2 * int(lens.max())   # balanced to within a chunk or two
    assert shard.plan_chunks([], k, 3) == [(0, 0)] * 3
    assert shard.plan_chunks([40], k, 2)[1] == (shard.plan_chunks([40], k, 2)[0][1], 1)
    with pytest.raises(ValueError):
        shard.plan_chunks([10], k, 2)
    roff = shard.result_offsets([31, 35, 60], k)
    assert roff.tolist() == [0, 1, 6, 36]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle_ffi import MODE_ALL, OracleIndex
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k = json.load(open(os.path.join(case, "meta.json")))["k"]
    prefix = os.path.join(case, "ms.fa")
    oi = OracleIndex.load(prefix, use_klcp=False)     # this rank's "replica"
    ms = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    kmers = synth.pack_rows(synth.kmer_queries(ms, k, 3001, 11))   # same seeded batch on every rank
    # single k-mers: presence (u8) and lookup ids (i64)
    pres = shard.sharded_query_kmers(kmers, lambda q: oi.query_packed(q, k, MODE_ALL, False).astype(np.uint8), rank, world)
    ids = shard.sharded_query_kmers(kmers, lambda q: oi.query_packed(q, k, 0, True), rank, world)
    # chunks: reads cut into ragged chunks; a chunk's k-mers are contiguous windows of the read
    reads = synth.read_queries(ms, 150, 40, 2)
    bases = np.concatenate(reads)
    offs, lens = [], []
    rng = np.random.default_rng(3)
    for r in range(len(reads)):
        cut = int(rng.integers(k, 150 - k))
        offs += [150 * r, 150 * r + cut - (k - 1)]
        lens += [cut, 150 - cut + (k - 1)]
    offs, lens = np.asarray(offs), np.asarray(lens)

    def chunk_compute(co, cl):
        ks = [synth.pack_kmers(bases[o:o + l], k) for o, l in zip(co, cl)]
        q = np.concatenate(ks) if ks else np.zeros(0, dtype=np.uint64)
        return oi.query_packed(q, k, MODE_ALL, False).astype(np.uint8)
    chunked = shard.sharded_query_chunks(offs, lens, k, chunk_compute, rank, world)
    if rank == 0:
        want_p = oi.query_packed(kmers, k, MODE_ALL, False).astype(np.uint8)
        want_i = oi.query_packed(kmers, k, 0, True)
        want_c = chunk_compute(offs, lens)
        ok = bool(np.array_equal(pres, want_p) and np.array_equal(ids, want_i) and np.array_equal(chunked, want_c)
                  and ids.dtype == np.int64 and pres.dtype == np.uint8)
        with open(out_path, "w") as f:
            json.dump({"ok": ok, "n": int(len(kmers)), "n_chunk_results": int(len(chunked))}, f)
    else:
        assert pres is None and ids is None and chunked is None
    oi.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo_shard_and_gather(tmp_path, oracle_built):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.json")
    case = os.path.join(GOLDEN, "syn_k31_max")
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["ok"] and res["n"] == 3001
    assert res["n_chunk_results"] == 40 * (150 - 31 + 1)


def _bench_worker(rank, world, port, case, out_path):
    """bench.py's multi-rank bookkeeping without a device: process group over gloo, max over ranks, the parity sample
    of EVERY rank gathered to and checked on rank 0 (the oracle stands in for the rank's GPU replica)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import argparse
    import bench
    from oracle_ffi import OracleIndex
    cx = bench.Ctx(argparse.Namespace(batch=0, no_parity=False))
    cx.init_dist()
    assert cx.world == world and cx.rank == rank
    assert cx.max_over_ranks(float(rank + 1)) == float(world)
    k = json.load(open(os.path.join(case, "meta.json")))["k"]
    prefix = os.path.join(case, "ms.fa")
    ms = synth.ascii_to_codes(open(prefix, "rb").read().split(b"\n")[1])
    wl = dict(name="t", prefix=prefix, k=k, genome=ms, codes=None)
    replica = OracleIndex.load(prefix, use_klcp=False)

    class Replica:  # what bench.py calls on the rank's index
        def __init__(self, wrong):
            self.wrong = wrong

        def query_kmers(self, q, kk, mode, output):
            r = replica.query_packed(q, kk, 1 if mode == 1 else 0, output == 1)
            if self.wrong and len(r):
                r = r.copy()
                r[0] ^= 1
            return r

    oi = replica if rank == 0 else None
    ok = bench.parity_kmers(cx, Replica(False), oi, wl, 500, 7, 1, 0, 1, False)
    bad = bench.parity_kmers(cx, Replica(rank == world - 1), oi, wl, 500, 8, 1, 0, 1, False)  # only the LAST rank answers wrongly
    reads = synth.read_queries(ms, 150, 20, 5 + rank)
    allk = np.concatenate([synth.pack_kmers(r, k) for r in reads])
    res = (replica.query_packed(allk, k, 1, False) == 1).astype(np.uint8)
    rok = bench.parity_reads(cx, oi, wl, dict(reads=reads, results=res), 1, False)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"ok": ok, "bad": bad, "reads": rok}, f)
    else:
        assert ok is None and bad is None and rok is None
    cx.barrier()
    cx.dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_bench_parity_bookkeeping(tmp_path, oracle_built):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.json")
    mp.spawn(_bench_worker, args=(2, _free_port(), os.path.join(GOLDEN, "syn_k31_max"), out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res == {"ok": True, "bad": False, "reads": True}  # a wrong answer on rank 1 is caught on rank 0
