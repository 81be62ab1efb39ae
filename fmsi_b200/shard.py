"""Query sharding across the GPUs of one box.

The query path shards by independent units — single k-mers, or chunks/reads for `-S` — with a full
index replica per GPU and NO data-path collective: rank r answers a contiguous range of the batch
and results return in query order. This module is the host-side plan (pure integer arithmetic, no
device work) plus the torch.distributed driver that applies it; the device work itself is
`Index.query_kmers` / `Index.query_chunks` on the rank's own GPU.

Reference context: the reference is single-threaded (`ms_query`, src/main.cpp:238-375); sharding at
record / chunk boundaries never changes per-k-mer results, only the strand
predictor's history, which the host replays over gathered both-strand results when needed.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def plan_kmers(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [begin, end) ranges of n independent k-mers, sizes differing by at most one."""
    if world < 1 or n < 0:
        raise ValueError("world >= 1 and n >= 0 required")
    base, extra = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def plan_chunks(chunk_len: Sequence[int], k: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [begin, end) CHUNK ranges balanced by k-mer count (chunk c holds chunk_len[c]-k+1
    k-mers). Chunks are never split: a streaming chunk is one unit of work (one lane)."""
    if world < 1:
        raise ValueError("world >= 1 required")
    lens = np.asarray(chunk_len, dtype=np.int64)
    if lens.size and int(lens.min()) < k:
        raise ValueError("every chunk must hold at least one k-mer")
    cum = np.concatenate([[0], np.cumsum(lens - (k - 1))]) if lens.size else np.zeros(1, dtype=np.int64)
    total = int(cum[-1])
    out, b = [], 0
    for r in range(world):
        target = (total * (r + 1)) // world
        e = int(np.searchsorted(cum, target, side="left")) if r + 1 < world else len(lens)
        e = max(e, b)
        out.append((b, min(e, len(lens))))
        b = out[-1][1]
    return out


def result_offsets(chunk_len: Sequence[int], k: int) -> np.ndarray:
    """res_off[c] of fmsi_gpu_query_chunks for back-to-back results; length n_chunks + 1."""
    lens = np.asarray(chunk_len, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lens - (k - 1))]).astype(np.uint64)


def sharded_query_kmers(kmers: np.ndarray, compute: Callable[[np.ndarray], np.ndarray], rank: int, world: int,
                        gather: Callable[[np.ndarray, list[int]], np.ndarray | None] | None = None) -> np.ndarray | None:
    """Rank `rank` answers its range of `kmers` with `compute` (the rank's GPU replica) and the shards
    are concatenated in query order on rank 0 by `gather` (default: torch.distributed.gather over the
    default process group — NCCL on GPUs, gloo in the CPU tests). Returns the full result on rank 0,
    None elsewhere."""
    plan = plan_kmers(len(kmers), world)
    b, e = plan[rank]
    mine = np.ascontiguousarray(compute(kmers[b:e]))
    sizes = [pe - pb for pb, pe in plan]
    if world == 1:
        return mine
    return (gather or dist_gather)(mine, sizes)


def sharded_query_chunks(chunk_off: np.ndarray, chunk_len: np.ndarray, k: int,
                         compute: Callable[[np.ndarray, np.ndarray], np.ndarray], rank: int, world: int,
                         gather: Callable[[np.ndarray, list[int]], np.ndarray | None] | None = None) -> np.ndarray | None:
    """Chunk-granular variant (reads / `-S`): `compute(chunk_off[b:e], chunk_len[b:e])` returns the
    results of those chunks back to back."""
    plan = plan_chunks(chunk_len, k, world)
    roff = result_offsets(chunk_len, k)
    b, e = plan[rank]
    mine = np.ascontiguousarray(compute(chunk_off[b:e], chunk_len[b:e]))
    sizes = [int(roff[pe] - roff[pb]) for pb, pe in plan]
    if len(mine) != sizes[rank]:
        raise RuntimeError("compute returned %d results for a shard of %d k-mers" % (len(mine), sizes[rank]))
    if world == 1:
        return mine
    return (gather or dist_gather)(mine, sizes)


def dist_gather(mine: np.ndarray, sizes: list[int]) -> np.ndarray | None:
    """Gather variable-size shards to rank 0 in rank (= query) order. Bookkeeping only: the query
    path itself needs no collective."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    pad = max(sizes) if sizes else 0
    item = mine.reshape(len(mine), -1)
    width = item.shape[1] if item.ndim == 2 and item.shape[0] else (1 if mine.ndim == 1 else int(np.prod(mine.shape[1:])))
    buf = torch.zeros((pad, width), dtype=torch.from_numpy(np.zeros(1, mine.dtype)).dtype, device=dev)
    if len(mine):
        buf[:len(mine)] = torch.from_numpy(item.copy()).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0)
    if rank != 0:
        return None
    parts = [o[:sizes[r]].cpu().numpy() for r, o in enumerate(outs)]
    full = np.concatenate(parts, axis=0)
    return full.reshape((-1,) + mine.shape[1:]) if mine.ndim > 1 else full.reshape(-1)
