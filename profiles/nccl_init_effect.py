#!/usr/bin/env python3
"""Does an initialised NCCL communicator slow the random-probe kernel (bench.py at N > 1 lost 5.5 % per rank with
NCCL bookkeeping)? One process, one GPU: human-scale fold tier timed before NCCL, after init_process_group (world 1)
and one all_reduce, and after destroying the group."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
from bench import device_genome, device_queries  # noqa: E402

n, k, batch = int(os.environ.get("GENOME", 3_100_000_000)), 31, 1 << 26
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
codes, ascii_ = device_genome(n, 4, k, dev)
gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=False, device=0, n=n, mem=fg.MEM_DEVICE, dict=2)
q = [device_queries(codes, k, batch, s, dev) for s in (1, 2)]
out = torch.empty(batch, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream


def timed(reps=10):
    for i in range(3):
        gi.query_kmers_ptr(q[i & 1].data_ptr(), batch, out.data_ptr(), k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, fg.MEM_DEVICE, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        gi.query_kmers_ptr(q[i & 1].data_ptr(), batch, out.data_ptr(), k, fg.MODE_ALL, fg.OUT_PRESENCE, fg.STRANDS_LAZY, fg.MEM_DEVICE, st)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


res = {"before_nccl_ms": timed()}
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
res["after_init_ms"] = timed()
t = torch.ones(4, device=dev)
dist.all_reduce(t)
dist.barrier()
torch.cuda.synchronize()
res["after_all_reduce_and_barrier_ms"] = timed()
dist.destroy_process_group()
res["after_destroy_ms"] = timed()
print(json.dumps(res))
