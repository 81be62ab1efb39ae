#!/usr/bin/env python3
"""Does enabling CUDA peer access (what an NCCL communicator and fmsi_gpu_pool_create's replication do) slow the
random-probe kernel? One process, GPU 0: human-scale fold tier, 2^26 queries per launch, timed (CUDA events) before
peer access, with it enabled 0<->1, after disabling it again, and with allocations made while it was enabled."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fmsi_b200 as fg  # noqa: E402
from bench import device_genome, device_queries  # noqa: E402

n, k, batch = int(os.environ.get("GENOME", 3_100_000_000)), 31, 1 << 26
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
import ctypes  # noqa: E402
import glob  # noqa: E402
_cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12", "libcudart.so"]
rt = None
for c in _cands:
    try:
        rt = ctypes.CDLL(c)
        break
    except OSError:
        pass


def build():
    codes, ascii_ = device_genome(n, 4, k, dev)
    gi = fg.Index.build(ascii_.data_ptr(), k, with_klcp=False, device=0, n=n, mem=fg.MEM_DEVICE, dict=2)
    q = [device_queries(codes, k, batch, s, dev) for s in (1, 2)]
    return gi, q


def timed(gi, q, reps=10):
    out = torch.empty(batch, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
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


res = {}
gi, q = build()
res["before_peer_access_ms"] = timed(gi, q)
rt.cudaSetDevice(0)
res["enable_0_to_1_rc"] = int(rt.cudaDeviceEnablePeerAccess(1, 0))
res["peer_enabled_0_to_1_ms"] = timed(gi, q)
rt.cudaSetDevice(1)
res["enable_1_to_0_rc"] = int(rt.cudaDeviceEnablePeerAccess(0, 0))
rt.cudaSetDevice(0)
torch.cuda.set_device(0)
res["peer_enabled_both_ways_ms"] = timed(gi, q)
gi.close()
del q
torch.cuda.empty_cache()
gi, q = build()
res["allocated_while_enabled_ms"] = timed(gi, q)
rt.cudaSetDevice(0)
res["disable_0_to_1_rc"] = int(rt.cudaDeviceDisablePeerAccess(1))
rt.cudaSetDevice(1)
res["disable_1_to_0_rc"] = int(rt.cudaDeviceDisablePeerAccess(0))
rt.cudaSetDevice(0)
torch.cuda.set_device(0)
res["after_disabling_ms"] = timed(gi, q)
print(json.dumps(res))
