#!/usr/bin/env python3
"""Host<->device link ceiling on this box (context for bench.py's `e2e`, which moves 8 B in + 1 B out per k-mer):
pinned-memory cudaMemcpyAsync in 128-MiB pieces, H2D alone, D2H alone, and both directions at once on two streams."""
import json

import torch

dev = torch.device("cuda", 0)
n = 128 << 20
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
h_out = [torch.empty(n // 8, dtype=torch.uint8).pin_memory() for _ in range(2)]
d_in = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
d_out = [torch.empty(n // 8, dtype=torch.uint8, device=dev) for _ in range(2)]
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, reps=24):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for r in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[r & 1].copy_(h_in[r & 1], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[r & 1].copy_(d_out[r & 1], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps


for _ in range(2):
    t_h = run(True, False)
    t_d = run(False, True)
    t_b = run(True, True)
print(json.dumps({"h2d_gbs_alone": round(n / t_h / 1e9, 2), "d2h_gbs_alone_16MiB_pieces": round(n / 8 / t_d / 1e9, 2),
                  "h2d_gbs_with_d2h_1_8th": round(n / t_b / 1e9, 2),
                  "kmers_per_s_ceiling_at_8B_in": round(n / t_b / 8 / 1e9, 3), "piece_mib": n >> 20}))
