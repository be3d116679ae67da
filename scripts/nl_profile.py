#!/usr/bin/env python3
"""kernel / op level profile of the device neighbor list build (torch profiler) on a bench workload"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile
from torchpme_b200.neighbors import neighbor_list, distances_from
from torchpme_b200.synthetic import rocksalt

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dtype = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
pos, q, cell, idx, d = rocksalt(n_side, dtype=dtype, device="cuda", cutoff=6.0)
for _ in range(3):
    out = neighbor_list(pos, cell, 6.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = neighbor_list(pos, cell, 6.0)
torch.cuda.synchronize()
print(f"n={pos.shape[0]} pairs={out[0].shape[0]} wall {1e3 * (time.perf_counter() - t0) / 5:.3f} ms per build")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        out = neighbor_list(pos, cell, 6.0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
