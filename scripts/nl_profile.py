#!/usr/bin/env python3
"""kernel-level profile (torch profiler) of the device neighbor list build, the pair-distance kernels and the
graphed positions-only step on a bench workload:  python scripts/nl_profile.py 64 f64 [shuffle]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile
import torchpme_b200 as tp
from torchpme_b200.neighbors import neighbor_list, distances_from
from torchpme_b200.synthetic import rocksalt

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dtype = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
pos, q, cell, idx, d = rocksalt(n_side, dtype=dtype, device="cuda", cutoff=6.0)
if "shuffle" in sys.argv:
    perm = torch.randperm(pos.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    pos, q = pos[perm].contiguous(), q[perm].contiguous()
for _ in range(3):
    out = neighbor_list(pos, cell, 6.0, index_dtype=torch.int32)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = neighbor_list(pos, cell, 6.0, index_dtype=torch.int32)
torch.cuda.synchronize()
print(f"n={pos.shape[0]} pairs={out[0].shape[0]} (generator {idx.shape[0]}) wall {1e3 * (time.perf_counter() - t0) / 5:.3f} ms per build")
n_mesh = {32: 64, 64: 128, 100: 256}.get(n_side, 64)
L = float(cell[0, 0])
calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=L / (n_mesh / 2 - 2))
step = tp.GraphedPositionsStep(calc, q, cell, pos, cutoff=6.0, host_io=True)


def once():
    i, d0, s = neighbor_list(pos, cell, 6.0, index_dtype=torch.int32)
    p = pos.clone().requires_grad_(True)
    dd = distances_from(p, cell, i, s)
    dd.sum().backward()
    step.replay()


once()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        once()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:28]:
    print(f"{e.device_time_total / e.count:9.1f} us x{e.count:3d}  {e.key[:110]}")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    step.replay()
ev[0].record()
for _ in range(20):
    step.replay()
ev[1].record()
torch.cuda.synchronize()
print(f"graphed positions-only step incl. host I/O: {ev[0].elapsed_time(ev[1]) / 20:.4f} ms, overflowed={step.overflowed()}")
