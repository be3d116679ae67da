#!/usr/bin/env python3
"""kernel-level times (torch profiler) of the tile sort / spread / gather on one workload"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile
from torchpme_b200 import _native
from torchpme_b200.mesh import CellGeometry
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tile_check import WORK

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
shuffle = len(sys.argv) > 2 and sys.argv[2] == "shuffle"
n_side, n_mesh, dtype, method = WORK[name]
gen = torch.Generator().manual_seed(0)
length = n_side * 2.82
ar = torch.arange(n_side)
sites = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3)
pos = (sites.double() * 2.82 + 0.1 * torch.randn(sites.shape, generator=gen, dtype=torch.float64)) % length
q = (1.0 - 2.0 * (sites.sum(1) % 2).double()).reshape(-1, 1)
if shuffle:
    perm = torch.randperm(pos.shape[0], generator=gen); pos, q = pos[perm], q[perm]
pos, q = pos.to("cuda", dtype).contiguous(), q.to("cuda", dtype).contiguous()
ns = (n_mesh,) * 3
r2u = CellGeometry(torch.eye(3, dtype=torch.float64) * length).r2u(ns)
_native.TILE_MODE = "on"
plan = _native.tile_plan(dtype, ns, 4, method, pos.shape[0])
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def once():
    flush.fill_(1.0)
    tiles = _native.TileSort(plan, pos, r2u)
    flush.fill_(1.0)
    rho = _native.spread(pos, q, r2u, ns, 4, method, tiles=tiles)
    flush.fill_(1.0)
    _native.gather(rho, pos, r2u, 4, method, want_grad=True, tiles=tiles)
for _ in range(3):
    once()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        once()
    torch.cuda.synchronize()
print(f"# {name} shuffle={shuffle} tile=({plan.tx},{plan.ty}) debug={os.environ.get('TPME_TILE_DEBUG')}")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if "tile" in e.key or "Memset" in e.key:
        print(f"{e.device_time_total / e.count:9.1f} us x{e.count:3d}  {e.key[:100]}")
