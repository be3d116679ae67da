#!/usr/bin/env python3
"""
GPU check of the tiled mesh kernels (csrc/tiles.cu) against the direct ones (csrc/interp.cu):
results of spread / gather / derivative gather / vjp on the bench workloads, and event timings of
both families with the L2 flushed between launches, for lattice-ordered and shuffled atoms.

    python scripts/tile_check.py [c2 c3 c4 ...] [--tile tx,ty] [--shuffle] [--reps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from torchpme_b200 import _native  # noqa: E402
from torchpme_b200.mesh import CellGeometry  # noqa: E402

WORK = {"c2": (32, 64, torch.float32, 0), "c3": (64, 128, torch.float64, 1), "c4": (100, 256, torch.float32, 0),
        "c5": (64, 128, torch.float32, 1), "t1": (12, 32, torch.float64, 0)}


def timed(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def run(name, shuffle, reps, out):
    n_side, n_mesh, dtype, method = WORK[name]
    dev = "cuda"
    gen = torch.Generator().manual_seed(0)
    length = n_side * 2.82
    ar = torch.arange(n_side)
    sites = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3)
    pos = (sites.double() * 2.82 + 0.1 * torch.randn(sites.shape, generator=gen, dtype=torch.float64)) % length
    q = (1.0 - 2.0 * (sites.sum(1) % 2).double()).reshape(-1, 1)
    if shuffle:
        perm = torch.randperm(pos.shape[0], generator=gen)
        pos, q = pos[perm], q[perm]
    pos, q = pos.to(dev, dtype).contiguous(), q.to(dev, dtype).contiguous()
    cell = torch.eye(3, dtype=torch.float64) * length
    ns = (n_mesh,) * 3
    r2u = CellGeometry(cell).r2u(ns)
    n = pos.shape[0]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    _native.TILE_SPREAD = "on"
    tiles = _native.TileSort(_native.tile_plan(dtype, ns, 4, method, n), pos, r2u)
    torch.cuda.synchronize()
    plan = tiles.plan
    res = {"workload": name, "shuffled": bool(shuffle), "tile": [plan.tx, plan.ty], "smem": plan.smem_bytes,
           "threads": [plan.spread_threads, plan.gather_threads]}
    # ---- results
    rho_d = _native.spread(pos, q, r2u, ns, 4, method)
    rho_t = _native.spread(pos, q, r2u, ns, 4, method, tiles=tiles)
    res["spread_rel"] = rel(rho_t.double(), rho_d.double())
    res["spread_sum"] = [float(rho_t.double().sum()), float(rho_d.double().sum()), float(q.double().sum())]
    phi = torch.randn(rho_d.shape, dtype=dtype, device=dev)
    v_d, dv_d = _native.gather(phi, pos, r2u, 4, method, want_grad=True)
    v_t, dv_t = _native.gather(phi, pos, r2u, 4, method, want_grad=True, tiles=tiles)
    res["gather_rel"] = rel(v_t.double(), v_d.double())
    res["dgather_rel"] = rel(dv_t.double(), dv_d.double())
    coef = torch.randn_like(q)
    g_d, vv_d, gr_d = _native.gather_vjp(phi, pos, coef, r2u, 4, method, want_values=True, want_grad_r2u=True)
    g_t, vv_t, gr_t = _native.gather_vjp(phi, pos, coef, r2u, 4, method, want_values=True, want_grad_r2u=True,
                                         tiles=tiles)
    res["vjp_rel"] = rel(g_t.double(), g_d.double())
    res["vjp_values_rel"] = rel(vv_t.double(), vv_d.double())
    res["grad_r2u_rel"] = rel(gr_t.double(), gr_d.double())
    # run-to-run reproducibility of the tiled spread (canonical bin order)
    tiles2 = _native.TileSort(plan, pos, r2u)
    rho_t2 = _native.spread(pos, q, r2u, ns, 4, method, tiles=tiles2)
    res["spread_bitwise_repeat"] = bool(torch.equal(rho_t, rho_t2))
    # ---- timings (us, median, L2 flushed before each launch)
    res["us"] = {
        "sort": timed(lambda: _native.TileSort(plan, pos, r2u), reps, flush),
        "spread_direct": timed(lambda: _native.spread(pos, q, r2u, ns, 4, method, out=rho_d), reps, flush),
        "spread_tiled": timed(lambda: _native.spread(pos, q, r2u, ns, 4, method, out=rho_t, tiles=tiles), reps, flush),
        "gather_direct": timed(lambda: _native.gather(phi, pos, r2u, 4, method, want_grad=True), reps, flush),
        "gather_tiled": timed(lambda: _native.gather(phi, pos, r2u, 4, method, want_grad=True, tiles=tiles), reps, flush),
        "vjp_direct": timed(lambda: _native.gather_vjp(phi, pos, coef, r2u, 4, method, want_values=True), reps, flush),
        "vjp_tiled": timed(lambda: _native.gather_vjp(phi, pos, coef, r2u, 4, method, want_values=True, tiles=tiles),
                           reps, flush),
    }
    res["us"] = {k: round(v, 2) for k, v in res["us"].items()}
    print(json.dumps(res), flush=True)
    out.append(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["t1", "c2", "c3", "c4"])
    ap.add_argument("--shuffle", action="store_true")
    ap.add_argument("--both", action="store_true", help="lattice order and shuffled")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    _native.TILE_MODE = "on"
    out = []
    for w in args.workloads:
        for sh in ([False, True] if args.both else [args.shuffle]):
            run(w, sh, args.reps, out)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
