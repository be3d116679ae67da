"""
GPU tests of the slab-decomposed (multi-GPU) path.

* single GPU, W *virtual* ranks: the slab kernels (``tpme_spread_slab``, ``tpme_gather_slab``,
  ``tpme_slab_fft_yz``, ``tpme_slab_fft_x_green``, ``tpme_slab_exchange_copy``) composed the way the
  peer-memory exchange composes them must reproduce the single-GPU kernels;
* single GPU, world_size 1 process group: ``SlabP3MCalculator`` / ``SlabPMECalculator`` (both
  transports) against the plain calculators and the oracle;
* two or more GPUs (skipped otherwise): ``torchrun`` of ``slab_gpu_worker.py`` -- every rank
  checks potentials and gradients against the numpy oracle for both transports.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err, rocksalt

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("method, nodes, channels, ns", [
    ("P3M", 4, 1, (32, 32, 32)), ("Lagrange", 5, 2, (16, 32, 64)), ("P3M", 3, 1, (64, 16, 256)),
    ("Lagrange", 7, 1, (8, 8, 8)),
])
def test_virtual_ranks_reproduce_single_gpu(method, nodes, channels, ns, world, dtype):
    from torchpme_b200 import _native
    from torchpme_b200.distributed import SlabLayout
    from torchpme_b200.mesh import geometry_of

    dev = "cuda"
    gen = torch.Generator().manual_seed(nodes + world)
    cell = (torch.eye(3, dtype=torch.float64) * 11.0 + 0.8 * torch.rand(3, 3, generator=gen, dtype=torch.float64)).to(dev)
    n = 3000
    frac = torch.rand(n, 3, generator=gen, dtype=torch.float64) * 1.4 - 0.2   # some atoms outside the cell
    pos = (frac.to(dev) @ cell).to(dtype)
    w = torch.randn(n, channels, generator=gen, dtype=torch.float64).to(dev).to(dtype)
    geom = geometry_of(cell)
    r2u = geom.r2u(ns)
    mid = _native.METHOD_ID[method]
    nx, ny, nz = ns
    nzh = nz // 2 + 1
    layouts = [SlabLayout(ns, world, r) for r in range(world)]
    tol = 1e-11 if dtype == torch.float64 else 2e-5

    # ---- spread: slabs concatenated == full mesh
    rho_full = _native.spread(pos, w, r2u, ns, nodes, mid)
    rho = [_native.spread(pos, w, r2u, ns, nodes, mid, slab=(lay.x0, lay.nxl)) for lay in layouts]
    assert rel_err(torch.cat(rho, dim=1), rho_full) < tol
    # the same through the per-slab point lists (only the atoms that reach into the slab are visited)
    plists = [_native.slab_select_points(pos, r2u, ns, nodes, (lay.x0, lay.nxl)) for lay in layouts]
    rho_l = [_native.spread(pos, w, r2u, ns, nodes, mid, slab=(lay.x0, lay.nxl), point_list=pl)
             for lay, pl in zip(layouts, plists)]
    assert rel_err(torch.cat(rho_l, dim=1), rho_full) < tol
    counts = [int(pl[1]) for pl in plists]
    assert all(0 < c <= n for c in counts) and sum(counts) >= n
    if nx >= 4 * nodes * world:
        assert sum(counts) < 2 * n

    # ---- filter through W virtual ranks (the peer-exchange data path on one device)
    green = _native.make_green(_native.GREEN_COULOMB, 0.41, geom.recip, geom.spacing(ns), smearing=1.3,
                               prefactor=1.1, p3m_nodes=nodes if method == "P3M" else 0)
    phi_full, _ = _native.kfilter_apply(rho_full, green)
    esize = 8 if dtype == torch.float32 else 16
    X = [torch.empty((channels, lay.nxl, ny, nzh, 2), dtype=dtype, device=dev) for lay in layouts]
    T = [torch.empty((channels, nx, lay.nyl, nzh, 2), dtype=dtype, device=dev) for lay in layouts]
    lay0 = layouts[0]
    blk, run = lay0.block, lay0.nyl * nzh
    for r, lay in enumerate(layouts):
        _native.slab_fft_yz(True, rho[r], X[r])
    for r, lay in enumerate(layouts):
        _native.slab_exchange_copy(X[r], [T[p].data_ptr() + r * blk * esize for p in range(world)], channels,
                                   world, lay.nxl, run, (lay.nxl * ny * nzh, run, ny * nzh), (world * blk, run))
    for r, lay in enumerate(layouts):
        _native.slab_fft_x_green(T[r], ns, lay.y0, green)
    for r, lay in enumerate(layouts):
        _native.slab_exchange_copy(T[r], [X[p].data_ptr() + r * run * esize for p in range(world)], channels,
                                   world, lay.nxl, run, (world * blk, blk, run), (lay.nxl * ny * nzh, ny * nzh))
    phi = []
    for r, lay in enumerate(layouts):
        out = torch.empty_like(rho[r])
        _native.slab_fft_yz(False, out, X[r])
        phi.append(out)
    assert rel_err(torch.cat(phi, dim=1), phi_full) < tol

    # ---- the same through the fused compute + exchange kernels (y pass / x pass push to the "peers")
    for t in X + T:
        t.fill_(float("nan"))
    for r in range(world):
        peers = _native.make_slab_peers(r, [t.data_ptr() for t in X], [t.data_ptr() for t in T])
        _native.slab_fft_yz_push(rho[r], ns, peers)
    for r in range(world):
        peers = _native.make_slab_peers(r, [t.data_ptr() for t in X], [t.data_ptr() for t in T])
        _native.slab_fft_x_green_push(dtype, torch.device(dev), channels, ns, green, peers)
    phi2 = []
    for r in range(world):
        out = torch.empty_like(rho[r])
        _native.slab_fft_yz(False, out, X[r])
        phi2.append(out)
    assert rel_err(torch.cat(phi2, dim=1), phi_full) < tol

    # ---- gather: partial sums over the slabs == full gather (values, dV/dr, vjp)
    v_full, dv_full = _native.gather(phi_full, pos, r2u, nodes, mid, want_grad=True)
    gp_full, _, _ = _native.gather_vjp(phi_full, pos, w, r2u, nodes, mid)
    v_sum, dv_sum, gp_sum = 0, 0, 0
    for r, lay in enumerate(layouts):
        v, dv = _native.gather(phi[r], pos, r2u, nodes, mid, want_grad=True, slab=(lay.x0, nx))
        gp, _, _ = _native.gather_vjp(phi[r], pos, w, r2u, nodes, mid, slab=(lay.x0, nx))
        # list mode leaves foreign points untouched: start from zeros / accumulate
        v_l = torch.zeros_like(v)
        dv_l = torch.zeros_like(dv)
        gp_l = torch.zeros_like(gp)
        lst, cnt = plists[r]
        sel = lst[: int(cnt)].long()
        _, dv_tmp = _native.gather(phi[r], pos, r2u, nodes, mid, want_grad=True, values_out=v_l, slab=(lay.x0, nx),
                                   point_list=plists[r])
        dv_l[sel] = dv_tmp[sel]
        _native.gather_vjp(phi[r], pos, w, r2u, nodes, mid, grad_positions=gp_l, slab=(lay.x0, nx),
                           point_list=plists[r])
        assert rel_err(v_l, v) < tol and rel_err(gp_l, gp) < tol * 10 and rel_err(dv_l, dv) < tol * 10
        v_sum, dv_sum, gp_sum = v_sum + v, dv_sum + dv, gp_sum + gp
    assert rel_err(v_sum, v_full) < tol
    if nodes > 1:
        assert rel_err(dv_sum, dv_full) < tol * 10
        assert rel_err(gp_sum, gp_full) < tol * 10


@pytest.fixture(scope="module")
def world1_group():
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "p2p", "p2p-copy"])
@pytest.mark.parametrize("method, dtype", [("P3M", torch.float32), ("Lagrange", torch.float64)])
def test_world_size_one_matches_plain_calculator(world1_group, method, dtype, transport):
    import torchpme_b200 as tp
    from torchpme_b200.distributed import SlabP3MCalculator, SlabPMECalculator

    pos, q, cell, idx, d = rocksalt(8, dtype=dtype, device="cuda")
    L = float(cell[0, 0])
    pot = tp.CoulombPotential(smearing=1.2).to("cuda")
    plain_cls, slab_cls = (tp.P3MCalculator, SlabP3MCalculator) if method == "P3M" else (tp.PMECalculator, SlabPMECalculator)
    gout = torch.randn(q.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to("cuda").to(dtype)
    res = []
    for calc in (plain_cls(pot, mesh_spacing=L / 6), slab_cls(pot, mesh_spacing=L / 6, transport=transport)):
        p = pos.clone().requires_grad_(True)
        qq = q.clone().requires_grad_(True)
        dd = d.clone().requires_grad_(True)
        V = calc(qq, cell, p, idx, dd)
        (V * gout).sum().backward()
        res.append((V.detach(), p.grad, qq.grad, dd.grad))
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    for a, b in zip(*res):
        assert rel_err(b, a) < tol


def test_two_gpus_against_oracle():
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    world = 4 if n_dev >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "slab_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(lines) >= 1
    for res in lines:
        tol = 1e-9 if res["dtype"] == "float64" else 1e-3
        for key in ("V", "dpos", "dq", "dd"):
            assert res[key] < tol, res


def test_two_ranks_on_one_gpu_against_oracle():
    """
    The real multi-process slab path on a ONE-GPU box: two processes share cuda:0 (torch.distributed
    over gloo, TPME_SLAB_ONE_GPU=1), so CUDA IPC mappings, the device-side flag barrier, the FFT kernels
    that store into the peer's arrays and the peer all-reduce all run between two processes, against
    the numpy oracle.  The GPU time-slices the two contexts, so the barriers take milliseconds.
    """
    world = 2
    env = dict(os.environ, TPME_SLAB_ONE_GPU="1", TPME_PEER_TIMEOUT="120")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "slab_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [json.loads(line) for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 8     # 2 transports x 2 dtypes x 2 methods
    for res in lines:
        tol = 1e-9 if res["dtype"] == "float64" else 1e-3
        for key in ("V", "dpos", "dq", "dd"):
            assert res[key] < tol, res
