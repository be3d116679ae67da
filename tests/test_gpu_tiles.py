"""
The tiled mesh kernels (csrc/tiles.cu: cell-sorted atoms, shared-memory pencils moved by TMA bulk
copies) through the C ABI, against the numpy oracle (mesh_interpolator.py:303-457 restated) and the
direct kernels of csrc/interp.cu: cubic and triclinic cells, atoms outside the cell, both stencil
families, fp32 / fp64, several channels, every output mode, non-cubic meshes, tile shapes.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import pme_oracle as oracle

pytestmark = pytest.mark.gpu


def _system(n, n_channels, triclinic, seed, box=20.0):
    rng = np.random.default_rng(seed)
    cell = np.eye(3) * box
    if triclinic:
        cell = cell + rng.uniform(-0.1, 0.1, (3, 3)) * box
    pos = rng.uniform(0, 1, (n, 3)) @ cell
    pos[: n // 10] += 2.0 * cell[0] - 1.0 * cell[2]     # atoms outside the cell: index wrap
    return pos, rng.normal(size=(n, n_channels)), cell


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("method", ["P3M", "Lagrange"])
@pytest.mark.parametrize("ns, triclinic, channels", [((32, 32, 32), False, 1), ((16, 32, 64), True, 2),
                                                      ((64, 16, 8), True, 3), ((8, 8, 128), False, 1)])
def test_tiled_kernels_match_oracle_and_direct(ns, triclinic, channels, method, dtype):
    from torchpme_b200 import _native
    from torchpme_b200.mesh import CellGeometry

    pos_np, w_np, cell_np = _system(3000, channels, triclinic, seed=sum(ns) + channels)
    dev = "cuda"
    pos = torch.tensor(pos_np, dtype=dtype, device=dev)
    w = torch.tensor(w_np, dtype=dtype, device=dev)
    r2u = CellGeometry(torch.tensor(cell_np)).r2u(ns)
    mid = _native.METHOD_ID[method]
    saved = _native.TILE_MODE, _native.TILE_SPREAD
    _native.TILE_MODE = _native.TILE_SPREAD = "on"
    try:
        tiles = _native.tile_sort(pos, r2u, ns, 4, mid)
    finally:
        _native.TILE_MODE, _native.TILE_SPREAD = saved
    assert tiles is not None, "this mesh / stencil is covered by the tiled kernels"
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    # the oracle sees the numbers the kernels saw
    p64, w64 = pos.double().cpu().numpy(), w.double().cpu().numpy()
    # ---- spread
    rho_t = _native.spread(pos, w, r2u, ns, 4, mid, tiles=tiles)
    rho_d = _native.spread(pos, w, r2u, ns, 4, mid)
    rho_o = oracle.points_to_mesh(w64, p64, cell_np, np.array(ns), 4, method)
    assert rel_err(rho_t, rho_o) < tol
    assert rel_err(rho_t, rho_d) < tol
    assert abs(float(rho_t.double().sum()) - w64.sum()) < 1e-3 * (1 + abs(w64).sum()) * (1e-9 if dtype == torch.float64 else 1)
    # ---- gather: values and dvalues/dr
    gen = torch.Generator().manual_seed(1)
    phi = torch.randn((channels,) + tuple(ns), generator=gen, dtype=torch.float64).to(dev, dtype)
    v_t, dv_t = _native.gather(phi, pos, r2u, 4, mid, want_grad=True, tiles=tiles)
    v_d, dv_d = _native.gather(phi, pos, r2u, 4, mid, want_grad=True)
    v_o = oracle.mesh_to_points(phi.double().cpu().numpy(), p64, cell_np, 4, method)
    assert rel_err(v_t, v_o) < tol
    assert rel_err(v_t, v_d) < tol and rel_err(dv_t, dv_d) < 10 * tol
    # values only / derivative only
    v_only, none = _native.gather(phi, pos, r2u, 4, mid, want_grad=False, tiles=tiles)
    assert none is None and rel_err(v_only, v_d) < tol
    # ---- vjp (+ values, + cell reduction, accumulate)
    coef = torch.randn(w.shape, generator=gen, dtype=torch.float64).to(dev, dtype)
    g_t, vv_t, gr_t = _native.gather_vjp(phi, pos, coef, r2u, 4, mid, want_values=True, want_grad_r2u=True, tiles=tiles)
    g_d, vv_d, gr_d = _native.gather_vjp(phi, pos, coef, r2u, 4, mid, want_values=True, want_grad_r2u=True)
    assert rel_err(g_t, g_d) < 10 * tol and rel_err(vv_t, vv_d) < tol
    assert rel_err(gr_t, gr_d) < (1e-10 if dtype == torch.float64 else 1e-3)
    # d/dr of sum_c coef * gather equals the contraction of dvalues
    assert rel_err(g_t, torch.einsum("ic,icd->id", coef, dv_t)) < 10 * tol
    acc = g_d.clone()
    _native.gather_vjp(phi, pos, coef, r2u, 4, mid, grad_positions=acc, tiles=tiles)
    assert rel_err(acc, 2 * g_d) < 10 * tol


@pytest.mark.parametrize("tile", ["8,16", "8,8", "4,8", "4,4", "2,2"])
def test_tile_shapes(tile, monkeypatch):
    """every pencil footprint gives the same mesh and the same gathered values"""
    from torchpme_b200 import _native
    from torchpme_b200.mesh import CellGeometry

    monkeypatch.setenv("TPME_TILE", tile)
    monkeypatch.setattr(_native, "_tile_plans", {})
    monkeypatch.setattr(_native, "TILE_MODE", "on")
    monkeypatch.setattr(_native, "TILE_SPREAD", "on")
    ns = (32, 64, 128)
    pos_np, w_np, cell_np = _system(20000, 1, True, seed=5, box=40.0)
    pos = torch.tensor(pos_np, dtype=torch.float32, device="cuda")
    w = torch.tensor(w_np, dtype=torch.float32, device="cuda")
    r2u = CellGeometry(torch.tensor(cell_np)).r2u(ns)
    tiles = _native.tile_sort(pos, r2u, ns, 4, 0)
    assert tiles is not None and [tiles.plan.tx, tiles.plan.ty] == [int(v) for v in tile.split(",")]
    # the sort is a permutation and every bin holds the atoms whose first node lies in it
    idx = tiles.idx.cpu().numpy()
    assert np.array_equal(np.sort(idx), np.arange(pos.shape[0]))
    start = tiles.bin_start.cpu().numpy()
    assert start[0] == 0 and start[-1] == pos.shape[0] and (np.diff(start) >= 0).all()
    rho_t = _native.spread(pos, w, r2u, ns, 4, 0, tiles=tiles)
    rho_d = _native.spread(pos, w, r2u, ns, 4, 0)
    assert rel_err(rho_t, rho_d) < 2e-5
    v_t, _ = _native.gather(rho_d, pos, r2u, 4, 0, tiles=tiles)
    v_d, _ = _native.gather(rho_d, pos, r2u, 4, 0)
    assert rel_err(v_t, v_d) < 2e-5


def test_tiled_path_is_the_default_for_large_systems_and_falls_back_otherwise():
    from torchpme_b200 import _native
    from torchpme_b200.mesh import CellGeometry

    r2u = CellGeometry(torch.eye(3, dtype=torch.float64) * 10).r2u((32, 32, 32))
    big = torch.rand(_native.TILE_MIN_POINTS, 3, device="cuda") * 10
    small = torch.rand(100, 3, device="cuda") * 10
    if _native.TILE_MODE == "auto":
        assert _native.tile_sort(big, r2u, (32, 32, 32), 4, 0) is not None
        assert _native.tile_sort(small, r2u, (32, 32, 32), 4, 0) is None          # too few atoms to pay for the sort
    assert _native.tile_sort(big, r2u, (32, 32, 32), 5, 0) is None                 # 5 nodes: direct kernels
    assert _native.tile_sort(big, CellGeometry(torch.eye(3, dtype=torch.float64) * 10).r2u((30, 32, 32)),
                             (30, 32, 32), 4, 0) is None                             # not a power of two


def test_empty_and_single_point():
    from torchpme_b200 import _native
    from torchpme_b200.mesh import CellGeometry

    ns = (16, 16, 16)
    r2u = CellGeometry(torch.eye(3, dtype=torch.float64) * 8).r2u(ns)
    saved, _native.TILE_MODE = _native.TILE_MODE, "on"
    saved_spread, _native.TILE_SPREAD = _native.TILE_SPREAD, "on"
    try:
        assert _native.tile_sort(torch.empty(0, 3, device="cuda"), r2u, ns, 4, 0) is None
        one = torch.tensor([[7.9, 0.1, 4.0]], device="cuda", dtype=torch.float64)
        tiles = _native.tile_sort(one, r2u, ns, 4, 1)
        w = torch.ones(1, 1, device="cuda", dtype=torch.float64)
        rho = _native.spread(one, w, r2u, ns, 4, 1, tiles=tiles)
        assert rel_err(rho, _native.spread(one, w, r2u, ns, 4, 1)) < 1e-13
        assert abs(float(rho.sum()) - 1.0) < 1e-12
    finally:
        _native.TILE_MODE, _native.TILE_SPREAD = saved, saved_spread
