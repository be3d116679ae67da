"""
Neighbor list built on the GPU (torchpme_b200.neighbors: cell-list kernels tpme_nl_count /
tpme_nl_fill) against the brute-force oracle (oracle.neighbor_list, the stand-in for the
reference's external `vesin` dependency, tests/helpers.py:240-275): the SAME set of (i, j, image)
pairs with the same distances, for cubic / triclinic / smaller-than-cutoff cells, half and full lists,
fp32 and fp64, 2-D periodic and open systems; at 1 M atoms against the pair list of the synthetic
generator; and the differentiable distances against autograd through the reference formula.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err, rocksalt
from oracle import pme_oracle as oracle
from test_neighbors import CELLS, _canonical

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("full", [False, True], ids=["half", "full"])
@pytest.mark.parametrize("cell_name", sorted(CELLS))
def test_pair_set_matches_oracle(cell_name, full, dtype):
    from torchpme_b200.neighbors import neighbor_list

    cell = CELLS[cell_name]
    rng = np.random.default_rng(11)
    n = 60 if "small" in cell_name else 400
    pos = rng.uniform(-0.3, 1.3, (n, 3)) @ cell        # some atoms outside the cell
    cutoff = 3.0
    idx, d, s = neighbor_list(torch.tensor(pos, dtype=dtype, device="cuda"), torch.tensor(cell, dtype=dtype, device="cuda"),
                              cutoff, full_neighbor_list=full)
    o_idx, o_d, o_s = oracle.neighbor_list(pos, cell, cutoff, full=full)
    # pairs closer to the cutoff than the working precision resolves may legitimately differ
    eps = 1e-9 if dtype == torch.float64 else 2e-5
    sure = np.abs(o_d - cutoff) > eps
    d_np = d.double().cpu().numpy()
    sure_gpu = np.abs(d_np - cutoff) > eps
    if full:
        got = np.concatenate([idx.cpu().numpy(), s.cpu().numpy()], axis=1)[sure_gpu]
        want = np.concatenate([o_idx, o_s], axis=1)[sure]
        got = got[np.lexsort(got.T[::-1])]
        want = want[np.lexsort(want.T[::-1])]
        assert got.shape == want.shape and np.array_equal(got, want)
    else:
        rows, dist = _canonical(idx.cpu().numpy()[sure_gpu], d_np[sure_gpu], s.cpu().numpy()[sure_gpu])
        o_rows, o_dist = _canonical(o_idx[sure], o_d[sure], o_s[sure])
        assert rows.shape == o_rows.shape and np.array_equal(rows, o_rows)
        assert np.abs(dist - o_dist).max() < (1e-12 if dtype == torch.float64 else 1e-5)


@pytest.mark.parametrize("periodic", [(True, True, False), (False, False, False)])
def test_partially_periodic(periodic):
    """slab / open boundaries: images only along the periodic axes (checked against a brute force)"""
    from torchpme_b200.neighbors import neighbor_list

    cell = np.eye(3) * 9.0
    rng = np.random.default_rng(3)
    pos = rng.uniform(0, 1, (300, 3)) @ cell
    cutoff = 2.5
    idx, d, s = neighbor_list(torch.tensor(pos, device="cuda"), torch.tensor(cell, device="cuda"), cutoff,
                              periodic=periodic)
    o_idx, o_d, o_s = oracle.neighbor_list(pos, cell, cutoff)
    ok = np.ones(len(o_d), dtype=bool)
    for a in range(3):
        if not periodic[a]:
            ok &= o_s[:, a] == 0
    rows, dist = _canonical(idx.cpu().numpy(), d.cpu().numpy(), s.cpu().numpy())
    o_rows, o_dist = _canonical(o_idx[ok], o_d[ok], o_s[ok])
    assert np.array_equal(rows, o_rows) and np.abs(dist - o_dist).max() < 1e-12


def test_distances_are_differentiable_like_the_reference_formula():
    """distances_from(positions, cell, idx, shifts): gradients to positions AND cell (the path MD codes use)"""
    from torchpme_b200.neighbors import distances_from, neighbor_list

    pos, q, cell, _, _ = rocksalt(5, dtype=torch.float64, device="cuda", cutoff=5.0)
    cell = cell + 0.3 * torch.rand(3, 3, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    idx, d, s = neighbor_list(pos, cell, 4.0)
    p = pos.clone().requires_grad_(True)
    c = cell.clone().requires_grad_(True)
    dd = distances_from(p, c, idx, s)
    assert rel_err(dd.detach(), d) < 1e-12
    w = torch.randn(dd.shape, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    (dd * w).sum().backward()
    # analytic: d|r_ij|/dr_j = r_ij / |r_ij|, d/dcell = S^T (r_ij / |r_ij|)
    rij = (pos[idx[:, 1]] - pos[idx[:, 0]] + s.double() @ cell)
    unit = rij / rij.norm(dim=1, keepdim=True) * w[:, None]
    gp = torch.zeros_like(pos).index_add_(0, idx[:, 1], unit).index_add_(0, idx[:, 0], -unit)
    gc = s.double().T @ unit
    assert rel_err(p.grad, gp) < 1e-12 and rel_err(c.grad, gc) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_one_million_atoms_against_the_generator(dtype):
    """c4-size system: same number of pairs and the same distance multiset as the lattice-topology list"""
    from torchpme_b200.neighbors import neighbor_list

    pos, q, cell, idx_ref, d_ref = rocksalt(100, dtype=torch.float64, device="cuda", cutoff=6.0)
    idx, d, s = neighbor_list(pos.to(dtype), cell.to(dtype), 6.0)
    assert idx.shape[0] > 16_000_000
    n = pos.shape[0]
    ones = lambda k: torch.ones(k, device="cuda")  # noqa: E731
    if dtype == torch.float32:
        # fp32 positions up to 282 A carry ~3e-5 A of rounding and ~8e6 pairs per Angstrom sit around the
        # cutoff: a few hundred pairs may fall on either side of ANY threshold.  Everything else agrees.
        assert abs(idx.shape[0] - idx_ref.shape[0]) < 2000
        assert abs(float(d.double().sum()) - float(d_ref.sum())) < 1e-5 * float(d_ref.sum())
        deg = torch.zeros(n, device="cuda").index_add_(0, idx.reshape(-1), ones(2 * idx.shape[0]))
        deg_ref = torch.zeros(n, device="cuda").index_add_(0, idx_ref.reshape(-1), ones(2 * idx_ref.shape[0]))
        assert float((deg - deg_ref).abs().max()) <= 2 and float((deg != deg_ref).sum()) < 4000
        return
    eps = 1e-9
    sure = (d - 6.0).abs() > eps
    sure_ref = (d_ref - 6.0).abs() > eps
    assert int(sure.sum()) == int(sure_ref.sum())
    a = torch.sort(d[sure]).values
    b = torch.sort(d_ref[sure_ref]).values
    assert float((a - b).abs().max()) < 1e-10
    # every unordered pair once: degree sums agree atom by atom
    deg = torch.zeros(n, device="cuda").index_add_(0, idx[sure].reshape(-1), ones(2 * int(sure.sum())))
    deg_ref = torch.zeros(n, device="cuda").index_add_(0, idx_ref[sure_ref].reshape(-1), ones(2 * int(sure_ref.sum())))
    assert torch.equal(deg, deg_ref)


def test_int32_indices_and_known_distances():
    from torchpme_b200.neighbors import distances_from, neighbor_list

    pos, q, cell, _, _ = rocksalt(6, dtype=torch.float32, device="cuda", cutoff=5.0)
    idx64, d64, s64 = neighbor_list(pos, cell, 4.5)
    idx32, d32, s32 = neighbor_list(pos, cell, 4.5, index_dtype=torch.int32)
    assert idx32.dtype == torch.int32 and idx64.dtype == torch.int64
    # the order inside a bin depends on atomics: compare as sets
    def key(i, s):
        image = ((s[:, 0] + 1) * 9 + (s[:, 1] + 1) * 3 + (s[:, 2] + 1)).long()
        return torch.sort((i[:, 0].long() * 4096 + i[:, 1].long()) * 27 + image).values

    assert torch.equal(key(idx64, s64), key(idx32, s32))
    p = pos.clone().requires_grad_(True)
    a = distances_from(p, cell, idx32, s32, known_distances=d32)
    b = distances_from(p, cell, idx32, s32)
    assert rel_err(a.detach(), b.detach()) < 1e-6
    (ga,) = torch.autograd.grad(a.square().sum(), p)
    (gb,) = torch.autograd.grad(b.square().sum(), p)
    assert rel_err(ga, gb) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("full", [False, True], ids=["half", "full"])
def test_fixed_capacity_list_through_the_calculator(dtype, full):
    """DeviceNeighborList (no host sync, padded buffers, pair count on the device): same potentials and
    forces as the exact-size list, padding entries never touched, overflow detected"""
    import torchpme_b200 as tp
    from torchpme_b200.neighbors import DeviceNeighborList, distances_from, neighbor_list

    pos, q, cell, _, _ = rocksalt(8, dtype=dtype, device="cuda", cutoff=5.0)
    cutoff = 5.0
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.1).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14,
                            full_neighbor_list=full)

    def step(idx, shifts, known):
        p = pos.clone().requires_grad_(True)
        d = distances_from(p, cell, idx, shifts, known_distances=known)
        V = calc(q, cell, p, idx, d)
        (g,) = torch.autograd.grad(V, p, grad_outputs=q)
        return V.detach(), g

    idx, d, s = neighbor_list(pos, cell, cutoff, full_neighbor_list=full)
    V_ref, F_ref = step(idx, s, None)
    nl = DeviceNeighborList(pos.shape[0], cell, cutoff, capacity=idx.shape[0] + 777, dtype=dtype,
                            full_neighbor_list=full)
    i2, d2, s2 = nl.build(pos)
    assert int(nl.n_pairs) == idx.shape[0] and not nl.overflowed()
    assert i2.shape[0] == idx.shape[0] + 777 and bool((i2[idx.shape[0]:] == 0).all())
    V, F = step(i2, s2, d2)
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    assert rel_err(V, V_ref) < tol and rel_err(F, F_ref) < tol
    # against the oracle's pair list too (independent of the device list)
    o_idx, o_d, o_s = oracle.neighbor_list(pos.double().cpu().numpy(), cell.double().cpu().numpy(), cutoff, full=full)
    V_o, F_o = step(torch.tensor(o_idx, device="cuda"), torch.tensor(o_s, device="cuda", dtype=torch.int32), None)
    assert rel_err(V, V_o) < tol and rel_err(F, F_o) < tol
    small = DeviceNeighborList(pos.shape[0], cell, cutoff, capacity=idx.shape[0] // 2, dtype=dtype, full_neighbor_list=full)
    small.build(pos)
    assert small.overflowed() and int(small.n_pairs) == idx.shape[0]


def test_graphed_positions_step_matches_the_list_based_step():
    """positions-only step captured as one CUDA graph (list build + calculator + backward), with and without
    host I/O, replayed for moved atoms"""
    import torchpme_b200 as tp
    from torchpme_b200.neighbors import distances_from, neighbor_list

    pos, q, cell, _, _ = rocksalt(10, dtype=torch.float64, device="cuda", cutoff=5.0)
    calc = tp.P3MCalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14)

    def eager(positions):
        idx, d, s = neighbor_list(positions, cell, 5.0)
        p = positions.clone().requires_grad_(True)
        V = calc(q, cell, p, idx, distances_from(p, cell, idx, s))
        E = (V * q).sum()
        (g,) = torch.autograd.grad(E, p)
        return E.detach(), g

    for host_io in (False, True):
        step = tp.GraphedPositionsStep(calc, q, cell, pos, cutoff=5.0, host_io=host_io)
        gen = torch.Generator("cuda").manual_seed(5)
        for _ in range(3):
            moved = pos + 0.2 * torch.randn(pos.shape, dtype=pos.dtype, device="cuda", generator=gen)
            if host_io:
                step.host["positions"].copy_(moved)
                step.replay()
                torch.cuda.synchronize()
                E, g = step.host["energy"].cuda(), step.host["grad_positions"].cuda()
            else:
                E, g = step(positions=moved)
            assert not step.overflowed()
            E_ref, g_ref = eager(moved)
            assert rel_err(E, E_ref) < 1e-11 and rel_err(g, g_ref) < 1e-10


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("calc_name", ["PMECalculator", "P3MCalculator"])
def test_forward_from_pairs_equals_the_composition(calc_name, dtype):
    """calculator.forward_from_pairs(q, cell, r, idx, S) == calculator(q, cell, r, idx, distances_from(r, cell, idx, S)),
    values and the complete position gradient (mesh + real space), with / without known distances, on an
    exact-size and on a fixed-capacity list, for a generic upstream gradient"""
    import torchpme_b200 as tp
    from torchpme_b200.neighbors import DeviceNeighborList, distances_from, neighbor_list

    pos, q, cell, _, _ = rocksalt(8, dtype=dtype, device="cuda", cutoff=5.0)
    calc = getattr(tp, calc_name)(tp.CoulombPotential(smearing=1.1).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14)
    idx, d, s = neighbor_list(pos, cell, 5.0, index_dtype=torch.int32)
    w = torch.randn(q.shape, dtype=dtype, device="cuda", generator=torch.Generator("cuda").manual_seed(3))

    p0 = pos.clone().requires_grad_(True)
    q0 = q.clone().requires_grad_(True)
    V0 = calc(q0, cell, p0, idx, distances_from(p0, cell, idx, s))
    gp0, gq0 = torch.autograd.grad(V0, (p0, q0), grad_outputs=w)
    tol = 1e-11 if dtype == torch.float64 else 3e-5
    nl = DeviceNeighborList(pos.shape[0], cell, 5.0, capacity=idx.shape[0] + 500, dtype=dtype)
    ci, cd, cs = nl.build(pos)
    for (i_, s_, known) in ((idx, s, None), (idx, s, d), (ci, cs, cd), (ci, cs, None)):
        p1 = pos.clone().requires_grad_(True)
        q1 = q.clone().requires_grad_(True)
        V1 = calc.forward_from_pairs(q1, cell, p1, i_, s_, known_distances=known)
        gp1, gq1 = torch.autograd.grad(V1, (p1, q1), grad_outputs=w)
        assert rel_err(V1.detach(), V0.detach()) < tol
        assert rel_err(gp1, gp0) < tol and rel_err(gq1, gq0) < tol
    # forces only (no charge gradient) and charges only
    p1 = pos.clone().requires_grad_(True)
    (gp1,) = torch.autograd.grad(calc.forward_from_pairs(q, cell, p1, idx, s), p1, grad_outputs=w)
    assert rel_err(gp1, gp0) < tol
    q1 = q.clone().requires_grad_(True)
    (gq1,) = torch.autograd.grad(calc.forward_from_pairs(q1, cell, pos, idx, s), q1, grad_outputs=w)
    assert rel_err(gq1, gq0) < tol


@pytest.mark.parametrize("full", [False, True], ids=["half", "full"])
@pytest.mark.parametrize("index_dtype", [torch.int64, torch.int32])
def test_crowded_atoms_take_the_second_search(full, index_dtype):
    """more partners per atom than the shared-memory hit list holds (32): the rest is written by a second walk;
    dense system, ~100 (half: ~50) neighbors per atom, triclinic cell, against the brute-force oracle"""
    from torchpme_b200.neighbors import DeviceNeighborList, neighbor_list

    rng = np.random.default_rng(5)
    cell = np.array([[8.0, 0.0, 0.0], [1.5, 7.5, 0.0], [-1.0, 0.8, 8.5]])
    pos = rng.uniform(0, 1, (200, 3)) @ cell
    cutoff = 4.0
    idx, d, s = neighbor_list(torch.tensor(pos, device="cuda"), torch.tensor(cell, device="cuda"), cutoff,
                              full_neighbor_list=full, index_dtype=index_dtype)
    o_idx, o_d, o_s = oracle.neighbor_list(pos, cell, cutoff, full=full)
    assert idx.shape[0] == o_idx.shape[0] and idx.shape[0] / 200 > (64 if full else 32)
    if full:
        got = np.concatenate([idx.cpu().numpy().astype(np.int64), s.cpu().numpy()], axis=1)
        want = np.concatenate([o_idx, o_s], axis=1)
        assert np.array_equal(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])
    else:
        rows, dist = _canonical(idx.cpu().numpy(), d.cpu().numpy(), s.cpu().numpy())
        o_rows, o_dist = _canonical(o_idx, o_d, o_s)
        assert np.array_equal(rows, o_rows) and np.abs(dist - o_dist).max() < 1e-12
    # a capacity that ends in the middle of a CTA's range: the pairs that fit are valid pairs, the count is complete
    cap = idx.shape[0] // 3
    nl = DeviceNeighborList(200, torch.tensor(cell, device="cuda"), cutoff, capacity=cap, dtype=torch.float64,
                            full_neighbor_list=full, index_dtype=index_dtype)
    ci, cd, cs = nl.build(torch.tensor(pos, device="cuda"))
    assert nl.overflowed() and int(nl.n_pairs) == idx.shape[0]
    key = lambda i, sh: set(map(tuple, np.concatenate([i.cpu().numpy().astype(np.int64), sh.cpu().numpy()], axis=1)))  # noqa: E731
    assert key(ci, cs) <= key(idx, s) and len(key(ci, cs)) == cap


def test_empty_and_tiny_systems():
    from torchpme_b200.neighbors import neighbor_list

    cell = torch.eye(3, dtype=torch.float64, device="cuda") * 5.0
    idx, d, s = neighbor_list(torch.empty((0, 3), dtype=torch.float64, device="cuda"), cell, 2.0)
    assert idx.shape == (0, 2) and d.shape == (0,) and s.shape == (0, 3)
    one = torch.tensor([[0.3, 0.4, 0.5]], dtype=torch.float64, device="cuda")
    idx, d, s = neighbor_list(one, cell, 2.0)                     # no image within the cutoff
    assert idx.shape[0] == 0
    idx, d, s = neighbor_list(one, cell, 5.5)                     # the six nearest self images: three in a half list
    assert idx.shape[0] == 3 and bool((idx == 0).all()) and torch.allclose(d, torch.full_like(d, 5.0))
    idx, d, s = neighbor_list(one, cell, 5.5, full_neighbor_list=True)
    assert idx.shape[0] == 6
