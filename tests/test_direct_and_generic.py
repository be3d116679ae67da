"""
Calculator routes next to the fused PME / P3M fast path:

* the direct (no smearing) ``Calculator`` on point-charge molecules with exact answers, as
  tests/calculators/test_values_direct.py:162-208 of the reference (rtol 1e-14 there) -- oracle
  on the CPU, CUDA path on the GPU;
* potentials the kernels do not know (here: a subclass of ``CoulombPotential``) go through the
  generic routes -- per-pair values from the potential's torch code, filter table from
  ``lr_from_k_sq`` -- and must agree with the in-kernel evaluation of the same potential;
* ``pair_mask`` equals removing the masked pairs from the neighbor list.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err, rocksalt
from molecules import all_pairs, molecule, rotations
from oracle import pme_oracle as oracle

CASES = [(m, s, f, r, full) for m in ("dimer", "triangle", "square", "tetrahedron")
         for s in ("alternating", "positive", "negative") for f in (0.079, 1.0, 5.54)
         for r in range(5) for full in (True, False)]


def test_oracle_direct_route_exact_molecules():
    spec = oracle.PotentialSpec("coulomb", None)
    for name, sign, scale, rot, full in CASES:
        pos, q, exact = molecule(name, sign)
        pos = scale * (pos @ rotations()[rot])
        idx, d = all_pairs(pos, full)
        v = oracle.compute_rspace(spec, q, idx, d, full_neighbor_list=full)
        np.testing.assert_allclose(v, exact / (2 * scale), rtol=1e-13, atol=2e-15)


def test_oracle_pair_mask_removes_pairs():
    rng = np.random.default_rng(3)
    pos = rng.random((12, 3)) * 3
    q = rng.standard_normal((12, 2))
    idx, d = all_pairs(pos, False)
    mask = rng.random(len(d)) < 0.6
    spec = oracle.PotentialSpec("coulomb", 0.7)
    a = oracle.compute_rspace(spec, q, idx, d, pair_mask=mask)
    b = oracle.compute_rspace(spec, q, idx[mask], d[mask])
    np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-15)


@pytest.mark.gpu
def test_direct_calculator_exact_molecules():
    import torchpme_b200 as tp

    dev = "cuda"
    for full in (True, False):
        calc = tp.Calculator(tp.CoulombPotential().to(dev), full_neighbor_list=full)
        for name, sign, scale, rot, f in CASES:
            if f != full or rot not in (0, 2, 4):
                continue
            pos, q, exact = molecule(name, sign)
            pos = scale * (pos @ rotations()[rot])
            idx, d = all_pairs(pos, full)
            V = calc.forward(torch.tensor(q, device=dev), torch.eye(3, dtype=torch.float64, device=dev),
                             torch.tensor(pos, device=dev), torch.tensor(idx, device=dev),
                             torch.tensor(d, device=dev))
            np.testing.assert_allclose(V.cpu().numpy(), exact / (2 * scale), rtol=1e-13, atol=2e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["P3M", "PME"])
def test_generic_routes_match_in_kernel_potential(method):
    import torchpme_b200 as tp

    class TorchCoulomb(tp.CoulombPotential):
        """same maths, but unknown to the kernels: forces the per-pair-value and filter-table routes"""

    dev, dt = "cuda", torch.float64
    pos, q, cell, idx, d = rocksalt(6, dtype=dt, device=dev, cutoff=5.0)
    q = torch.cat([q, 0.3 * q + 0.1], dim=1)
    cls = tp.P3MCalculator if method == "P3M" else tp.PMECalculator
    L = float(cell[0, 0])
    gen = torch.Generator().manual_seed(11)
    gout = torch.randn(q.shape, generator=gen, dtype=dt).to(dev)
    mask = (torch.rand(idx.shape[0], generator=gen) < 0.8).to(dev)
    results = []
    for pot in (tp.CoulombPotential(smearing=1.1, prefactor=1.7), TorchCoulomb(smearing=1.1, prefactor=1.7)):
        calc = cls(pot.to(dev), mesh_spacing=L / 14)
        assert (pot._native_descriptor() is None) == isinstance(pot, TorchCoulomb)
        out = []
        for pair_mask in (None, mask):
            p = pos.clone().requires_grad_(True)
            qq = q.clone().requires_grad_(True)
            dd = d.clone().requires_grad_(True)
            V = calc(qq, cell, p, idx, dd, pair_mask=pair_mask)
            (V * gout).sum().backward()
            out += [V.detach(), p.grad, qq.grad, dd.grad]
        results.append(out)
    for a, b in zip(*results):
        assert rel_err(b, a) < 1e-9
    # pair_mask == dropping the masked pairs
    calc = cls(tp.CoulombPotential(smearing=1.1, prefactor=1.7).to(dev), mesh_spacing=L / 14)
    with torch.no_grad():
        Vm = calc(q, cell, pos, idx, d, pair_mask=mask)
        Vd = calc(q, cell, pos, idx[mask], d[mask])
    assert rel_err(Vm, Vd) < 1e-12
