"""
CPU tensors (device dispatch to torchpme_b200._cpu, the package's own torch formulation) against the
golden tensors of the unmodified reference: BASELINE config c1 (CsCl, PMECalculator, fp64, CPU) and
every other calculator golden incl. the 2-D periodic ones, values and all gradients; the reference's
workflow expectations (dtype / device preserved, `examples/basic-usage.py` numbers); the inspection
attributes of MeshInterpolator (lib/mesh_interpolator.py:65-79); double backward.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, b200_potential, case_arrays, load_calculator_cases

CASES, DATA = load_calculator_cases()
with open(os.path.join(GOLDEN, "periodic_cases.json")) as f:
    P_CASES = json.load(f)
P_DATA = np.load(os.path.join(GOLDEN, "periodic_cases.npz"))
ALL = [(c, DATA) for c in CASES] + [(c, P_DATA) for c in P_CASES]


def _norm(ref, natural):
    return max(float(np.abs(ref).max()), 1e-2 * natural)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("case, data", ALL, ids=[c["name"] for c, _ in ALL])
def test_cpu_calculators_match_reference_golden(case, data, dtype):
    import torchpme_b200 as tp

    g = case_arrays(data, case["name"])
    q = torch.tensor(g["charges"], dtype=dtype, requires_grad=True)
    cell = torch.tensor(g["cell"], dtype=dtype, requires_grad=True)
    pos = torch.tensor(g["positions"], dtype=dtype, requires_grad=True)
    d = torch.tensor(g["neighbor_distances"], dtype=dtype, requires_grad=True)
    idx = torch.tensor(g["neighbor_indices"])
    cls = tp.PMECalculator if case["calc"] == "pme" else tp.P3MCalculator
    calc = cls(b200_potential(tp, case["pot"], dtype=dtype), mesh_spacing=case["mesh_spacing"],
               interpolation_nodes=case["nodes"], full_neighbor_list=case["full"])
    periodic = torch.tensor(case["periodic"]) if "periodic" in case else None
    V = calc.forward(q, cell, pos, idx, d, periodic=periodic)
    assert V.dtype == dtype and V.device.type == "cpu"          # tests/calculators/test_workflow.py:112-123
    (V * torch.tensor(g["grad_out"], dtype=dtype)).sum().backward()
    tol = 1e-9 if dtype == torch.float64 else 1e-3
    scale = float(np.abs(g["V"]).max())
    assert np.abs(V.detach().numpy() - g["V"]).max() / scale < tol
    for name, t in (("dq", q), ("dd", d), ("dpos", pos), ("dcell", cell)):
        assert np.abs(t.grad.numpy() - g[name]).max() / _norm(g[name], scale) < tol, name


def test_c1_cscl_madelung_on_cpu():
    """BASELINE config c1 / examples/basic-usage.py: CsCl, PMECalculator, fp64, CPU -> Madelung constant"""
    import torchpme_b200 as tp

    pos = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], dtype=torch.float64)
    q = torch.tensor([[1.0], [-1.0]], dtype=torch.float64)
    cell = torch.eye(3, dtype=torch.float64)
    from oracle import pme_oracle as oracle   # test utility: the neighbor list only
    idx, d, _ = oracle.neighbor_list(pos.numpy(), cell.numpy(), 1.0)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=0.2), mesh_spacing=0.05).to(torch.float64)
    V = calc(q, cell, pos, torch.tensor(idx), torch.tensor(d))
    madelung = -float((V * q).sum())          # unit cell edge 1: tests/helpers.py:34 (2.0353610945), rtol of the reference's test
    assert abs(madelung - 2.0353610945) < 9e-4 * 2.0353610945, madelung


def test_mesh_interpolator_inspection_attributes():
    """interpolation_weights, x/y/z_indices, x/y/z_shifts as the reference exposes them"""
    import torchpme_b200 as tp

    rng = np.random.default_rng(0)
    cell = torch.tensor(np.eye(3) * 5.0 + rng.uniform(-0.3, 0.3, (3, 3)))
    pos = torch.tensor(rng.uniform(-2, 7, (11, 3)))
    ns = torch.tensor([6, 7, 8])
    for method, nodes in (("P3M", 3), ("Lagrange", 4), ("P3M", 5)):
        mi = tp.lib.MeshInterpolator(cell, ns, nodes, method)
        mi.compute_weights(pos)
        w = mi.interpolation_weights
        assert w.shape == (nodes, 11, 3)
        assert torch.allclose(w.sum(0), torch.ones(11, 3, dtype=torch.float64), atol=1e-12)   # partition of unity
        for name, n_axis in (("x", 6), ("y", 7), ("z", 8)):
            ind, sh = getattr(mi, f"{name}_indices"), getattr(mi, f"{name}_shifts")
            assert ind.shape == (nodes ** 3, 11) and sh.shape == (nodes ** 3,)
            assert int(ind.min()) >= 0 and int(ind.max()) < n_axis
        # the attributes reproduce points_to_mesh exactly (mesh_interpolator.py:411-426)
        weights = torch.tensor(rng.normal(size=(11, 1)))
        rho = mi.points_to_mesh(weights)
        xs, ys, zs = mi.x_shifts, mi.y_shifts, mi.z_shifts
        ref = torch.zeros(1, 6, 7, 8, dtype=torch.float64)
        contrib = weights[:, 0] * w[xs, :, 0] * w[ys, :, 1] * w[zs, :, 2]
        ref[0].index_put_((mi.x_indices, mi.y_indices, mi.z_indices), contrib, accumulate=True)
        assert torch.allclose(rho, ref, atol=1e-13)
    try:
        from _reference_import import available, import_reference
    except ImportError:
        return
    if available():
        ref = import_reference()
        rmi = ref.lib.MeshInterpolator(cell, ns, 4, "Lagrange")
        rmi.compute_weights(pos)
        mi = tp.lib.MeshInterpolator(cell, ns, 4, "Lagrange")
        mi.compute_weights(pos)
        assert torch.allclose(mi.interpolation_weights, rmi.interpolation_weights, atol=1e-13)
        for name in ("x_indices", "y_indices", "z_indices", "x_shifts", "y_shifts", "z_shifts"):
            assert torch.equal(getattr(mi, name), getattr(rmi, name)), name


def test_double_backward_on_cpu():
    """force matching needs create_graph=True through the calculator (the reference supports it)"""
    import torchpme_b200 as tp

    g = case_arrays(DATA, "rand_p3m_n4_coulomb")
    dt = torch.float64
    q = torch.tensor(g["charges"], dtype=dt)
    cell = torch.tensor(g["cell"], dtype=dt)
    pos = torch.tensor(g["positions"], dtype=dt, requires_grad=True)
    idx = torch.tensor(g["neighbor_indices"])
    pot = tp.CoulombPotential(smearing=0.9, prefactor=1.7).to(dt)
    pot.smearing.requires_grad_(True)        # a potential parameter as a leaf
    calc = tp.P3MCalculator(pot, mesh_spacing=0.8, full_neighbor_list=False)
    d = torch.tensor(g["neighbor_distances"], dtype=dt)
    V = calc(q, cell, pos, idx, d)
    (forces,) = torch.autograd.grad((V * q).sum(), pos, create_graph=True)
    loss = (forces ** 2).sum()
    (g_smearing,) = torch.autograd.grad(loss, pot.smearing)
    assert torch.isfinite(g_smearing) and float(g_smearing.abs()) > 0


def test_torch_compile_on_cpu():
    """tests/calculators/test_workflow.py:146-153 on the CPU path (Dynamo + AOT autograd; eager kernels)"""
    import torchpme_b200 as tp

    pos = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], dtype=torch.float64)
    q = torch.tensor([[1.0], [-1.0]], dtype=torch.float64)
    cell = torch.eye(3, dtype=torch.float64)
    idx, d = torch.tensor([[0, 1]]), torch.tensor([0.8660254], dtype=torch.float64)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=0.2), mesh_spacing=0.1).to(torch.float64)
    ref = calc(q, cell, pos, idx, d)
    torch._dynamo.reset()
    compiled = torch.compile(calc, backend="aot_eager")
    p = pos.clone().requires_grad_(True)
    out = compiled(q, cell, p, idx, d)
    assert type(out) is torch.Tensor and torch.allclose(out, ref, atol=1e-12)
    out.sum().backward()
    assert torch.isfinite(p.grad).all()


def test_forward_from_pairs_on_cpu_is_the_composition():
    """pair list as indices + image shifts on CPU tensors: distances_from + forward, gradients through both"""
    import torchpme_b200 as tp
    from oracle import pme_oracle as oracle
    from torchpme_b200.neighbors import distances_from

    rng = np.random.default_rng(0)
    cell_np = np.array([[6.0, 0.0, 0.0], [0.8, 5.5, 0.0], [-0.4, 0.6, 6.2]])
    pos_np = rng.uniform(0, 1, (24, 3)) @ cell_np
    idx_np, d_np, s_np = oracle.neighbor_list(pos_np, cell_np, 3.0)
    q = torch.tensor(rng.normal(size=(24, 1)))
    cell = torch.tensor(cell_np)
    idx, shifts = torch.tensor(idx_np), torch.tensor(s_np, dtype=torch.int32)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=0.8), mesh_spacing=0.4)
    p0 = torch.tensor(pos_np, requires_grad=True)
    d = distances_from(p0, cell, idx, shifts)
    np.testing.assert_allclose(d.detach().numpy(), d_np, rtol=1e-12)
    V0 = calc(q, cell, p0, idx, d)
    (g0,) = torch.autograd.grad((V0 * q).sum(), p0)
    p1 = torch.tensor(pos_np, requires_grad=True)
    V1 = calc.forward_from_pairs(q, cell, p1, idx, shifts)
    (g1,) = torch.autograd.grad((V1 * q).sum(), p1)
    np.testing.assert_allclose(V1.detach().numpy(), V0.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(g1.numpy(), g0.numpy(), rtol=1e-10, atol=1e-12)
    # the complete force (mesh + real space) sums to ~0 over the atoms (up to the mesh discretisation error)
    assert float(g1.sum(0).abs().max()) < 1e-2 * float(g1.abs().max())
