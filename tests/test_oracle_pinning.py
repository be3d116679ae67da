"""
Pin the numpy oracle (oracle/pme_oracle.py) to the reference:
  (i)  golden tensors produced by the unmodified reference (tests/golden/make_golden.py),
  (ii) the reference's own known answers: Madelung constants
       (tests/calculators/test_values_ewald.py:65-152, rtol 9e-4 for pme/p3m), GROMACS SPME
       energies / forces (test_values_ewald.py:223-356, rtol 1e-4 / 5e-3), closed forms of the
       short-range potentials (tests/test_potentials.py:86-182), E1 against scipy
       (tests/lib/test_math.py:12-31),
  (iii) when /root/reference is present (build container only): a live comparison.
"""
import json
import os

import numpy as np
import pytest
import scipy.special

import crystals
from helpers import GOLDEN, case_arrays, load_calculator_cases, oracle_potential, rel_err
from oracle import pme_oracle as oracle

CASES, DATA = load_calculator_cases()


def _method(case):
    return "Lagrange" if case["calc"] == "pme" else "P3M"


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    g = case_arrays(DATA, case["name"])
    pot = oracle_potential(case["pot"])
    args = (pot, g["charges"], g["cell"], g["positions"], g["neighbor_indices"], g["neighbor_distances"],
            case["mesh_spacing"], case["nodes"], _method(case), case["full"])
    assert tuple(oracle.get_ns_mesh(g["cell"], case["mesh_spacing"])) == tuple(g["ns_mesh"])
    V = oracle.calculator_forward(*args)
    assert rel_err(V, g["V"]) < 1e-12
    step = oracle.calculator_step(*args, grad_out=g["grad_out"])
    scale = max(np.abs(g["V"]).max(), 1e-30)
    assert rel_err(step["V"], g["V"]) < 1e-12
    assert rel_err(step["dq"], g["dq"]) < 1e-11
    assert rel_err(step["dd"], g["dd"]) < 1e-8  # exclusion variant uses a finite difference
    assert np.abs(step["dpos"] - g["dpos"]).max() / max(np.abs(g["dpos"]).max(), scale) < 1e-11
    energy = oracle.calculator_step(*args)
    assert np.abs(energy["dpos"] - g["dpos_energy"]).max() / max(np.abs(g["dpos_energy"]).max(), scale) < 1e-11
    assert rel_err(energy["dd"], g["dd_energy"]) < 1e-8


def test_oracle_blocks_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, "block_cases.npz"))
    cell, pos, w, ns, mesh_in = g["cell"], g["positions"], g["weights"], g["ns"], g["mesh_in"]
    for method, nodes_list in (("P3M", (1, 2, 3, 4, 5)), ("Lagrange", (3, 4, 5, 6, 7))):
        for nodes in nodes_list:
            key = f"{method}_{nodes}"
            assert rel_err(oracle.points_to_mesh(w, pos, cell, ns, nodes, method), g[key + "_rho"]) < 1e-13
            vals, dvals = oracle.mesh_to_points(mesh_in, pos, cell, nodes, method, gradient=True)
            assert rel_err(vals, g[key + "_vals"]) < 1e-13
            dpos = np.einsum("ic,icd->id", g[key + "_g"], dvals)
            if nodes > 1:
                assert rel_err(dpos, g[key + "_dpos"]) < 1e-12
    kv = oracle.kvectors_for_mesh(cell, ns)
    k_sq = np.linalg.norm(kv, axis=3) ** 2
    coul = oracle.PotentialSpec("coulomb", 0.8)
    assert rel_err(coul.lr_from_k_sq(k_sq), g["kfilter_coulomb"]) < 1e-13
    assert rel_err(oracle.p3m_influence(kv, cell, ns, 4) * coul.lr_from_k_sq(k_sq), g["kfilter_p3m"]) < 1e-13
    for p in range(1, 7):
        ipl = oracle.PotentialSpec("ipl", 0.8, exponent=p)
        assert rel_err(ipl.lr_from_k_sq(k_sq), g[f"kfilter_ipl{p}"]) < 1e-12
    for fn, inn in (("ortho", "ortho"), ("backward", "forward"), ("forward", "backward"), ("backward", "backward")):
        out = oracle.kspace_filter(mesh_in, g["kfilter_coulomb"], fn, inn)
        assert rel_err(out, g[f"filter_{fn}_{inn}"]) < 1e-12


@pytest.mark.parametrize("calc", ["pme", "p3m"])
@pytest.mark.parametrize("scale", [1 / 2.0353610, 1.0, 3.4951291])
@pytest.mark.parametrize("name", crystals.NAMES)
def test_oracle_madelung(name, scale, calc):
    """reference: tests/calculators/test_values_ewald.py:65-152 (same hyper-parameters, rtol 9e-4)"""
    pos, q, cell, madelung, units = crystals.get(name, scale)
    cutoff = 2 * scale
    smearing = cutoff / 5.0
    idx, d, _ = oracle.neighbor_list(pos, cell, cutoff, full=(calc == "pme"))
    pot = oracle.PotentialSpec("ipl" if calc == "pme" else "coulomb", smearing, exponent=1)
    V = oracle.calculator_forward(pot, q, cell, pos, idx, d, smearing / 8, 4,
                                  "Lagrange" if calc == "pme" else "P3M", full_neighbor_list=(calc == "pme"))
    assert abs(-(V * q).sum() / units - madelung) / madelung < 9e-4


@pytest.mark.parametrize("calc", ["pme", "p3m"])
@pytest.mark.parametrize("scale", [0.43, 1.33])
@pytest.mark.parametrize("frame", [0, 1])
def test_oracle_gromacs_energy_forces(frame, scale, calc):
    """reference: tests/calculators/test_values_ewald.py:223-314 (energy rtol 1e-4, forces 5e-3)"""
    with open(os.path.join(GOLDEN, "gromacs_frames.json")) as f:
        fr = json.load(f)["frames"][frame]
    pos = scale * np.array(fr["positions"])
    cell = scale * np.array(fr["cell"])
    q = np.array(fr["charges"]).reshape(-1, 1)
    cutoff = 5.54 * scale
    smearing = cutoff / 6.0
    idx, d, shifts = oracle.neighbor_list(pos, cell, cutoff, full=False)
    pot = oracle.PotentialSpec("coulomb", smearing, prefactor=14.399645478425667)
    method = "Lagrange" if calc == "pme" else "P3M"
    step = oracle.calculator_step(pot, q, cell, pos, idx, d, smearing / 8.0, 4, method)
    energy = (step["V"] * q).sum()
    assert abs(energy - fr["energy"] / scale) / abs(fr["energy"] / scale) < 1e-4
    # total force = k-space part (dpos) + real-space chain rule through the distances
    vec = pos[idx[:, 1]] + shifts @ cell - pos[idx[:, 0]]
    unit = vec / np.linalg.norm(vec, axis=1, keepdims=True)
    grad = step["dpos"].copy()
    np.add.at(grad, idx[:, 1], step["dd"][:, None] * unit)
    np.add.at(grad, idx[:, 0], -step["dd"][:, None] * unit)
    forces_ref = np.array(fr["forces"]) / scale**2
    assert np.abs(-grad - forces_ref).max() / np.abs(forces_ref).max() < 5e-3


def test_oracle_short_range_closed_forms():
    """SR + LR = 1/r^p and the p = 1, 2, 3 closed forms (reference tests/test_potentials.py:61-111)"""
    d = np.linspace(0.3, 6.0, 200)
    s = 1.1
    x = d**2 / (2 * s**2)
    expect = {1: scipy.special.erfc(np.sqrt(x)) / d, 2: np.exp(-x) / d**2,
              3: (scipy.special.erfc(np.sqrt(x)) + 2 * np.sqrt(x / np.pi) * np.exp(-x)) / d**3}
    for p in range(1, 7):
        pot = oracle.PotentialSpec("ipl", s, exponent=p)
        np.testing.assert_allclose(pot.sr_from_dist(d) + pot.lr_from_dist(d), d ** (-float(p)), rtol=1e-13)
        np.testing.assert_allclose(pot.sr_from_dist_closed(d), pot.sr_from_dist(d), rtol=1e-9, atol=1e-15)
        if p in expect:
            np.testing.assert_allclose(pot.sr_from_dist(d), expect[p], rtol=1e-10, atol=1e-15)
        v, dv = pot.sr_from_dist_closed(d, deriv=True)
        h = 1e-6
        fd = (pot.sr_from_dist_closed(d + h) - pot.sr_from_dist_closed(d - h)) / (2 * h)
        np.testing.assert_allclose(dv, fd, rtol=1e-6, atol=1e-12)
    coul = oracle.PotentialSpec("coulomb", s)
    np.testing.assert_allclose(coul.sr_from_dist(d), expect[1], rtol=1e-10, atol=1e-15)


def test_oracle_exp1_matches_scipy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e-4, 1, 5000), rng.uniform(1, 60, 5000)])
    np.testing.assert_allclose(oracle.exp1(x), scipy.special.exp1(x), rtol=2e-14)


def test_oracle_cell_gradient_by_finite_difference_matches_golden():
    """dcell of the golden set (reference autograd) against central differences of the oracle."""
    case = next(c for c in CASES if c["name"] == "rand_p3m_n4_coulomb")
    g = case_arrays(DATA, case["name"])
    pot = oracle_potential(case["pot"])
    ns = tuple(g["ns_mesh"])

    def loss(cell):
        V = oracle.calculator_forward(pot, g["charges"], cell, g["positions"], g["neighbor_indices"],
                                      g["neighbor_distances"], case["mesh_spacing"], case["nodes"],
                                      _method(case), case["full"], ns=ns)
        return (V * g["grad_out"]).sum()

    fd = np.zeros((3, 3))
    h = 1e-6
    for a in range(3):
        for b in range(3):
            e = np.zeros((3, 3)); e[a, b] = h
            fd[a, b] = (loss(g["cell"] + e) - loss(g["cell"] - e)) / (2 * h)
    assert np.abs(fd - g["dcell"]).max() / np.abs(g["dcell"]).max() < 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/torchpme"), reason="reference tree only in the build container")
def test_oracle_against_live_reference():
    import torch
    from _reference_import import import_reference

    tp = import_reference()
    rng = np.random.default_rng(11)
    cell = np.eye(3) * 7.0 + rng.uniform(-0.6, 0.6, (3, 3))
    pos = rng.uniform(0, 1, (50, 3)) @ cell
    q = rng.normal(size=(50, 1))
    idx, d, _ = oracle.neighbor_list(pos, cell, 3.0)
    for cls, method, nodes in ((tp.PMECalculator, "Lagrange", 5), (tp.P3MCalculator, "P3M", 3)):
        calc = cls(tp.CoulombPotential(smearing=0.9), mesh_spacing=0.6, interpolation_nodes=nodes).to(torch.float64)
        ref = calc(torch.tensor(q), torch.tensor(cell), torch.tensor(pos), torch.tensor(idx), torch.tensor(d)).numpy()
        mine = oracle.calculator_forward(oracle.PotentialSpec("coulomb", 0.9), q, cell, pos, idx, d, 0.6, nodes, method)
        assert rel_err(mine, ref) < 1e-12


# --------------------------------------------------------------------------------------
# interpolator invariants the reference asserts (tests/lib/test_mesh_interpolator.py:17-328)
# --------------------------------------------------------------------------------------
STENCILS = [("P3M", n) for n in (1, 2, 3, 4, 5)] + [("Lagrange", n) for n in (3, 4, 5, 6, 7)]


@pytest.mark.parametrize("method, nodes", STENCILS)
def test_oracle_interpolator_invariants(method, nodes):
    rng = np.random.default_rng(100 + nodes)
    # charge conservation, cubic cell, odd and even mesh sizes (:23-58)
    for n_mesh in (19, 20, 25):
        cell = np.eye(3) * 6.28318530717
        pos = rng.random((8, 3)) * 6.28318530717
        w = 3 * rng.standard_normal((8, 5))
        mesh = oracle.points_to_mesh(w, pos, cell, (n_mesh,) * 3, nodes, method)
        np.testing.assert_allclose(mesh.sum(axis=(1, 2, 3)), w.sum(0), rtol=1e-12, atol=1e-12)
    # charge conservation, triclinic cell, anisotropic mesh, atoms outside the cell (:62-99)
    cell = rng.standard_normal((3, 3)) * 2.718281828
    pos = (rng.random((11, 3)) * 3 - 1) @ cell
    w = 3 * rng.standard_normal((11, 2))
    ns = (11, 14, 17)
    mesh = oracle.points_to_mesh(w, pos, cell, ns, nodes, method)
    np.testing.assert_allclose(mesh.sum(axis=(1, 2, 3)), w.sum(0), rtol=1e-12, atol=1e-12)
    # total mass: interpolating a constant mesh returns the constant (:236-278)
    const = np.full((2,) + ns, 0.37)
    const[1] = -1.9
    vals, dvals = oracle.mesh_to_points(const, pos, cell, nodes, method, gradient=True)
    np.testing.assert_allclose(vals, np.broadcast_to([0.37, -1.9], vals.shape), rtol=1e-12)
    np.testing.assert_allclose(dvals, 0.0, atol=1e-10)
    # spread and gather are adjoint: <gather(m), w> == <m, spread(w)>
    m = rng.standard_normal((2,) + ns)
    lhs = (oracle.mesh_to_points(m, pos, cell, nodes, method) * w).sum()
    np.testing.assert_allclose(lhs, (m * mesh).sum(), rtol=1e-11)
    # derivative of the interpolated value by central differences (:282-328); skip points that sit
    # within the step of a stencil switch (P3M n >= 2 is C0 or better, Lagrange is only C0 there)
    if nodes > 1:
        eps = 1e-6
        u = oracle.mesh_coordinates(pos, cell, ns)
        frac = u - np.floor(u) if nodes % 2 == 0 else u - np.rint(u) + 0.5
        safe = np.all((frac > 1e-3) & (frac < 1 - 1e-3), axis=1)
        _, grad = oracle.mesh_to_points(m, pos, cell, nodes, method, gradient=True)
        for axis in range(3):
            step = np.zeros(3)
            step[axis] = eps
            fd = (oracle.mesh_to_points(m, pos + step, cell, nodes, method)
                  - oracle.mesh_to_points(m, pos - step, cell, nodes, method)) / (2 * eps)
            np.testing.assert_allclose(grad[safe, :, axis], fd[safe], rtol=1e-5, atol=1e-6)


def test_oracle_exact_agreement_on_mesh_points():
    """nodes = 1, 2 (P3M): atoms sitting on mesh points put their whole weight there (:103-147)"""
    rng = np.random.default_rng(8794329)
    for nodes in (1, 2):
        for n_mesh in (7, 10, 12):
            cell = rng.standard_normal((3, 3)) * 0.28209478
            # distinct mesh points
            flat = rng.choice(n_mesh ** 3, size=10, replace=False)
            ind = np.stack(np.unravel_index(flat, (n_mesh,) * 3), axis=1)
            pos = (ind / n_mesh) @ cell
            if nodes == 2:
                pos = pos + (0.5 / n_mesh) * cell.sum(0)     # even stencils are centred on cell midpoints
            w = 3 * rng.standard_normal((10, 3))
            mesh = oracle.points_to_mesh(w, pos, cell, (n_mesh,) * 3, nodes, "P3M")
            if nodes == 1:
                np.testing.assert_allclose(mesh[:, ind[:, 0], ind[:, 1], ind[:, 2]].T, w, rtol=1e-9, atol=1e-12)
            else:
                # x = 0: the two nodes of every axis share the weight equally -> 1/8 on the base corner
                np.testing.assert_allclose(mesh.sum(axis=(1, 2, 3)), w.sum(0), rtol=1e-12)
