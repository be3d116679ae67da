"""
Host-side logic that needs no GPU: argument validation texts (the reference's tests match
them by regex, tests/calculators/test_calculator.py:51-243), constructor checks, the torch
implementations of the potential interface against the oracle, mesh-size rule.
"""
import numpy as np
import pytest
import torch

import torchpme_b200 as tp
from oracle import pme_oracle as oracle
from torchpme_b200._checks import validate_parameters
from torchpme_b200.mesh import CellGeometry


def _good():
    return dict(charges=torch.ones(4, 1), cell=torch.eye(3), positions=torch.zeros(4, 3),
                neighbor_indices=torch.zeros(5, 2, dtype=torch.int64), neighbor_distances=torch.ones(5))


def test_valid_inputs_pass():
    validate_parameters(**_good())


@pytest.mark.parametrize("patch, exc, match", [
    (dict(positions=torch.zeros(4, 5)), ValueError, r"`positions` must be a tensor with shape \[n_atoms, 3\], got tensor with shape \[4, 5\]"),
    (dict(cell=torch.eye(2)), ValueError, r"`cell` must be a tensor with shape \[3, 3\], got tensor with shape \[2, 2\]"),
    (dict(cell=torch.eye(3, dtype=torch.float64)), TypeError, r"type of `cell` \(torch.float64\) must be same as that of the `positions` class \(torch.float32\)"),
    (dict(cell=torch.eye(3, device="meta")), ValueError, r"device of `cell` \(meta\) must be same as that of the `positions` class \(cpu\)"),
    (dict(charges=torch.ones(4)), ValueError, r"`charges` must be a 2-dimensional tensor, got tensor with 1 dimension\(s\) and shape \[4\]"),
    (dict(charges=torch.ones(6, 2)), ValueError, r"`charges` must be a tensor with shape \[n_atoms, n_channels\], with `n_atoms` being the same as the variable `positions`. Got tensor with shape \[6, 2\] where positions contains 4 atoms"),
    (dict(charges=torch.ones(4, 1, dtype=torch.float64)), TypeError, r"type of `charges` \(torch.float64\) must be same as that of the `positions` class \(torch.float32\)"),
    (dict(charges=torch.ones(4, 1, device="meta")), ValueError, r"device of `charges` \(meta\) must be same as that of the `positions` class \(cpu\)"),
    (dict(neighbor_indices=torch.zeros(5, 3, dtype=torch.int64)), ValueError, r"neighbor_indices is expected to have shape \[num_neighbors, 2\], but got \[5, 3\] for one structure"),
    (dict(neighbor_indices=torch.zeros(5, 2, dtype=torch.int64, device="meta")), ValueError, r"device of `neighbor_indices` \(meta\) must be same as that of the `positions` class \(cpu\)"),
    (dict(neighbor_distances=torch.ones(7)), ValueError, r"`neighbor_indices` and `neighbor_distances` need to have shapes \[num_neighbors, 2\] and \[num_neighbors\], but got \[5, 2\] and \[7\]"),
    (dict(neighbor_distances=torch.ones(5, device="meta")), ValueError, r"device of `neighbor_distances` \(meta\) must be same as that of the `positions` class \(cpu\)"),
    (dict(neighbor_distances=torch.ones(5, dtype=torch.float64)), TypeError, r"type of `neighbor_distances` \(torch.float64\) must be same as that of the `positions` class \(torch.float32\)"),
    (dict(periodic=torch.ones(2, dtype=torch.bool)), ValueError, r"`periodic` must be a tensor of shape \(3,\), got tensor with shape \[2\]"),
    (dict(pair_mask=torch.ones(4, dtype=torch.bool)), ValueError, r"`pair_mask` must have the same shape as the number of neighbors, got tensor with shape \[4\] while the number of neighbors is 5"),
    (dict(pair_mask=torch.ones(5)), TypeError, r"type of `pair_mask` \(torch.float32\) must be torch.bool"),
    (dict(node_mask=torch.ones(3, dtype=torch.bool)), ValueError, r"`node_mask` must have shape \[n_atoms\], got tensor with shape \[3\] where n_atoms is 4"),
    (dict(node_mask=torch.ones(4)), TypeError, r"type of `node_mask` \(torch.float32\) must be torch.bool"),
    (dict(kvectors=torch.ones(3, 2)), ValueError, r"`kvectors` must be a tensor of shape \[n_kvecs, 3\], got tensor with shape \[3, 2\]"),
    (dict(kvectors=torch.ones(3, 3, dtype=torch.float64)), TypeError, r"type of `kvectors` \(torch.float64\) must be same as that of the `positions` class \(torch.float32\)"),
])
def test_validation_messages(patch, exc, match):
    kw = _good()
    kw.update(patch)
    with pytest.raises(exc, match=match):
        validate_parameters(**kw)


def test_constructor_errors():
    with pytest.raises(TypeError, match="Potential must be an instance of Potential"):
        tp.Calculator(potential="coulomb")
    with pytest.raises(ValueError, match="Must specify smearing to use a potential with PMECalculator"):
        tp.PMECalculator(tp.CoulombPotential(), mesh_spacing=0.1)
    with pytest.raises(ValueError, match="`smearing` is -1.0 but must be positive"):
        tp.P3MCalculator(tp.CoulombPotential(smearing=-1.0), mesh_spacing=0.1)
    with pytest.raises(ValueError, match="`interpolation_nodes` is 8 but only values from 3 to 7 for method 'Lagrange' are allowed"):
        tp.PMECalculator(tp.CoulombPotential(smearing=1.0), mesh_spacing=0.1, interpolation_nodes=8)
    with pytest.raises(ValueError, match="`interpolation_nodes` is 6 but only values from 1 to 5 for method 'P3M' are allowed"):
        tp.P3MCalculator(tp.CoulombPotential(smearing=1.0), mesh_spacing=0.1, interpolation_nodes=6)
    with pytest.raises(ValueError, match="method 'cubic' is not supported. Choose from 'Lagrange' or 'P3M'"):
        tp.lib.MeshInterpolator(torch.eye(3), torch.tensor([4, 4, 4]), 4, "cubic")
    with pytest.raises(ValueError, match="Invalid option 'x' for the `fft_norm` parameter."):
        tp.lib.KSpaceFilter(torch.eye(3), torch.tensor([4, 4, 4]), tp.CoulombPotential(smearing=1.0), fft_norm="x")
    with pytest.raises(ValueError, match="Unsupported exponent: 7"):
        tp.InversePowerLawPotential(exponent=7, smearing=1.0)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.0), mesh_spacing=0.5)
    with pytest.raises(NotImplementedError, match="Batching not implemented for mesh-based calculators"):
        calc._compute_kspace(torch.ones(2, 1), torch.eye(3), torch.zeros(2, 3), node_mask=torch.ones(2, dtype=torch.bool))


def test_block_shape_errors():
    mi = tp.lib.MeshInterpolator(torch.eye(3), torch.tensor([4, 4, 4]), 4, "P3M")
    with pytest.raises(ValueError, match=r"cell of shape \[2, 2\] should be of shape \(3, 3\)"):
        mi.update(cell=torch.eye(2))
    with pytest.raises(ValueError, match=r"shape \[2\] of `ns_mesh` has to be \(3,\)"):
        mi.update(ns_mesh=torch.tensor([4, 4]))
    with pytest.raises(ValueError, match=r"shape \[5, 2\] of `positions` has to be \(N, 3\)"):
        mi.compute_weights(torch.zeros(5, 2))
    with pytest.raises(ValueError, match="`positions` device meta is not the same as instance device cpu"):
        mi.compute_weights(torch.zeros(5, 3, device="meta"))
    mi.compute_weights(torch.zeros(5, 3))
    with pytest.raises(ValueError, match="`particle_weights` of dimension 1 has to be of dimension 2"):
        mi.points_to_mesh(torch.zeros(5))
    with pytest.raises(ValueError, match="`mesh_vals` of dimension 3 has to be of dimension 4"):
        mi.mesh_to_points(torch.zeros(4, 4, 4))
    kf = tp.lib.KSpaceFilter(torch.eye(3), torch.tensor([4, 4, 4]), tp.CoulombPotential(smearing=1.0))
    with pytest.raises(ValueError, match="`mesh_values` needs to be a 4 dimensional tensor, got 3"):
        kf.forward(torch.zeros(4, 4, 4))
    with pytest.raises(ValueError, match="The real-space mesh is inconsistent with the k-space grid."):
        kf.forward(torch.zeros(1, 4, 4, 5))


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_potential_interface_matches_oracle(p):
    """torch implementations of the potential interface (used by the table route) vs the oracle"""
    s = 0.9
    pot = tp.InversePowerLawPotential(exponent=p, smearing=s, prefactor=1.3)
    ref = oracle.PotentialSpec("ipl", s, exponent=p, prefactor=1.3)
    d = torch.linspace(0.4, 5.0, 64, dtype=torch.float64)
    k_sq = torch.cat([torch.zeros(1, dtype=torch.float64), torch.linspace(0.01, 40.0, 80, dtype=torch.float64)])
    np.testing.assert_allclose(pot.from_dist(d).numpy(), ref.from_dist(d.numpy()), rtol=1e-13)
    np.testing.assert_allclose(pot.lr_from_dist(d).numpy(), ref.lr_from_dist(d.numpy()), rtol=1e-12)
    np.testing.assert_allclose(pot.sr_from_dist(d).numpy(), ref.sr_from_dist(d.numpy()), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(pot.lr_from_k_sq(k_sq).numpy(), ref.lr_from_k_sq(k_sq.numpy()), rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(float(pot.self_contribution()), ref.self_contribution(), rtol=1e-13)
    np.testing.assert_allclose(float(pot.background_correction()), ref.background_correction(), rtol=1e-13)
    if p == 1:
        c = tp.CoulombPotential(smearing=s, prefactor=1.3)
        np.testing.assert_allclose(c.lr_from_k_sq(k_sq).numpy(), pot.lr_from_k_sq(k_sq).numpy(), rtol=1e-13)
        np.testing.assert_allclose(c.sr_from_dist(d).numpy(), pot.sr_from_dist(d).numpy(), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(float(c.self_contribution()), ref.self_contribution(), rtol=1e-13)
        np.testing.assert_allclose(float(c.background_correction()), ref.background_correction(), rtol=1e-13)


def test_exclusion_cutoff_function():
    pot = tp.CoulombPotential(smearing=1.0, exclusion_radius=2.0, exclusion_degree=3)
    ref = oracle.PotentialSpec("coulomb", 1.0, exclusion_radius=2.0, exclusion_degree=3)
    d = torch.linspace(0.1, 4.0, 50, dtype=torch.float64)
    np.testing.assert_allclose(pot.f_cutoff(d).numpy(), ref.f_cutoff(d.numpy()), rtol=1e-13)
    np.testing.assert_allclose(pot.sr_from_dist(d).numpy(), ref.sr_from_dist(d.numpy()), rtol=1e-12)
    with pytest.raises(ValueError, match="Cannot compute cutoff function when `exclusion_radius` is not set"):
        tp.CoulombPotential(smearing=1.0).f_cutoff(d)


def test_exp1_and_gradient():
    import scipy.special

    x = torch.cat([torch.rand(2000, dtype=torch.float64), 1 + 50 * torch.rand(2000, dtype=torch.float64)]).requires_grad_(True)
    y = tp.lib.exp1(x)
    np.testing.assert_allclose(y.detach().numpy(), scipy.special.exp1(x.detach().numpy()), rtol=1e-13)
    y.sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), -np.exp(-x.detach().numpy()) / x.detach().numpy(), rtol=1e-13)


def test_mesh_size_rule_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(50):
        cell = np.eye(3) * rng.uniform(1, 30) + rng.uniform(-1, 1, (3, 3))
        h = rng.uniform(0.05, 2.0)
        assert CellGeometry(torch.tensor(cell)).ns_mesh(h) == tuple(oracle.get_ns_mesh(cell, h))
        np.testing.assert_array_equal(tp.lib.get_ns_mesh(torch.tensor(cell), h).numpy(), oracle.get_ns_mesh(cell, h))


def test_kvectors_match_oracle():
    rng = np.random.default_rng(6)
    cell = np.eye(3) * 4.0 + rng.uniform(-0.4, 0.4, (3, 3))
    for ns in ((4, 5, 6), (7, 3, 8), (1, 1, 1)):
        kv = tp.lib.generate_kvectors_for_mesh(torch.tensor(cell), torch.tensor(ns))
        np.testing.assert_allclose(kv.numpy(), oracle.kvectors_for_mesh(cell, ns), rtol=1e-13, atol=1e-14)


def test_fused_config_is_host_only_and_cached():
    """the by-value launch parameters of the fast path are pure host arithmetic"""
    calc = tp.P3MCalculator(tp.CoulombPotential(smearing=1.2, prefactor=2.0), mesh_spacing=1.5,
                            interpolation_nodes=4)
    cell = torch.eye(3, dtype=torch.float64) * 12.0
    cfg = calc._fused_config(cell)
    assert cfg.ns == (32, 32, 32) and cfg.nodes == 4 and cfg.method == 0 and cfg.full_list is False
    assert abs(cfg.half_ivolume - 0.5 / 12.0**3) < 1e-18
    assert abs(cfg.self_half - 0.5 * 2.0 * np.sqrt(2 / np.pi) / 1.2) < 1e-14
    assert abs(cfg.background_ivolume - 2.0 * np.pi * 1.2**2 / 12.0**3) < 1e-16
    assert cfg.green_args["p3m_nodes"] == 4 and cfg.green_args["kind"] == 1
    np.testing.assert_allclose(np.asarray(cfg.r2u).reshape(3, 3), np.eye(3) * 32 / 12.0, rtol=1e-15)
    assert calc._fused_config(cell) is cfg                      # same cell tensor: cached
    pme = tp.PMECalculator(tp.InversePowerLawPotential(exponent=6, smearing=1.0), mesh_spacing=1.5)
    cfg2 = pme._fused_config(cell)
    assert cfg2.method == 1 and cfg2.green_args["p3m_nodes"] == 0 and cfg2.green_args["exponent"] == 6
    assert cfg2.background_ivolume == 0.0
    with pytest.raises(tp.NativeLibraryError, match="CUDA-only"):
        calc.energy_and_gradients(torch.ones(2, 1, dtype=torch.float64), cell, torch.zeros(2, 3, dtype=torch.float64),
                                  torch.zeros(1, 2, dtype=torch.int64), torch.ones(1, dtype=torch.float64))


def test_tuning_timings_on_cpu():
    """TuningTimings (tuning/tuner.py:283-373): same constructor / forward contract, CPU tensors timed on the host"""
    import torchpme_b200 as tp
    from torchpme_b200.tuning import TuningTimings

    pos = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], dtype=torch.float64)
    q = torch.tensor([[1.0], [-1.0]], dtype=torch.float64)
    cell = torch.eye(3, dtype=torch.float64)
    idx = torch.tensor([[0, 1]])
    d = torch.tensor([0.8660254], dtype=torch.float64)
    timings = TuningTimings(q, cell, pos, idx, d, n_repeat=2, n_warmup=1, run_backward=True)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=0.2), mesh_spacing=0.1).to(torch.float64)
    seconds = timings(calc)
    assert isinstance(seconds, float) and 0 < seconds < 10
    with pytest.raises(ValueError):       # the reference validates the structure in the constructor
        TuningTimings(q, torch.eye(2, dtype=torch.float64), pos, idx, d)


def test_device_cell_registers_geometry_from_host_values(monkeypatch):
    """device_cell(): the geometry of the returned tensor comes from the host values -- geometry_of() must not
    read the tensor back (on CUDA that read is the one stream sync of a step with a new cell)"""
    from torchpme_b200 import mesh

    rng = np.random.default_rng(11)
    box = np.eye(3) * 9.0 + rng.uniform(-0.3, 0.3, (3, 3))
    for dtype in (torch.float64, torch.float32):
        cell = tp.device_cell(box, "cpu", dtype)
        assert cell.dtype == dtype and cell.shape == (3, 3)
        built = []
        original = mesh.CellGeometry.__init__
        monkeypatch.setattr(mesh.CellGeometry, "__init__", lambda self, c: (built.append(1), original(self, c))[1])
        geom = mesh.geometry_of(cell)
        assert not built                                          # cache hit: nothing rebuilt
        monkeypatch.setattr(mesh.CellGeometry, "__init__", original)
        fresh = mesh.CellGeometry(cell.clone())                   # what a device read-back would have given
        np.testing.assert_array_equal(geom.cell, fresh.cell)
        np.testing.assert_array_equal(np.asarray(geom.recip), np.asarray(fresh.recip))
        assert geom.volume == fresh.volume and geom.ns_mesh(0.7) == fresh.ns_mesh(0.7)
        cell.mul_(1.01)                                           # in-place change: the cached geometry is stale
        assert mesh.geometry_of(cell).volume != geom.volume
    assert tp.device_cell([[1, 0, 0], [0, 2, 0], [0, 0, 3]], "cpu").dtype == torch.get_default_dtype()
    with pytest.raises(ValueError, match="should be of shape"):
        tp.device_cell(np.eye(2), "cpu")
