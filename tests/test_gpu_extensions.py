"""
GPU tests of the entry points that round 1 left unvalidated (they ran green on a B200 at the start of
round 2 and are part of the default GPU suite since): the one-filter-pass energy + gradients step,
the device-built neighbor list feeding the calculators, and the spline / combined potentials through
the generic table and per-pair-value routes.
"""
import numpy as np
import pytest
import torch

from helpers import rel_err, rocksalt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method, dtype", [("P3M", torch.float64), ("PME", torch.float64), ("P3M", torch.float32)])
def test_energy_and_gradients_matches_autograd(method, dtype):
    import torchpme_b200 as tp

    pos, q, cell, idx, d = rocksalt(8, dtype=dtype, device="cuda")
    q = torch.cat([q, 0.4 * q + 0.2], dim=1)
    cls = tp.P3MCalculator if method == "P3M" else tp.PMECalculator
    calc = cls(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14)
    p = pos.clone().requires_grad_(True)
    dd = d.clone().requires_grad_(True)
    V = calc(q, cell, p, idx, dd)
    energy = (V * q).sum()
    gp, gd = torch.autograd.grad(energy, (p, dd))
    e2, gp2, gd2, V2 = calc.energy_and_gradients(q, cell, pos, idx, d)
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    assert rel_err(V2, V.detach()) < tol
    assert abs(float(e2) - float(energy)) < tol * abs(float(energy))
    assert rel_err(gd2, gd) < tol
    assert rel_err(gp2, gp) < tol * 10
    graphed = tp.GraphedStep(calc, q, cell, pos, idx, d, fused_energy_gradients=True)
    graphed.replay()
    torch.cuda.synchronize()
    assert rel_err(graphed.grad_positions, gp) < tol * 10 and rel_err(graphed.grad_distances, gd) < tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("full", [False, True])
def test_gpu_neighbor_list_matches_oracle_and_feeds_the_calculator(full, dtype):
    import numpy as np

    import torchpme_b200 as tp
    from oracle import pme_oracle as oracle
    from torchpme_b200.neighbors import distances_from, neighbor_list

    pos, q, cell, idx_ref, d_ref = rocksalt(6, dtype=torch.float64, device="cuda", cutoff=5.0)
    idx, d, s = neighbor_list(pos.to(dtype), cell.to(dtype), 5.0, full_neighbor_list=full)
    o_idx, o_d, o_s = oracle.neighbor_list(pos.cpu().numpy(), cell.cpu().numpy(), 5.0, full=full)
    assert idx.shape[0] == o_idx.shape[0]
    assert abs(float(d.double().sum()) - float(o_d.sum())) < 1e-4 * float(o_d.sum())
    assert rel_err(d, distances_from(pos.to(dtype), cell.to(dtype), idx, s)) < (1e-12 if dtype == torch.float64 else 1e-5)
    # the calculator gives the same potentials with this list as with the synthetic generator's
    calc = tp.P3MCalculator(tp.CoulombPotential(smearing=1.0).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14,
                            full_neighbor_list=full)
    V = calc(q.to(dtype), cell.to(dtype), pos.to(dtype), idx, d)
    if not full:
        V_ref = calc(q.to(dtype), cell.to(dtype), pos.to(dtype), idx_ref, d_ref.to(dtype))
        assert rel_err(V, V_ref) < (1e-10 if dtype == torch.float64 else 1e-4)
    assert np.isfinite(V.cpu().numpy()).all()


def test_spline_and_combined_potentials_through_the_generic_routes():
    """a spline of the smeared Coulomb potential reproduces CoulombPotential's long-range part"""
    import math

    import torchpme_b200 as tp

    dt = torch.float64
    pos, q, cell, idx, d = rocksalt(6, dtype=dt, device="cuda", cutoff=5.0)
    smearing = 1.0
    r = torch.logspace(-2, 2, 800, dtype=dt)
    y = torch.erf(r / smearing / 2 ** 0.5) / r
    spline = tp.SplinePotential(r, y, reciprocal=True, y_at_zero=math.sqrt(2 / math.pi) / smearing,
                                yhat_at_zero=0.0, smearing=smearing).to("cuda")
    coulomb = tp.CoulombPotential(smearing=smearing).to("cuda")
    spacing = float(cell[0, 0]) / 14
    v_spline = tp.PMECalculator(spline, mesh_spacing=spacing)._compute_kspace(q, cell, pos)
    v_coulomb = tp.PMECalculator(coulomb, mesh_spacing=spacing)._compute_kspace(q, cell, pos)
    assert rel_err(v_spline, v_coulomb) < 1e-3
    combined = tp.CombinedPotential([coulomb, tp.InversePowerLawPotential(exponent=4, smearing=smearing).to("cuda")],
                                    initial_weights=torch.tensor([1.0, 0.0], dtype=dt), smearing=smearing).to("cuda")
    p = pos.clone().requires_grad_(True)
    V = tp.PMECalculator(combined, mesh_spacing=spacing)(q, cell, p, idx, d)
    V_ref = tp.PMECalculator(coulomb, mesh_spacing=spacing)(q, cell, pos, idx, d)
    assert rel_err(V.detach(), V_ref) < 1e-9
    (V * q).sum().backward()
    assert combined.weights.grad is not None and p.grad is not None


def test_tuning_timings_measure_device_time():
    """event-timed TuningTimings: close to the CUDA-graph replayed step time, not to the Python launch overhead"""
    import torchpme_b200 as tp
    from torchpme_b200.tuning import TuningTimings

    pos, q, cell, idx, d = rocksalt(16, dtype=torch.float32, device="cuda")
    calc_fine = tp.P3MCalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=float(cell[0, 0]) / 62)
    calc_coarse = tp.P3MCalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14)
    timings = TuningTimings(q, cell, pos, idx, d, n_repeat=4, n_warmup=4, flush_l2=True)
    t_fine, t_coarse = timings(calc_fine), timings(calc_coarse)
    assert 0 < t_coarse < 0.05 and 0 < t_fine < 0.05
    assert t_fine > t_coarse * 0.5      # a 128^3 mesh is not cheaper than a 32^3 one


def test_torch_compile_wraps_the_calculator():
    """tests/calculators/test_workflow.py:146-153: a torch.compile'd calculator runs and returns a tensor"""
    import torchpme_b200 as tp

    pos, q, cell, idx, d = rocksalt(6, dtype=torch.float32, device="cuda", cutoff=5.0)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.0).to("cuda"), mesh_spacing=float(cell[0, 0]) / 14)
    ref = calc(q, cell, pos, idx, d)
    # Dynamo + AOT autograd are what touch this package (the kernels sit behind ctypes calls inside
    # autograd.Function nodes: graph breaks there, eager execution of those frames); Inductor's code
    # generation only concerns the torch ops around them and depends on the host toolchain
    for backend in ("aot_eager", "inductor"):
        torch._dynamo.reset()
        compiled = torch.compile(calc, backend=backend)
        p = pos.clone().requires_grad_(True)
        try:
            out = compiled(q, cell, p, idx, d)
        except Exception as exc:      # a missing host compiler for Inductor is not this package's business
            if backend == "inductor" and "InductorError" in type(exc).__name__:
                continue
            raise
        assert type(out) is torch.Tensor and rel_err(out, ref) < 1e-5
        (out * q).sum().backward()
        assert p.grad is not None and torch.isfinite(p.grad).all()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("mode, order", [(1, 1), (1, 2), (2, 4), (3, 6), (0, 2)])
@pytest.mark.parametrize("triclinic", [False, True])
def test_p3m_influence_modes_in_kernel(mode, order, triclinic, dtype):
    """P3MKSpaceFilter modes 1-3 (lib/kspace_filter.py:307-347) evaluated per k-point inside the filter kernel
    against the table route (the reference's torch expression, CPU-checked against the reference in
    test_spline_potentials.py).  The filter TABLES are compared (tpme_green_table writes what the fused x pass
    multiplies by).  For modes >= 1 the planes through a Nyquist frequency are left out: there the
    finite-difference operator D vanishes analytically and the formula (k.D)^m / (U^2 |D|^4m) divides
    rounding noise by rounding noise -- in the reference too (sin(pi) = 1.2e-16, not 0)."""
    import torchpme_b200 as tp
    from torchpme_b200 import _native
    from torchpme_b200.lib import P3MKSpaceFilter
    from torchpme_b200.mesh import geometry_of

    gen = torch.Generator().manual_seed(mode * 10 + order)
    cell = torch.eye(3, dtype=torch.float64) * 9.0
    if triclinic:
        cell = cell + 0.8 * torch.rand(3, 3, generator=gen, dtype=torch.float64)
    ns = (16, 32, 16)
    pot = tp.CoulombPotential(smearing=0.9).to("cuda")
    c_grad = cell.to("cuda", dtype).requires_grad_(True)          # a cell gradient selects the table route
    f_tab = P3MKSpaceFilter(c_grad, torch.tensor(ns).cuda(), 4, pot, mode=mode, differential_order=order,
                            fft_norm="backward", ifft_norm="forward")
    assert f_tab._wants_table()
    want = f_tab._kfilter.detach()
    f_gpu = P3MKSpaceFilter(cell.to("cuda", dtype), torch.tensor(ns).cuda(), 4, pot, mode=mode,
                            differential_order=order, fft_norm="backward", ifft_norm="forward")
    assert not f_gpu._wants_table()
    geom = geometry_of(f_gpu.cell)
    green = _native.make_green(kind=_native.GREEN_COULOMB, scale=1.0, recip=geom.recip, spacing=geom.spacing(ns),
                               smearing=0.9, prefactor=1.0, p3m_nodes=4, p3m_mode=mode, differential_order=order)
    got = _native.green_table(dtype, ns, green, "cuda")
    keep = torch.ones(want.shape, dtype=torch.bool, device="cuda")
    if mode > 0:
        keep[ns[0] // 2, :, :] = False
        keep[:, ns[1] // 2, :] = False
        keep[:, :, ns[2] // 2] = False
    tol = 1e-10 if dtype == torch.float64 else 5e-4
    scale = want[keep].abs().max()
    assert float((got[keep] - want[keep]).abs().max() / scale) < tol
    # and through the filter itself.  Mode 0 only: for modes >= 1 the k-points (n/2, 0, 0) ... carry
    # rounding-noise / rounding-noise values of order 1e17 in the table AND in the kernel (see above), which
    # multiply the FFT round-off of any mesh -- comparing two such results says nothing
    if mode == 0:
        mesh = torch.randn((2,) + ns, generator=gen, dtype=torch.float64).to("cuda", dtype)
        a, b = f_gpu(mesh), f_tab(mesh).detach()
        assert rel_err(a, b) < (1e-10 if dtype == torch.float64 else 2e-4)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("reciprocal", [True, False], ids=["reciprocal-axis", "direct"])
@pytest.mark.parametrize("p3m", [False, True], ids=["pme", "p3m"])
def test_spline_kernel_evaluated_in_the_filter_kernel(p3m, reciprocal, dtype):
    """SplinePotential.lr_from_k_sq (potentials/spline.py:151-157: cubic spline in k^2, direct or on a reciprocal
    axis) evaluated per k-point inside the filter kernel (green kinds 3 / 4) against the table the potential's
    own torch code builds; triclinic cell, a mesh that reaches below the first and beyond the last knot"""
    import math

    import torchpme_b200 as tp
    from torchpme_b200 import _native
    from torchpme_b200.lib import KSpaceFilter, P3MKSpaceFilter
    from torchpme_b200.mesh import geometry_of

    smearing = 1.0
    if reciprocal:
        r = torch.logspace(-2, 2, 400, dtype=torch.float64)
        y = torch.erf(r / smearing / 2 ** 0.5) / r
        pot = tp.SplinePotential(r, y, reciprocal=True, y_at_zero=math.sqrt(2 / math.pi) / smearing,
                                 yhat_at_zero=0.0, smearing=smearing, prefactor=1.7)
    else:
        r = torch.linspace(0.0, 12.0, 300, dtype=torch.float64)
        y = torch.exp(-0.5 * (r / 1.3) ** 2)
        k = torch.linspace(0.0, 9.0, 200, dtype=torch.float64)
        pot = tp.SplinePotential(r, y, k_grid=k, smearing=smearing, prefactor=0.6)
    pot = pot.to("cuda")
    gen = torch.Generator().manual_seed(7)
    cell = (torch.eye(3, dtype=torch.float64) * 8.0 + 0.7 * torch.rand(3, 3, generator=gen, dtype=torch.float64)).to("cuda", dtype)
    ns = (16, 32, 16)
    make = (lambda c: P3MKSpaceFilter(c, torch.tensor(ns).cuda(), 4, pot, fft_norm="backward", ifft_norm="forward")) \
        if p3m else (lambda c: KSpaceFilter(c, torch.tensor(ns).cuda(), pot, fft_norm="backward", ifft_norm="forward"))
    f_gpu, f_tab = make(cell), make(cell.clone().requires_grad_(True))
    assert not f_gpu._wants_table() and f_tab._wants_table()
    want = f_tab._kfilter.detach()
    geom = geometry_of(f_gpu.cell)
    green = _native.make_green(scale=1.0, recip=geom.recip, spacing=geom.spacing(ns), p3m_nodes=4 if p3m else 0,
                               **pot._native_filter())
    assert green.kind == (4 if reciprocal else 3)
    got = _native.green_table(dtype, ns, green, "cuda")
    tol = 1e-11 if dtype == torch.float64 else 3e-5
    assert float((got - want.to(dtype)).abs().max() / want.abs().max()) < tol
    mesh = torch.randn((2,) + ns, generator=gen, dtype=torch.float64).to("cuda", dtype)
    assert rel_err(f_gpu(mesh), f_tab(mesh).detach()) < (1e-10 if dtype == torch.float64 else 2e-4)


def test_debug_bounds_check_catches_a_bad_pair_list(monkeypatch):
    """TPME_DEBUG_BOUNDS: out-of-range neighbor indices raise instead of writing out of bounds"""
    import torchpme_b200 as tp
    from torchpme_b200 import _native

    pos, q, cell, idx, d = rocksalt(4, dtype=torch.float32, device="cuda", cutoff=5.0)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.0).to("cuda"), mesh_spacing=float(cell[0, 0]) / 6)
    monkeypatch.setattr(_native, "DEBUG_BOUNDS", True)
    calc(q, cell, pos, idx, d)
    bad = idx.clone()
    bad[3, 1] = pos.shape[0]
    with pytest.raises(IndexError, match="neighbor_indices must lie in"):
        calc(q, cell, pos, bad, d)


def test_nan_guard_on_the_fused_path():
    """the reference raises on NaNs in the filtered mesh (lib/kspace_filter.py:189-195): the fused calculator
    node does too, unless the guard is switched off"""
    import torchpme_b200 as tp

    pos, q, cell, idx, d = rocksalt(4, dtype=torch.float32, device="cuda", cutoff=5.0)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.0).to("cuda"), mesh_spacing=float(cell[0, 0]) / 6)
    bad = q.clone()
    bad[3, 0] = float("nan")
    tp.set_nan_check(True)
    with pytest.raises(ValueError, match="NaNs detected in the k-space filter result"):
        calc(bad, cell, pos, idx, d)
    tp.set_nan_check(False)
    try:
        assert torch.isnan(calc(bad, cell, pos, idx, d)).any()
    finally:
        tp.set_nan_check(True)


def test_device_cell_step_has_no_host_sync():
    """
    A step with a NEW cell tensor made by `device_cell` (geometry registered from the host values) runs without
    any host synchronisation: forward + backward under torch's sync-debug "error" mode.
    """
    import torchpme_b200 as tp
    from helpers import rocksalt

    pos, q, cell, idx, d = rocksalt(8, dtype=torch.float32, device="cuda")
    mesh_spacing = float(cell[0, 0]) / (16 / 2 - 2)
    calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=mesh_spacing)
    box = cell.cpu().numpy().astype(np.float64)
    tp.set_nan_check(False)          # the reference's NaN guard on the filtered mesh is a host read by design
    try:
        p = pos.clone().requires_grad_(True)
        V0 = calc(q, cell, p, idx, d)
        (V0 * q).sum().backward()     # warm: plans, allocator
        torch.cuda.synchronize()
        torch.cuda.set_sync_debug_mode("error")
        try:
            cell2 = tp.device_cell(box * 1.0005, "cuda", torch.float32)
            p2 = pos.clone().requires_grad_(True)
            V = calc(q, cell2, p2, idx, d)
            (V * q).sum().backward()
        finally:
            torch.cuda.set_sync_debug_mode("default")
        torch.cuda.synchronize()
        assert torch.isfinite(V).all() and torch.isfinite(p2.grad).all()
        assert float((V - V0).abs().max()) < 1e-2 * float(V0.abs().max())      # a 0.05 % larger box
    finally:
        tp.set_nan_check(True)
