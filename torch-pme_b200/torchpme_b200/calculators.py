"""
Calculators: real-space pair sum + (PME | P3M) mesh pipeline behind the reference's
``forward(charges, cell, positions, neighbor_indices, neighbor_distances, ...)`` call.

Host-side mirror of ``src/torchpme/calculators/{calculator,pme,p3m}.py``.  Every stage is a
CUDA kernel of ``libtorchpme_b200.so``; tensors must live on a CUDA device -- there is no
CPU path (a CPU tensor raises ``NativeLibraryError``).
"""

from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _native
from ._checks import validate_parameters
from .mesh import KSpaceFilter, MeshInterpolator, P3MKSpaceFilter, geometry_of
from .potentials import Potential


class _PairSum(torch.autograd.Function):
    """
    out[i, c] = 1/2 sum_pairs q[j, c] v(d) (+ mirrored term for half lists).  ``pair_input`` is
    the distance for the in-kernel potentials and the per-pair value v for the generic route.
    """

    @staticmethod
    def forward(ctx, charges, pair_input, neighbor_indices, mask_u8, pot, full_list):
        q = charges.detach().contiguous()
        x = pair_input.detach().contiguous()
        idx = neighbor_indices.contiguous()
        dist, values = (None, x) if pot.kind == 0 else (x, None)
        ctx.pot, ctx.full_list = pot, full_list
        ctx.save_for_backward(q, x, idx, mask_u8)
        return _native.pair_forward(q, idx, dist, values, mask_u8, full_list, pot)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, x, idx, mask_u8 = ctx.saved_tensors
        need_q, need_x = ctx.needs_input_grad[:2]
        dist, values = (None, x) if ctx.pot.kind == 0 else (x, None)
        g_q, g_x = _native.pair_backward(q, idx, dist, values, mask_u8, grad_out.contiguous(),
                                         ctx.full_list, ctx.pot, want_charges=need_q, want_pairs=need_x)
        return g_q, g_x, None, None, None, None


class Calculator(torch.nn.Module):
    """
    Real-space part V_i = 1/2 sum_j q_j v(r_ij) over a neighbor list, plus the long-range
    part of subclasses (reference: ``calculators/calculator.py:8-189``).
    """

    def __init__(self, potential: Potential, full_neighbor_list: bool = False):
        super().__init__()
        if not isinstance(potential, Potential):
            raise TypeError(f"Potential must be an instance of Potential, got {type(potential)}")
        self.potential = potential
        self.full_neighbor_list = full_neighbor_list

    def _compute_rspace(self, charges, neighbor_indices, neighbor_distances, pair_mask=None):
        pot = self.potential
        mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
        descriptor = pot._native_descriptor()
        if descriptor is not None and pot.smearing is not None:
            kind, exponent = descriptor
            smearing, prefactor = pot._scalars()
            native_pot = _native.make_pair_potential(
                kind, smearing, prefactor, exponent, pot.exclusion_radius, pot.exclusion_degree)
            return _PairSum.apply(charges, neighbor_distances, neighbor_indices, mask_u8,
                                  native_pot, self.full_neighbor_list)
        # generic potentials: v(d) from the potential's own torch code, pair sum in the kernel
        if pot.smearing is None:
            bare = pot.from_dist(neighbor_distances, pair_mask)
            if pot.exclusion_radius is not None:
                bare = bare * (1 - pot.f_cutoff(neighbor_distances, pair_mask))
        else:
            bare = pot.sr_from_dist(neighbor_distances, pair_mask)
        native_pot = _native.make_pair_potential(0)
        return _PairSum.apply(charges, bare.to(charges.dtype), neighbor_indices, mask_u8,
                              native_pot, self.full_neighbor_list)

    def _compute_kspace(self, charges, cell, positions, periodic=None, node_mask=None, kvectors=None):
        raise NotImplementedError(f"`compute_kspace` not implemented for {self.__class__.__name__}")

    def forward(self, charges, cell, positions, neighbor_indices, neighbor_distances,
                periodic=None, node_mask=None, pair_mask=None, kvectors=None):
        validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances,
                            periodic, pair_mask, node_mask, kvectors)
        potential_sr = self._compute_rspace(charges, neighbor_indices, neighbor_distances, pair_mask)
        if self.potential.smearing is None:
            return potential_sr
        potential_lr = self._compute_kspace(charges, cell, positions, periodic=periodic,
                                            kvectors=kvectors, node_mask=node_mask)
        return potential_sr + potential_lr


class PMECalculator(Calculator):
    """
    Particle-mesh Ewald with Lagrange interpolation (3..7 nodes)
    (reference: ``calculators/pme.py:10-143``).
    """

    _method = "Lagrange"

    def __init__(self, potential: Potential, mesh_spacing: float, interpolation_nodes: int = 4,
                 full_neighbor_list: bool = False):
        super().__init__(potential=potential, full_neighbor_list=full_neighbor_list)
        if potential.smearing is None:
            raise ValueError("Must specify smearing to use a potential with PMECalculator")
        if potential.smearing <= 0:
            raise ValueError(f"`smearing` is {potential.smearing} but must be positive")
        self.mesh_spacing = mesh_spacing
        self.interpolation_nodes = interpolation_nodes
        unit = torch.eye(3, device=potential.smearing.device, dtype=potential.smearing.dtype)
        ones = torch.ones(3, dtype=torch.int64, device=unit.device)
        self.kspace_filter = self._make_filter(unit, ones)
        self.mesh_interpolator = MeshInterpolator(unit, ones, interpolation_nodes, self._method)

    def _make_filter(self, cell, ns):
        return KSpaceFilter(cell, ns, kernel=self.potential, fft_norm="backward", ifft_norm="forward")

    def _compute_kspace(self, charges, cell, positions, periodic=None, node_mask=None, kvectors=None):
        if node_mask is not None or kvectors is not None:
            raise NotImplementedError("Batching not implemented for mesh-based calculators")
        pot = self.potential
        geom = geometry_of(cell)
        ns = geom.ns_mesh(self.mesh_spacing)
        self.mesh_interpolator._update_host(cell, ns)
        self.kspace_filter._update_host(cell, ns)

        self.mesh_interpolator.compute_weights(positions)
        rho = self.mesh_interpolator.points_to_mesh(charges)
        phi = self.kspace_filter.forward(rho)
        if cell.requires_grad:
            ivolume = torch.abs(torch.linalg.det(cell)).pow(-1)
        else:
            ivolume = 1.0 / geom.volume
        out = self.mesh_interpolator.mesh_to_points(phi) * ivolume

        # self term, neutralising background (x2: everything is halved below), slab term
        out = out - charges * pot.self_contribution().to(charges.dtype)
        background = pot.background_correction().to(charges.dtype)
        out = out - (2 * background * ivolume) * charges.sum(dim=0)
        if periodic is not None:
            out = out + pot.pbc_correction(periodic, positions, cell, charges).to(charges.dtype)
        return out / 2


class P3MCalculator(PMECalculator):
    """
    Particle-particle particle-mesh: P3M charge assignment (1..5 nodes) and influence function
    (reference: ``calculators/p3m.py:9-84``).
    """

    _method = "P3M"

    def _make_filter(self, cell, ns):
        return P3MKSpaceFilter(cell, ns, interpolation_nodes=self.interpolation_nodes,
                               kernel=self.potential, mode=0, differential_order=2,
                               fft_norm="backward", ifft_norm="forward")
