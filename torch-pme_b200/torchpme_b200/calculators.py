"""
Calculators: real-space pair sum + (PME | P3M) mesh pipeline behind the reference's
``forward(charges, cell, positions, neighbor_indices, neighbor_distances, ...)`` call.

Host-side mirror of ``src/torchpme/calculators/{calculator,pme,p3m}.py``.  Every stage is a
CUDA kernel of ``libtorchpme_b200.so``; tensors must live on a CUDA device -- there is no
CPU path (a CPU tensor raises ``NativeLibraryError``).
"""

from __future__ import annotations

import os
import weakref
from typing import Optional

import torch
from torch.autograd.function import once_differentiable

from . import _cpu, _native
from ._checks import validate_parameters
from .mesh import KSpaceFilter, MeshInterpolator, P3MKSpaceFilter, check_filter_result, geometry_of
from .potentials import Potential


# --------------------------------------------------------------------------------------
# TorchScript boundary.  ``torch.jit.script(calculator)`` (and ``jit.save`` / ``jit.load``, which the
# reference's tests/calculators/test_workflow.py:130-162 exercise) compiles ``forward``; everything behind
# it here is ctypes + autograd.Function code that TorchScript cannot compile.  ``forward`` is therefore a
# scriptable shell: eagerly it calls ``_forward_impl`` directly, under TorchScript it calls the operator
# ``torchpme_b200::calculator_forward`` registered below with the PyTorch dispatcher, which looks the
# live calculator up by its handle and runs the same ``_forward_impl`` -- as a CompositeImplicitAutograd
# kernel, so the autograd nodes of the implementation are recorded on the tape as usual.  A scripted
# calculator is thus a shell around a dispatcher operator that needs this package (and the calculator it
# was scripted from) in the same process; it does not make the module deployable without Python.
# --------------------------------------------------------------------------------------
_CALCULATORS: "weakref.WeakValueDictionary[int, Calculator]" = weakref.WeakValueDictionary()
_OPS = torch.library.Library("torchpme_b200", "DEF")
_OPS.define("calculator_forward(Tensor charges, Tensor cell, Tensor positions, Tensor neighbor_indices, "
            "Tensor neighbor_distances, Tensor? periodic, Tensor? node_mask, Tensor? pair_mask, "
            "Tensor? kvectors, int handle) -> Tensor")


def _calculator_forward_op(charges, cell, positions, neighbor_indices, neighbor_distances, periodic,
                           node_mask, pair_mask, kvectors, handle):
    calc = _CALCULATORS.get(int(handle))
    if calc is None:
        raise RuntimeError(
            "torchpme_b200::calculator_forward: the calculator this scripted module was created from is "
            "not alive in this process (a scripted torchpme_b200 calculator is a shell around the Python "
            "implementation, see torchpme_b200/calculators.py)")
    return calc._forward_impl(charges, cell, positions, neighbor_indices, neighbor_distances, periodic,
                              node_mask, pair_mask, kvectors)


_OPS.impl("calculator_forward", _calculator_forward_op, "CompositeImplicitAutograd")


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


_side_streams: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """one auxiliary stream per device for work that is independent of the mesh pipeline"""
    key = torch.device(device).index
    stream = _side_streams.get(key)
    if stream is None:
        stream = _side_streams[key] = torch.cuda.Stream(device=device)
    return stream


_zero_streams: dict = {}
# TPME_PREZERO=0: the spreads zero-fill their mesh themselves, on the critical path (A/B switch)
_PREZERO = os.environ.get("TPME_PREZERO", "1") != "0"


def _zeroed_meshes(like, n_channels: int, ns, count: int, branch: bool = True):
    """
    `count` zero-filled (C, nx, ny, nz) meshes for the spreads of one fused step.  The fills run on their own
    stream -- a parallel branch of a captured graph -- next to the tile sort instead of between the sort and
    the spread (c3: two 16.8 MB fills, ~10 us each, leave the mesh chain); the caller's stream joins that
    branch in :func:`_join_zeroed` right before the first spread.  The mesh of the backward spread is
    filled here as well, so the backward pass starts with its spread.
    """
    dev = like.device
    shape = (n_channels, int(ns[0]), int(ns[1]), int(ns[2]))
    meshes = [torch.empty(shape, dtype=like.dtype, device=dev) for _ in range(count)]
    if not (_PREZERO and branch):
        return meshes, None
    key = torch.device(dev).index
    stream = _zero_streams.get(key)
    if stream is None:
        stream = _zero_streams[key] = torch.cuda.Stream(device=dev)
    stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(stream):
        for m in meshes:
            m.zero_()
    return meshes, stream


def _join_zeroed(device, stream) -> bool:
    """the caller's stream waits for the fills; returns whether the meshes are zero (spread with accumulate)"""
    if stream is None:
        return False
    torch.cuda.current_stream(device).wait_stream(stream)
    return True


class _FusedStepConfig:
    """by-value launch parameters of one fused PME / P3M evaluation"""
    __slots__ = ("r2u", "ns", "nodes", "method", "green_args", "pair_pot", "full_list",
                 "half_ivolume", "self_half", "background_ivolume", "defer_join")


class _FusedMeshPotential(torch.autograd.Function):
    """
    Whole calculator forward as one autograd node (fast path: in-kernel potential, no cell
    gradient, 3-D periodic):

        V = pair_sum(q, d) + 1/(2 Vol) gather(A spread(q)) - q self/2 - bg sum(q)/Vol

    Six kernel launches forward (pair, spread, FFT . G . iFFT, gather with the O(N) corrections
    fused as epilogue, which also emits dV/dr) and five backward (pair backward, spread of the
    incoming gradient, FFT . G . iFFT, gather + derivative gather with the saved dV/dr folded in).
    """

    @staticmethod
    def forward(ctx, charges, positions, distances, neighbor_indices, mask_u8, cfg, shifts=None, cell_host=None):
        q = charges.detach().contiguous()
        pos = positions.detach().contiguous()
        idx = neighbor_indices.contiguous()
        need_pos = ctx.needs_input_grad[1]
        out = torch.empty_like(q)
        # `shifts` given: the distances are |r_j + S . cell - r_i| of THESE positions (forward_from_pairs):
        # they are computed on the real-space branch unless the caller already has them, and the backward
        # carries dL/dd on to the positions on that branch too
        ctx.shifts, ctx.cell_host = shifts, cell_host
        d = None if distances is None else distances.detach().contiguous()
        # the pair sum is independent of the mesh pipeline until the gather epilogue: run it (and
        # the zero fill of its accumulator) on a side stream -- a parallel branch when the step is
        # captured in a CUDA graph.  `out` is next touched on the main stream after the join.
        main = torch.cuda.current_stream(q.device)
        side = _side_stream(q.device)
        side.wait_stream(main)
        need_dist = d is None
        if need_dist:       # allocated on the main stream like `out`: both branches join before anything is freed
            d = torch.zeros(idx.shape[0], dtype=q.dtype, device=q.device)
        # defer_join (set by GraphedStep for its private graph only): the two branches never join inside this
        # node -- the pair sum goes to its own buffer and is added to the mesh result ON THE SIDE STREAM, so the
        # mesh pipeline (forward and backward) does not wait for a pair list that is still crossing PCIe.  The
        # returned potentials (and dL/dd in backward) are then complete on the side stream only; the caller joins.
        defer = bool(cfg.defer_join) and not ctx.needs_input_grad[0]
        ctx.defer = defer
        pair_out = torch.empty_like(q) if defer else out
        with torch.cuda.stream(side):
            pair_out.zero_()
            if need_dist:
                _native.pair_distances(pos, cell_host, idx, shifts, _native.pair_count_of(idx), d)
            _native.pair_forward(q, idx, d, None, mask_u8, cfg.full_list, cfg.pair_pot, out=pair_out)
        if defer:
            out.zero_()
        # the meshes of this spread and of the backward one are zero-filled on a branch of their own
        # (not with deferred joins: there the mesh chain hides behind the pair-list copy, nothing to gain)
        n_spreads = 2 if (ctx.needs_input_grad[0] or need_pos) and not defer else 1
        meshes, zs = _zeroed_meshes(q, q.shape[1], cfg.ns, n_spreads, branch=not defer)
        # atoms binned by mesh tile once per step: the spread, the gather and both backward launches
        # stage their pencil of the mesh in shared memory (csrc/tiles.cu); None = direct kernels
        tiles = _native.tile_sort(pos, cfg.r2u, cfg.ns, cfg.nodes, cfg.method)
        zeroed = _join_zeroed(q.device, zs)
        rho = _native.spread(pos, q, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, tiles=tiles, out=meshes[0],
                             accumulate=zeroed)
        ctx.zero_mesh = meshes[1] if zeroed and len(meshes) > 1 else None
        green = _native.make_green(scale=1.0, **cfg.green_args)
        phi, _, dc = _native.kfilter_apply(rho, green, want_dc=True)
        check_filter_result(phi)      # the reference's NaN guard (a host sync; see mesh.set_nan_check)
        if not defer:
            main.wait_stream(side)
        epi = _native.make_epilogue(q, dc, cfg.half_ivolume, cfg.self_half, cfg.background_ivolume)
        _, dvalues = _native.gather(phi, pos, cfg.r2u, cfg.nodes, cfg.method, want_grad=need_pos,
                                    values_out=out, epilogue=epi, tiles=tiles)
        if defer:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out.add_(pair_out)
        ctx.cfg, ctx.tiles = cfg, tiles
        ctx.save_for_backward(q, pos, d, idx, mask_u8, dvalues)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, pos, d, idx, mask_u8, dvalues = ctx.saved_tensors
        cfg, tiles = ctx.cfg, ctx.tiles
        need_q, need_pos, need_d = ctx.needs_input_grad[:3]
        shifts = ctx.shifts
        if shifts is not None:           # distances derived from the positions: dL/dd ends in dL/dpositions
            need_d = need_pos
        g = grad_out.contiguous()
        g_q = torch.zeros_like(q) if need_q else None
        g_d = None
        main = torch.cuda.current_stream(q.device)
        side = _side_stream(q.device)
        forked = False
        if need_q or need_d:
            # dL/dd is produced by the pair kernel alone -> side stream; when dL/dq is wanted too the
            # gather epilogue accumulates into the same buffer, so the join happens before it
            g_d = None
            if need_d:     # padding entries of a counted (fixed-capacity) list are never written: zeros there
                counted = _native.pair_count_of(idx) is not None
                g_d = (torch.zeros if counted else torch.empty)(idx.shape[0], dtype=q.dtype, device=q.device)
            g_pos4 = None
            if shifts is not None and need_d:
                g_pos4 = torch.zeros((pos.shape[0], 4), dtype=q.dtype, device=q.device)
            side.wait_stream(main)
            forked = True
            with torch.cuda.stream(side):
                _native.pair_backward(q, idx, d, None, mask_u8, g, cfg.full_list, cfg.pair_pot,
                                      want_charges=need_q, want_pairs=need_d, grad_charges_out=g_q,
                                      grad_pairs_out=g_d)
                if g_pos4 is not None:
                    _native.pair_distances_backward(pos, ctx.cell_host, idx, shifts, g_d,
                                                    _native.pair_count_of(idx), g_pos4, None)
        g_pos = None
        if need_q or need_pos:
            zero_mesh, ctx.zero_mesh = ctx.zero_mesh, None       # filled during forward; good for one backward
            rho_g = _native.spread(pos, g, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, tiles=tiles, out=zero_mesh,
                                   accumulate=zero_mesh is not None)
            green = _native.make_green(scale=1.0, **cfg.green_args)
            psi, _, dc_g = _native.kfilter_apply(rho_g, green, want_dc=True)
            if forked and need_q:
                main.wait_stream(side)
                forked = False
            epi = _native.make_epilogue(g, dc_g, cfg.half_ivolume, cfg.self_half, cfg.background_ivolume,
                                        coef2=g if need_pos else None, dvalues2=dvalues,
                                        vjp_scale=cfg.half_ivolume)
            if need_pos:
                g_pos, _, _ = _native.gather_vjp(psi, pos, q, cfg.r2u, cfg.nodes, cfg.method,
                                                 values_out=g_q, epilogue=epi, tiles=tiles)
            else:
                _native.gather(psi, pos, cfg.r2u, cfg.nodes, cfg.method, values_out=g_q, epilogue=epi,
                               tiles=tiles)
        if forked and not (ctx.defer and not need_q):
            main.wait_stream(side)
        if shifts is not None:
            if need_d:
                if ctx.defer and forked and not need_q:
                    main.wait_stream(side)      # the real-space forces are added to the mesh forces here
                g_pos = g_pos4[:, :3] if g_pos is None else g_pos + g_pos4[:, :3]
            g_d = None
        return g_q, g_pos, g_d, None, None, None, None, None


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
        self._handle: int = id(self)
        _CALCULATORS[self._handle] = self

    @torch.jit.unused
    def _compute_rspace(self, charges, neighbor_indices, neighbor_distances, pair_mask=None):
        pot = self.potential
        mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
        descriptor = pot._native_descriptor()
        # the in-kernel potentials take smearing / prefactor by value: when any of them needs a
        # gradient the per-pair values come from the potential's own differentiable torch code
        # (reference: potentials/potential.py:106-138 feeds autograd the same way)
        if descriptor is not None and (not charges.is_cuda or any(
                t.requires_grad for t in list(pot.parameters()) + list(pot.buffers()))):
            descriptor = None
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
        if not charges.is_cuda:               # device dispatch: CPU tensors use the torch formulation
            return _cpu.pair_sum(charges, neighbor_indices, bare.to(charges.dtype), self.full_neighbor_list)
        native_pot = _native.make_pair_potential(0)
        return _PairSum.apply(charges, bare.to(charges.dtype), neighbor_indices, mask_u8,
                              native_pot, self.full_neighbor_list)

    @torch.jit.unused
    def _compute_kspace(self, charges, cell, positions, periodic=None, node_mask=None, kvectors=None):
        raise NotImplementedError(f"`compute_kspace` not implemented for {self.__class__.__name__}")

    def forward(self, charges: torch.Tensor, cell: torch.Tensor, positions: torch.Tensor,
                neighbor_indices: torch.Tensor, neighbor_distances: torch.Tensor,
                periodic: Optional[torch.Tensor] = None, node_mask: Optional[torch.Tensor] = None,
                pair_mask: Optional[torch.Tensor] = None, kvectors: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not torch.jit.is_scripting():
            return self._forward_impl(charges, cell, positions, neighbor_indices, neighbor_distances,
                                      periodic, node_mask, pair_mask, kvectors)
        return torch.ops.torchpme_b200.calculator_forward(
            charges, cell, positions, neighbor_indices, neighbor_distances, periodic, node_mask, pair_mask,
            kvectors, self._handle)

    @torch.jit.unused
    @torch.compiler.disable   # ctypes launches: torch.compile runs this frame eagerly (one graph break)
    def _forward_impl(self, charges, cell, positions, neighbor_indices, neighbor_distances,
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
        self._fused_cfg = self._fused_key = self._fused_geom = None
        self._fused_pot_key = self._fused_pot_terms = None
        unit = torch.eye(3, device=potential.smearing.device, dtype=potential.smearing.dtype)
        ones = torch.ones(3, dtype=torch.int64, device=unit.device)
        self.kspace_filter = self._make_filter(unit, ones)
        self.mesh_interpolator = MeshInterpolator(unit, ones, interpolation_nodes, self._method)

    def _make_filter(self, cell, ns):
        return KSpaceFilter(cell, ns, kernel=self.potential, fft_norm="backward", ifft_norm="forward")

    def _fast_path_ok(self, cell, periodic, node_mask, kvectors) -> bool:
        pot = self.potential
        if periodic is not None or node_mask is not None or kvectors is not None:
            return False
        if pot._native_descriptor() is None or cell.requires_grad:
            return False
        return not any(t.requires_grad for t in list(pot.parameters()) + list(pot.buffers()))

    @torch.jit.unused
    @torch.compiler.disable   # ctypes launches: torch.compile runs this frame eagerly (one graph break)
    def _forward_impl(self, charges, cell, positions, neighbor_indices, neighbor_distances,
                      periodic=None, node_mask=None, pair_mask=None, kvectors=None):
        # the fused node is CUDA only; CPU tensors go through the modular blocks, which dispatch to the
        # torch formulation of _cpu.py (CUDA tensors never do)
        if not positions.is_cuda or not self._fast_path_ok(cell, periodic, node_mask, kvectors):
            return super()._forward_impl(charges, cell, positions, neighbor_indices, neighbor_distances,
                                         periodic, node_mask, pair_mask, kvectors)
        validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances,
                            periodic, pair_mask, node_mask, kvectors)
        cfg = self._fused_config(cell)
        mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
        return _FusedMeshPotential.apply(charges, positions, neighbor_distances, neighbor_indices,
                                         mask_u8, cfg, None, None)

    def forward_from_pairs(self, charges, cell, positions, neighbor_indices, neighbor_shifts,
                           known_distances=None, pair_mask=None):
        """
        ``forward`` for a pair list given as indices + integer image shifts (what a neighbor search returns:
        ``neighbors.neighbor_list``, vesin's ``"PS"`` quantities): equal to

            d = neighbors.distances_from(positions, cell, neighbor_indices, neighbor_shifts)
            V = self(charges, cell, positions, neighbor_indices, d)

        including the gradient that reaches ``positions`` through the distances -- but on the fast path the
        distance kernels run on the real-space branch of the step, next to the pair kernels and concurrent
        with the mesh pipeline, instead of as separate autograd nodes around it.  ``known_distances``: the
        distances of these very positions when the caller has them already (skips their evaluation).
        """
        from .neighbors import distances_from

        if positions.is_cuda and not cell.requires_grad and self._fast_path_ok(cell, None, None, None):
            validate_parameters(charges, cell, positions, neighbor_indices,
                                known_distances if known_distances is not None else
                                torch.empty(neighbor_indices.shape[0], dtype=positions.dtype, device=positions.device),
                                None, pair_mask, None, None)
            cfg = self._fused_config(cell)
            mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
            return _FusedMeshPotential.apply(charges, positions, known_distances, neighbor_indices, mask_u8, cfg,
                                             neighbor_shifts.contiguous(), geometry_of(cell).cell)
        d = distances_from(positions, cell, neighbor_indices, neighbor_shifts, known_distances=known_distances)
        return self(charges, cell, positions, neighbor_indices, d, pair_mask=pair_mask)

    def _fused_config(self, cell) -> _FusedStepConfig:
        """by-value launch parameters of the fast path, cached per (cell geometry, potential scalars)"""
        pot = self.potential
        geom = geometry_of(cell)
        ns = geom.ns_mesh(self.mesh_spacing)
        kind, exponent = pot._native_descriptor()
        smearing, prefactor = pot._scalars()
        defer_join = bool(getattr(self, "_defer_join", False))
        key = (id(geom), ns, kind, exponent, smearing, prefactor, pot.exclusion_radius,
               pot.exclusion_degree, self.full_neighbor_list, defer_join)
        cfg = self._fused_cfg if self._fused_key == key else None
        if cfg is None:
            cfg = _FusedStepConfig()
            cfg.defer_join = defer_join
            cfg.r2u, cfg.ns = geom.r2u(ns), ns
            cfg.nodes, cfg.method = self.interpolation_nodes, _native.METHOD_ID[self._method]
            cfg.green_args = dict(kind=kind, exponent=exponent, smearing=smearing, prefactor=prefactor,
                                  recip=geom.recip, spacing=geom.spacing(ns),
                                  p3m_nodes=self.interpolation_nodes if self._method == "P3M" else 0)
            cfg.pair_pot = _native.make_pair_potential(kind, smearing, prefactor, exponent,
                                                       pot.exclusion_radius, pot.exclusion_degree)
            cfg.full_list = self.full_neighbor_list
            ivolume = 1.0 / geom.volume
            cfg.half_ivolume = 0.5 * ivolume
            # self / background terms depend on the potential's scalars only: evaluated (on the device, by the
            # potential's own torch code) and read back once per potential state, not once per cell
            pot_key = (kind, exponent, smearing, prefactor)
            if self._fused_pot_key != pot_key:
                self._fused_pot_terms = (float(pot.self_contribution()), float(pot.background_correction()))
                self._fused_pot_key = pot_key
            cfg.self_half = 0.5 * self._fused_pot_terms[0]
            cfg.background_ivolume = self._fused_pot_terms[1] * ivolume
            self._fused_cfg, self._fused_key, self._fused_geom = cfg, key, geom
        return cfg

    @torch.no_grad()
    def energy_and_gradients(self, charges, cell, positions, neighbor_indices, neighbor_distances,
                             pair_mask=None):
        """
        EXPERIMENTAL (not part of the reference API, not yet measured on a GPU):
        ``E = sum_ic q_ic V_ic`` with ``dE/dpositions`` and ``dE/dneighbor_distances`` from ONE
        spread, ONE filter pass and ONE gather.  The autograd backward of ``E`` spreads the incoming
        gradient ``dE/dV = q`` and filters it again; because the filter is self-adjoint that mesh
        equals the forward one, so the two derivative gathers coincide:

            dE/dr_i = (1 / Vol) sum_c q_ic  d/dr_i [gather(phi_c)]_i ,   phi = A spread(q)

        ("fused energy+forces path").  Same restrictions as the fast path of
        :meth:`forward`; returns ``(E, dE/dpositions, dE/dneighbor_distances, V)``.
        """
        if not self._fast_path_ok(cell, None, None, None):
            raise NotImplementedError(
                "energy_and_gradients needs an in-kernel potential and no cell / parameter gradients")
        validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances,
                            None, pair_mask, None, None)
        if not positions.is_cuda:
            raise _native.NativeLibraryError(
                f"`positions` lives on {positions.device}; torchpme_b200 is a CUDA-only implementation "
                "(no CPU fallback). Move the inputs to a CUDA device.")
        cfg = self._fused_config(cell)
        q, pos = charges.detach().contiguous(), positions.detach().contiguous()
        d, idx = neighbor_distances.detach().contiguous(), neighbor_indices.contiguous()
        mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
        out = torch.empty_like(q)
        g_d = torch.empty(idx.shape[0], dtype=q.dtype, device=q.device)
        main = torch.cuda.current_stream(q.device)
        side = _side_stream(q.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            out.zero_()
            _native.pair_forward(q, idx, d, None, mask_u8, cfg.full_list, cfg.pair_pot, out=out)
            _native.pair_backward(q, idx, d, None, mask_u8, q, cfg.full_list, cfg.pair_pot,
                                  want_charges=False, want_pairs=True, grad_pairs_out=g_d)
        meshes, zs = _zeroed_meshes(q, q.shape[1], cfg.ns, 1)
        tiles = _native.tile_sort(pos, cfg.r2u, cfg.ns, cfg.nodes, cfg.method)
        zeroed = _join_zeroed(q.device, zs)
        rho = _native.spread(pos, q, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, tiles=tiles, out=meshes[0],
                             accumulate=zeroed)
        green = _native.make_green(scale=1.0, **cfg.green_args)
        phi, _, dc = _native.kfilter_apply(rho, green, want_dc=True)
        main.wait_stream(side)
        epi = _native.make_epilogue(q, dc, cfg.half_ivolume, cfg.self_half, cfg.background_ivolume)
        _, dvalues = _native.gather(phi, pos, cfg.r2u, cfg.nodes, cfg.method, want_grad=True,
                                    values_out=out, epilogue=epi, tiles=tiles)
        g_pos = torch.einsum("ic,icd->id", q, dvalues) * (2.0 * cfg.half_ivolume)
        energy = (out * q).sum()
        return energy, g_pos, g_d, out

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
