"""
CUDA-graph capture of one energy + forces step.

Small systems are launch-latency bound: a 32k-atom P3M step is ~20 kernel launches of 3-8 us
each, while Python needs ~0.5 ms to issue them.  :class:`GraphedStep` captures

    V = calculator(charges, cell, positions, neighbor_indices, neighbor_distances)
    E = (V * charges).sum();   dE/dpositions, dE/dneighbor_distances = autograd.grad(E, ...)

once into a ``torch.cuda.CUDAGraph`` over static buffers and replays it; new inputs (host or
device tensors of the captured shapes) are copied into the static buffers first.  The cell
geometry, mesh size and neighbor-list length are frozen at capture time.

With ``host_io=True`` the host <-> device traffic is part of the graph: the step's inputs are read
from pinned host staging tensors (``.host["positions"]`` ...), the energy and forces land in pinned
host tensors (``.host["energy"]``, ``.host["grad_positions"]``).  The neighbor-list copy runs on the
graph branch of the real-space kernels, so the mesh pipeline (which only needs positions and
charges) overlaps it; one ``replay()`` + ``synchronize()`` is a complete host-to-host step.
"""

from __future__ import annotations

import os

import torch

from . import mesh as _mesh
from .mesh import set_nan_check


def _capture_stream(dev, dtype, calculator=None) -> "torch.cuda.Stream":
    """
    The stream a step is captured on carries the mesh chain, the critical path of the step; the pair kernels of
    the real-space branch (side stream) fill the GPU for 40-100 us each while the short mesh kernels wait for
    CTA slots.  TPME_GRAPH_PRIORITY=1 captures on a high-priority stream, which gives the kernel nodes of the mesh
    chain a higher launch priority than the branch kernels.  Opt-in, because the measured effect depends on the
    workload (profiles/r02_summary.md section 12): c5 0.268 -> 0.250 ms, c2 unchanged, c3 0.290 -> 0.293 ms,
    c4 0.836 -> 0.850 ms.  Never applied to a slab-decomposed calculator (its exchange / barrier kernels keep the
    stream layout they were validated with).
    """
    slab = getattr(calculator, "_slab_cfg", None) is not None or hasattr(calculator, "transport")
    high = os.environ.get("TPME_GRAPH_PRIORITY", "0") == "1" and not slab
    return torch.cuda.Stream(device=dev, priority=-1 if high else 0)


class GraphedStep:
    def __init__(self, calculator, charges, cell, positions, neighbor_indices, neighbor_distances,
                 warmup: int = 3, host_io: bool = False, fused_energy_gradients: bool = False):
        dev = positions.device
        # EXPERIMENTAL: one filter pass per step through calculator.energy_and_gradients()
        self.fused_energy_gradients = fused_energy_gradients
        if dev.type != "cuda":
            raise ValueError("GraphedStep needs CUDA tensors")
        self.calculator = calculator
        self.stream = _capture_stream(dev, positions.dtype, calculator)
        self.aux = torch.cuda.Stream(device=dev)     # side branch of the energy reduction
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        # the eager NaN guard is a host sync: off for the warm-up steps of the capture stream (it skips
        # itself while capturing), restored afterwards
        nan_check_before = _mesh._nan_check
        set_nan_check(False)
        with torch.cuda.stream(self.stream):
            # autograd ties a leaf to the stream it was created on: create them here
            self.charges = charges.detach().clone()
            self.cell = cell.detach().clone()
            self.positions = positions.detach().clone().requires_grad_(True)
            self.neighbor_indices = neighbor_indices.detach().clone()
            self.neighbor_distances = neighbor_distances.detach().clone().requires_grad_(True)
            self.host = None
            if host_io:
                self.host = {
                    "positions": positions.detach().cpu().pin_memory(),
                    "charges": charges.detach().cpu().pin_memory(),
                    "neighbor_indices": neighbor_indices.detach().cpu().pin_memory(),
                    "neighbor_distances": neighbor_distances.detach().cpu().pin_memory(),
                    "energy": torch.empty((), dtype=positions.dtype).pin_memory(),
                    "grad_positions": torch.empty(positions.shape, dtype=positions.dtype).pin_memory(),
                }
            # Inside this private graph the two branches of the calculator node need not join where the node
            # ends (calculators._FusedMeshPotential, defer_join): the mesh pipeline -- forward AND backward,
            # the latter only needs dE/dV = q -- and the D2H copy of the forces run while the pair list is
            # still crossing PCIe; what depends on the list (pair sum, energy, dE/dd) follows on the side
            # stream and everything joins once, at the end of the graph.
            defer = host_io and not fused_energy_gradients and hasattr(calculator, "_fused_config")
            self._deferred = defer
            had = getattr(calculator, "_defer_join", False)
            if defer:
                calculator._defer_join = True
            try:
                for _ in range(max(1, warmup)):
                    self._step(defer)
                    self._join()
                self.stream.synchronize()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    if host_io:
                        self._copy_in()
                    self.energy, self.grad_positions, self.grad_distances = self._step(defer)
                    if host_io:     # the forces are complete on this stream: their copy need not wait for the join
                        self.host["grad_positions"].copy_(self.grad_positions, non_blocking=True)
                    self._join()
                    if host_io:
                        self.host["energy"].copy_(self.energy, non_blocking=True)
            finally:
                if defer:
                    calculator._defer_join = had
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        set_nan_check(nan_check_before)
        # exchange buffers of a slab-decomposed calculator are baked into the graph: keep them alive
        self._keepalive = getattr(calculator, "_slab_cfg", None)

    def release(self) -> None:
        """drop the captured graph (call before destroying a process group whose collectives it holds)"""
        self.graph.reset()
        self._keepalive = None

    @torch.no_grad()
    def _copy_in(self):
        """captured host -> device copies; the pair list goes over the real-space branch"""
        from .calculators import _side_stream

        main = torch.cuda.current_stream(self.positions.device)
        side = _side_stream(self.positions.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.neighbor_distances.copy_(self.host["neighbor_distances"], non_blocking=True)
            self.neighbor_indices.copy_(self.host["neighbor_indices"], non_blocking=True)
        self.positions.copy_(self.host["positions"], non_blocking=True)
        self.charges.copy_(self.host["charges"], non_blocking=True)
        # the calculator forks the real-space kernels onto `side` (ordered after the copies above)
        # and joins before it needs them; every other consumer of the pair list sits behind that join

    def _join(self):
        """the capture stream waits for the energy branch and for the real-space branch of the calculator"""
        from .calculators import _side_stream

        if self.fused_energy_gradients:      # energy_and_gradients joins its own branches; `aux` is unused
            return
        main = torch.cuda.current_stream(self.positions.device)
        main.wait_stream(self.aux)
        if self._deferred:
            main.wait_stream(_side_stream(self.positions.device))

    def _step(self, defer: bool = False):
        from .calculators import _side_stream

        if self.fused_energy_gradients:
            energy, g_pos, g_d, _ = self.calculator.energy_and_gradients(
                self.charges, self.cell, self.positions, self.neighbor_indices, self.neighbor_distances)
            return energy, g_pos, g_d
        V = self.calculator(self.charges, self.cell, self.positions, self.neighbor_indices,
                            self.neighbor_distances)
        # E = sum_i q_i V_i: its reduction is a side branch of the graph, and the backward is seeded
        # directly with dE/dV = q (a vector-Jacobian product) instead of going through the tape of
        # the multiply + sum -- same numbers, three small kernels fewer on the critical path
        main = torch.cuda.current_stream(self.positions.device)
        # with deferred joins V is complete on the calculator's side stream, not here
        self.aux.wait_stream(_side_stream(self.positions.device) if defer else main)
        with torch.cuda.stream(self.aux):
            energy = (V.detach() * self.charges).sum()
        g_pos, g_d = torch.autograd.grad(V, (self.positions, self.neighbor_distances),
                                         grad_outputs=self.charges)
        return energy, g_pos, g_d    # the caller joins (self._join)

    def replay(self) -> None:
        self.graph.replay()

    @torch.no_grad()
    def __call__(self, positions=None, charges=None, neighbor_indices=None, neighbor_distances=None):
        """
        Copy the given inputs (any may be omitted to keep the previous values) into the static
        buffers, replay, and return ``(energy, dE/dpositions, dE/dneighbor_distances)`` -- views
        of static output buffers that the next call overwrites.
        """
        if positions is not None:
            self.positions.copy_(positions, non_blocking=True)
        if charges is not None:
            self.charges.copy_(charges, non_blocking=True)
        if neighbor_indices is not None:
            self.neighbor_indices.copy_(neighbor_indices, non_blocking=True)
        if neighbor_distances is not None:
            self.neighbor_distances.copy_(neighbor_distances, non_blocking=True)
        self.graph.replay()
        return self.energy, self.grad_positions, self.grad_distances


class GraphedPositionsStep:
    """
    Energy and forces from positions alone, as ONE CUDA graph: the neighbor list is rebuilt on the device
    inside the graph (``neighbors.DeviceNeighborList``: fixed-capacity buffers, the pair count never
    leaves the device), the distances are tied to the positions by ``neighbors.distances_from`` and the
    forces collect the mesh part and the real-space part:

        idx, d, S = list.build(positions);  d = distances_from(positions, cell, idx, S)
        V = calculator(charges, cell, positions, idx, d);  E = sum(q V);  F = -dE/dpositions

    With ``host_io=True`` the graph starts with the H2D copies of positions and charges from the pinned
    tensors ``.host["positions"]`` / ``.host["charges"]`` and ends with the D2H copies of the energy,
    ``dE/dpositions`` and the pair count -- 16 + 12 bytes per atom cross PCIe instead of the pair list.
    ``capacity`` defaults to 1.2 x the pair count of the positions given here; after a step
    ``overflowed()`` tells whether the list still fitted (a step that overflowed is incomplete: make a
    new object with a larger capacity).
    """

    def __init__(self, calculator, charges, cell, positions, cutoff: float, capacity: int | None = None,
                 host_io: bool = False, warmup: int = 3, index_dtype: torch.dtype = torch.int32):
        from .neighbors import DeviceNeighborList, neighbor_list

        dev = positions.device
        if dev.type != "cuda":
            raise ValueError("GraphedPositionsStep needs CUDA tensors")
        full = bool(getattr(calculator, "full_neighbor_list", False))
        if capacity is None:
            exact = neighbor_list(positions, cell, cutoff, full_neighbor_list=full, index_dtype=index_dtype)[0].shape[0]
            capacity = int(1.2 * exact) + 1024
        self.calculator, self.cutoff, self.capacity = calculator, float(cutoff), int(capacity)
        self.stream = _capture_stream(dev, positions.dtype, calculator)
        self.aux = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        nan_check_before = _mesh._nan_check
        set_nan_check(False)
        with torch.cuda.stream(self.stream):
            self.charges = charges.detach().clone()
            self.cell = cell.detach().clone()
            self.positions = positions.detach().clone().requires_grad_(True)
            self.list = DeviceNeighborList(positions.shape[0], self.cell, cutoff, self.capacity, dtype=positions.dtype,
                                           device=dev, full_neighbor_list=full, index_dtype=index_dtype)
            self.host = None
            if host_io:
                self.host = {
                    "positions": positions.detach().cpu().pin_memory(),
                    "charges": charges.detach().cpu().pin_memory(),
                    "energy": torch.empty((), dtype=positions.dtype).pin_memory(),
                    "grad_positions": torch.empty(positions.shape, dtype=positions.dtype).pin_memory(),
                    "n_pairs": torch.zeros((), dtype=torch.int64).pin_memory(),
                }
            for _ in range(max(1, warmup)):
                self._step()
            self.stream.synchronize()
            with torch.cuda.graph(self.graph, stream=self.stream):
                if host_io:
                    with torch.no_grad():
                        self.positions.copy_(self.host["positions"], non_blocking=True)
                        self.charges.copy_(self.host["charges"], non_blocking=True)
                self.energy, self.grad_positions = self._step()
                if host_io:
                    self.host["grad_positions"].copy_(self.grad_positions, non_blocking=True)
                    self.host["energy"].copy_(self.energy, non_blocking=True)
                    self.host["n_pairs"].copy_(self.list.n_pairs, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        set_nan_check(nan_check_before)

    def _step(self):
        from .calculators import _side_stream

        # the list is built on the real-space branch of the step (the stream the calculator runs its pair
        # kernels on): the mesh pipeline, which only needs positions and charges, does not wait for it
        main = torch.cuda.current_stream(self.positions.device)
        side = _side_stream(self.positions.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            idx, d, shifts = self.list.build(self.positions)
        if hasattr(self.calculator, "forward_from_pairs"):
            V = self.calculator.forward_from_pairs(self.charges, self.cell, self.positions, idx, shifts,
                                                   known_distances=d)
        else:
            from .neighbors import distances_from

            main.wait_stream(side)
            dist = distances_from(self.positions, self.cell, idx, shifts, known_distances=d)
            V = self.calculator(self.charges, self.cell, self.positions, idx, dist)
        self.aux.wait_stream(main)
        with torch.cuda.stream(self.aux):
            energy = (V.detach() * self.charges).sum()
        (g_pos,) = torch.autograd.grad(V, (self.positions,), grad_outputs=self.charges)
        main.wait_stream(self.aux)
        return energy, g_pos

    def replay(self) -> None:
        self.graph.replay()

    def overflowed(self) -> bool:
        """did the last step's pair list exceed the capacity? (host_io: reads the copied count, no extra sync)"""
        n = int(self.host["n_pairs"]) if self.host is not None else int(self.list.n_pairs)
        return n > self.capacity

    @torch.no_grad()
    def __call__(self, positions=None, charges=None):
        """copy the given inputs into the static buffers, replay, return ``(energy, dE/dpositions)``"""
        if positions is not None:
            self.positions.copy_(positions, non_blocking=True)
        if charges is not None:
            self.charges.copy_(charges, non_blocking=True)
        self.graph.replay()
        return self.energy, self.grad_positions
