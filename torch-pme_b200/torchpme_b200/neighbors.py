"""
Neighbor list on the GPU.

The reference takes ``neighbor_indices`` / ``neighbor_distances`` from the external ``vesin``
package (tests/helpers.py:240-275, examples/basic-usage.py:161-169); at 1 M atoms that CPU step and
the copy of its pair list dwarf a sub-millisecond GPU evaluation.  Here the list is built on the
device by ``libtorchpme_b200.so`` (``csrc/neighbors.cu``): atoms are wrapped into the cell, binned into
slabs between lattice planes and counting-sorted by bin; one thread per sorted atom walks the
neighbouring bins once (half of them for half lists), parks its partners in shared memory, and the CTA
reserves its output range with one atomic on the device-resident pair counter.

* :func:`neighbor_list` -- exact-size result, one host synchronisation (the pair count);
* :class:`DeviceNeighborList` -- fixed-capacity buffers, no host synchronisation: the number of pairs
  stays on the device and the pair kernels read it there, so list construction + calculator + backward
  form one CUDA graph (``graphs.GraphedPositionsStep``);
* :func:`distances_from` -- the differentiable ``|r_j + S . cell - r_i|`` (forward and backward are one
  kernel each on CUDA) that carries the forces from ``neighbor_distances`` back to the positions.

Validated on a B200 against the brute-force oracle (``tests/test_gpu_neighbors.py``) and on the CPU
through a host build of the same per-atom code (``tests/test_neighbors.py``).
"""

from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _native
from .mesh import geometry_of


_NeighborSearch = _native._NeighborSearch


def search_layout(cell: np.ndarray, cutoff: float, periodic=(True, True, True), bins_per_cutoff: int = 2,
                  max_bins: int = 1 << 21):
    """
    Bins per lattice direction and how many of them the search visits on each side.  Bins are
    slabs between lattice planes of thickness ``h_a / n_bins[a]`` (``h_a`` = perpendicular height of
    the cell); two points closer than ``cutoff`` differ by at most ``ceil(cutoff / thickness)`` bins.
    Non-periodic directions use one bin and no images.
    """
    inv = np.linalg.inv(np.asarray(cell, dtype=np.float64))
    heights = 1.0 / np.linalg.norm(inv, axis=0)
    n_bins = [max(1, int(math.floor(bins_per_cutoff * heights[a] / cutoff))) if periodic[a] else 1
              for a in range(3)]
    while n_bins[0] * n_bins[1] * n_bins[2] > max_bins:      # very large cells: coarser bins
        a = int(np.argmax(n_bins))
        n_bins[a] = max(1, n_bins[a] // 2)
    reach = [int(math.ceil(cutoff / (heights[a] / n_bins[a]) - 1e-12)) if periodic[a] else 0 for a in range(3)]
    return n_bins, reach


def _make_search(cell_np: np.ndarray, cutoff: float, full: bool, periodic) -> _NeighborSearch:
    n_bins, reach = search_layout(cell_np, cutoff, periodic)
    search = _NeighborSearch()
    for k, v in enumerate(np.asarray(cell_np, dtype=np.float64).reshape(-1)):
        search.cell[k] = float(v)
    for a in range(3):
        search.n_bins[a], search.reach[a], search.periodic[a] = n_bins[a], reach[a], int(periodic[a])
    search.full_list, search.cutoff = int(full), float(cutoff)
    return search


def _check_inputs(positions, cell, cutoff):
    if positions.dim() != 2 or positions.shape[1] != 3:
        raise ValueError(f"`positions` must be a tensor with shape [n_atoms, 3], got {list(positions.shape)}")
    if cell.shape != (3, 3):
        raise ValueError(f"`cell` must be a tensor with shape [3, 3], got {list(cell.shape)}")
    if not cutoff > 0:
        raise ValueError(f"`cutoff` must be positive, got {cutoff}")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Workspace:
    """device arrays of one build (sorted atoms, bin table, pair counter)"""

    def __init__(self, n: int, search: _NeighborSearch, dtype, device):
        lib = _native.load()
        bins = search.n_bins[0] * search.n_bins[1] * search.n_bins[2]
        self.n = n
        self.scratch = torch.empty(int(lib.tpme_nl_scratch_ints(n, ctypes.byref(search))), dtype=torch.int32, device=device)
        self.bin_start = torch.empty(bins + 1, dtype=torch.int32, device=device)
        self.sorted_rec = torch.empty((max(n, 1), 4), dtype=dtype, device=device)
        self.sorted_shift = torch.empty((max(n, 1), 4), dtype=torch.int32, device=device)
        self.n_pairs = torch.zeros((), dtype=torch.int64, device=device)


def _sort(positions, search, ws: _Workspace):
    lib = _native.load()
    with _native._on(positions, "positions"):
        _native._check(lib.tpme_nl_sort(_native._dtype_id(positions), _ptr(positions), ws.n, ctypes.byref(search),
                                        _ptr(ws.scratch), _ptr(ws.bin_start), _ptr(ws.sorted_rec),
                                        _ptr(ws.sorted_shift), _native._stream()), "tpme_nl_sort")
    _native._count(3)


def _pairs(positions, search, ws: _Workspace, indices=None, distances=None, shifts=None):
    """the search; without output buffers it only counts (ws.n_pairs)"""
    lib = _native.load()
    with _native._on(positions, "positions"):
        _native._check(lib.tpme_nl_pairs(_native._dtype_id(positions), _ptr(ws.sorted_rec), _ptr(ws.sorted_shift),
                                         _ptr(ws.bin_start), ws.n, ctypes.byref(search),
                                         0 if indices is None else indices.shape[0],
                                         int(indices is not None and indices.dtype == torch.int64), _ptr(indices),
                                         _ptr(distances), _ptr(shifts), _ptr(ws.n_pairs), _native._stream()),
                       "tpme_nl_pairs")
    _native._count()


def _host_build(lib, positions, search, index_dtype):
    """the same per-atom code compiled for the host (tests/native/nl_host.cpp): count call, then fill call"""
    pos = positions.detach().contiguous()
    n = pos.shape[0]
    args = (0 if pos.dtype == torch.float32 else 1, _ptr(pos), ctypes.c_int64(n), search.cell, search.n_bins,
            search.reach, search.periodic, ctypes.c_int(search.full_list), ctypes.c_double(search.cutoff))
    lib.nl_host_build.restype = ctypes.c_int64
    n_pairs = int(lib.nl_host_build(*args, ctypes.c_int64(0), 0, None, None, None))
    assert n_pairs >= 0
    indices = torch.empty((n_pairs, 2), dtype=index_dtype)
    distances = torch.empty(n_pairs, dtype=pos.dtype)
    shifts = torch.empty((n_pairs, 3), dtype=torch.int32)
    got = int(lib.nl_host_build(*args, ctypes.c_int64(n_pairs), int(index_dtype == torch.int64), _ptr(indices),
                                _ptr(distances), _ptr(shifts)))
    assert got == n_pairs
    return indices, distances, shifts


@torch.no_grad()
def neighbor_list(positions: torch.Tensor, cell: torch.Tensor, cutoff: float, full_neighbor_list: bool = False,
                  periodic=(True, True, True), index_dtype: torch.dtype = torch.int64, _host_library=None):
    """
    All pairs ``(i, j, S)`` with ``|r_j + S . cell - r_i| < cutoff`` (except ``(i, i, 0)``): every
    unordered pair once (``i < j`` for any ``S``, self images with ``S`` lexicographically positive)
    or, with ``full_neighbor_list``, in both directions.

    Returns ``neighbor_indices (P, 2)`` (``index_dtype``: int64 like the reference's lists, or int32),
    ``neighbor_distances (P,)`` and ``shifts (P, 3) int32`` on the device of ``positions``; the pairs
    found from one atom are contiguous, the order of those runs is not fixed.  One host synchronisation
    (the pair count).  For
    differentiable distances recompute them from the result with :func:`distances_from`.
    """
    _check_inputs(positions, cell, cutoff)
    periodic = tuple(bool(p) for p in periodic)
    search = _make_search(geometry_of(cell).cell, cutoff, full_neighbor_list, periodic)
    if _host_library is not None:
        return _host_build(_host_library, positions, search, index_dtype)
    if not positions.is_cuda:
        raise _native.NativeLibraryError(
            f"`positions` lives on {positions.device}; torchpme_b200 is a CUDA-only implementation "
            "(no CPU fallback). Move the inputs to a CUDA device.")
    pos = positions.detach().contiguous()
    ws = _Workspace(pos.shape[0], search, pos.dtype, pos.device)
    _sort(pos, search, ws)
    _pairs(pos, search, ws)                        # count only
    n_pairs = int(ws.n_pairs)                      # the one synchronisation
    indices = torch.empty((n_pairs, 2), dtype=index_dtype, device=pos.device)
    distances = torch.empty(n_pairs, dtype=pos.dtype, device=pos.device)
    shifts = torch.empty((n_pairs, 3), dtype=torch.int32, device=pos.device)
    if n_pairs:
        _pairs(pos, search, ws, indices, distances, shifts)
    return indices, distances, shifts


class DeviceNeighborList:
    """
    Neighbor list in fixed-capacity device buffers, rebuilt without a host synchronisation.

    ``build(positions)`` returns ``(neighbor_indices (capacity, 2), neighbor_distances (capacity,),
    shifts (capacity, 3))``; the number of valid pairs stays on the device (``n_pairs``, a 0-dim int64
    view) and is attached to the index buffer, where the real-space kernels of the calculators and
    :func:`distances_from` pick it up: entries beyond it are ignored (their gradients stay zero).
    All launches are CUDA-graph capturable.  Check :meth:`overflowed` (one host read) at a point
    where the host waits anyway; a list that did not fit is incomplete -- rebuild with a larger
    ``capacity``.
    """

    def __init__(self, n_atoms: int, cell: torch.Tensor, cutoff: float, capacity: int, dtype=None,
                 device=None, full_neighbor_list: bool = False, periodic=(True, True, True),
                 index_dtype: torch.dtype = torch.int32):
        if not cutoff > 0:
            raise ValueError(f"`cutoff` must be positive, got {cutoff}")
        dtype = dtype or cell.dtype
        device = torch.device(device or cell.device)
        if device.type != "cuda":
            raise _native.NativeLibraryError("DeviceNeighborList is a CUDA-only implementation (no CPU fallback).")
        self.search = _make_search(geometry_of(cell).cell, cutoff, full_neighbor_list,
                                   tuple(bool(p) for p in periodic))
        self.capacity = int(capacity)
        self.ws = _Workspace(n_atoms, self.search, dtype, device)
        self.indices = torch.zeros((self.capacity, 2), dtype=index_dtype, device=device)
        self.distances = torch.ones(self.capacity, dtype=dtype, device=device)
        self.shifts = torch.zeros((self.capacity, 3), dtype=torch.int32, device=device)
        self.n_pairs = self.ws.n_pairs
        _native.register_pair_count(self.indices, self.n_pairs)

    @torch.no_grad()
    def build(self, positions: torch.Tensor):
        if positions.shape != (self.ws.n, 3) or positions.dtype != self.distances.dtype:
            raise ValueError("`positions` must have the shape / dtype this list was created for")
        pos = positions.detach().contiguous()
        _sort(pos, self.search, self.ws)
        _pairs(pos, self.search, self.ws, self.indices, self.distances, self.shifts)
        return self.indices, self.distances, self.shifts

    def overflowed(self) -> bool:
        return int(self.n_pairs) > self.capacity


class _PairDistances(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, cell, indices, shifts, known):
        pos = positions.contiguous()
        geom = geometry_of(cell)
        count = _native.pair_count_of(indices)
        if known is not None:
            out = known.view_as(known)     # computed by the list construction from the same positions
        else:
            out = torch.zeros(indices.shape[0], dtype=pos.dtype, device=pos.device)
            _native.pair_distances(pos, geom.cell, indices, shifts, count, out)
        ctx.save_for_backward(pos, indices, shifts)
        ctx.cell_host, ctx.count = geom.cell, count
        ctx.cell_meta = (cell.dtype, cell.device)
        ctx.mark_non_differentiable(indices, shifts)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_d):
        pos, indices, shifts = ctx.saved_tensors
        g_pos4 = torch.zeros((pos.shape[0], 4), dtype=pos.dtype, device=pos.device)   # padded rows: vector reductions
        g_cell = torch.zeros(9, dtype=pos.dtype, device=pos.device) if ctx.needs_input_grad[1] else None
        _native.pair_distances_backward(pos, ctx.cell_host, indices, shifts, grad_d.contiguous(), ctx.count,
                                        g_pos4, g_cell)
        g_pos = g_pos4[:, :3]
        if g_cell is not None:
            g_cell = g_cell.view(3, 3).to(ctx.cell_meta[0])
        return g_pos if ctx.needs_input_grad[0] else None, g_cell, None, None, None


def distances_from(positions: torch.Tensor, cell: torch.Tensor, neighbor_indices: torch.Tensor,
                   shifts: torch.Tensor, known_distances: torch.Tensor | None = None) -> torch.Tensor:
    """
    ``|r_j + S . cell - r_i|``, differentiable in positions and cell.  On CUDA one kernel forward and one
    backward (``tpme_pair_distances``); ``known_distances`` (the values a list construction from the same
    positions returned) skips the forward kernel.  CPU tensors use the torch formulation.
    """
    if not positions.is_cuda:
        i, j = neighbor_indices[:, 0].long(), neighbor_indices[:, 1].long()
        delta = positions[j] - positions[i] + shifts.to(cell.dtype) @ cell
        return torch.linalg.norm(delta, dim=1)
    return _PairDistances.apply(positions, cell, neighbor_indices, shifts.contiguous(), known_distances)
