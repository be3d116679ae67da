"""
Neighbor list on the GPU.

The reference takes ``neighbor_indices`` / ``neighbor_distances`` from the external ``vesin``
package (tests/helpers.py:240-275, examples/basic-usage.py:166-169); at 1 M atoms that CPU step and
the copy of its int64 pair list dwarf a sub-millisecond GPU evaluation.  :func:`neighbor_list`
builds the list on the device with a cell list: atoms are wrapped into the cell, binned into slabs
between lattice planes and sorted by bin with torch ops (plumbing); the search itself runs in two
kernels of ``libtorchpme_b200.so`` (count, fill; ``csrc/neighbors_core.h``).

Validated on a B200 against the brute-force oracle (``tests/test_gpu_neighbors.py``: identical pair
sets for cubic / triclinic / smaller-than-cutoff cells, half and full lists, fp32 / fp64, slab and open
boundaries, 1 M atoms) and on the CPU through a host build of the same search loop
(``tests/test_neighbors.py``).
"""

from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _native


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


def _native_search(dtype_id, wrapped, wrap_shift, atom_bins, order, bin_start, n, search, offsets=None,
                   outputs=None):
    """count (offsets is None) or fill pass through the C ABI"""
    lib = _native.load()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    if not wrapped.is_cuda:
        raise _native.NativeLibraryError(
            f"`positions` lives on {wrapped.device}; torchpme_b200 is a CUDA-only implementation "
            "(no CPU fallback). Move the inputs to a CUDA device.")
    stream = ctypes.c_void_p(torch.cuda.current_stream(wrapped.device).cuda_stream)
    with torch.cuda.device(wrapped.device):
        if offsets is None:
            counts = torch.empty(n, dtype=torch.int32, device=wrapped.device)
            _native._check(lib.tpme_neighbor_count(dtype_id, ptr(wrapped), ptr(wrap_shift), ptr(atom_bins),
                                                   ptr(order), ptr(bin_start), n, ctypes.byref(search),
                                                   ptr(counts), stream), "tpme_neighbor_count")
            _native._count()
            return counts
        indices, dist_sq, shifts = outputs
        _native._check(lib.tpme_neighbor_fill(dtype_id, ptr(wrapped), ptr(wrap_shift), ptr(atom_bins),
                                              ptr(order), ptr(bin_start), n, ctypes.byref(search),
                                              ptr(offsets), ptr(indices), ptr(dist_sq), ptr(shifts), stream),
                       "tpme_neighbor_fill")
        _native._count()
    return None


@torch.no_grad()
def neighbor_list(positions: torch.Tensor, cell: torch.Tensor, cutoff: float, full_neighbor_list: bool = False,
                  periodic=(True, True, True), _search=None):
    """
    All pairs ``(i, j, S)`` with ``|r_j + S . cell - r_i| < cutoff`` (except ``(i, i, 0)``): every
    unordered pair once (``i < j`` for any ``S``, self images with ``S`` lexicographically positive)
    or, with ``full_neighbor_list``, in both directions.

    Returns ``neighbor_indices (P, 2) int64``, ``neighbor_distances (P,)`` and ``shifts (P, 3) int32``
    on the device of ``positions``; the pairs are grouped by atom ``i`` in cell-list order.  One host
    synchronisation (the pair count).  For differentiable distances recompute them from the result
    with :func:`distances_from`.
    """
    if positions.dim() != 2 or positions.shape[1] != 3:
        raise ValueError(f"`positions` must be a tensor with shape [n_atoms, 3], got {list(positions.shape)}")
    if cell.shape != (3, 3):
        raise ValueError(f"`cell` must be a tensor with shape [3, 3], got {list(cell.shape)}")
    if not cutoff > 0:
        raise ValueError(f"`cutoff` must be positive, got {cutoff}")
    dtype, device, n = positions.dtype, positions.device, positions.shape[0]
    periodic = tuple(bool(p) for p in periodic)
    cell_np = cell.detach().to("cpu", torch.float64).numpy()
    n_bins, reach = search_layout(cell_np, cutoff, periodic)
    search = _NeighborSearch()
    for k, v in enumerate(cell_np.reshape(-1)):
        search.cell[k] = float(v)
    for a in range(3):
        search.n_bins[a], search.reach[a], search.periodic[a] = n_bins[a], reach[a], int(periodic[a])
    search.full_list, search.cutoff = int(full_neighbor_list), float(cutoff)

    # ---- plumbing: wrap, bin, sort (torch ops) ------------------------------------------------
    cell64 = cell.detach().to(torch.float64)
    frac = positions.detach().to(torch.float64) @ torch.linalg.inv(cell64)
    per = torch.tensor(periodic, device=device)
    k = torch.where(per, torch.floor(frac), torch.zeros_like(frac))
    fw = (frac - k).clamp_(0.0, 1.0)
    wrapped = (positions.detach().to(torch.float64) - k @ cell64).to(dtype).contiguous()
    nb = torch.tensor(n_bins, device=device, dtype=torch.float64)
    bins = torch.minimum((fw * nb).to(torch.int64), (nb - 1).to(torch.int64))
    bins = torch.where(per, bins, torch.zeros_like(bins))
    linear = (bins[:, 0] * n_bins[1] + bins[:, 1]) * n_bins[2] + bins[:, 2]
    sorted_linear, order = torch.sort(linear, stable=True)
    total_bins = n_bins[0] * n_bins[1] * n_bins[2]
    bin_start = torch.searchsorted(sorted_linear, torch.arange(total_bins + 1, device=device)).to(torch.int32)
    wrap_shift = k.to(torch.int32).contiguous()
    atom_bins = bins.to(torch.int32).contiguous()
    order = order.to(torch.int32).contiguous()

    # ---- two-pass search ----------------------------------------------------------------------
    run = _search or _native_search
    dtype_id = 0 if dtype == torch.float32 else 1
    counts = run(dtype_id, wrapped, wrap_shift, atom_bins, order, bin_start, n, search)
    offsets = torch.cumsum(counts.to(torch.int64), 0) - counts.to(torch.int64)
    n_pairs = int(counts.sum()) if n else 0
    indices = torch.empty((n_pairs, 2), dtype=torch.int64, device=device)
    dist_sq = torch.empty(n_pairs, dtype=dtype, device=device)
    shifts = torch.empty((n_pairs, 3), dtype=torch.int32, device=device)
    if n_pairs:
        run(dtype_id, wrapped, wrap_shift, atom_bins, order, bin_start, n, search, offsets.contiguous(),
            (indices, dist_sq, shifts))
    return indices, torch.sqrt(dist_sq), shifts


def distances_from(positions: torch.Tensor, cell: torch.Tensor, neighbor_indices: torch.Tensor,
                   shifts: torch.Tensor) -> torch.Tensor:
    """``|r_j + S . cell - r_i|`` with plain torch ops (differentiable in positions and cell)"""
    delta = positions[neighbor_indices[:, 1]] - positions[neighbor_indices[:, 0]] + shifts.to(cell.dtype) @ cell
    return torch.linalg.norm(delta, dim=1)
