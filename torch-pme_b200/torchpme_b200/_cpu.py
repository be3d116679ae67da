"""
CPU tensors: the same four stages written with differentiable torch ops.

BASELINE config c1 ("CsCl, PMECalculator fp64 on CPU") and the reference's own basic usage
(``examples/basic-usage.py``, ``tests/calculators/test_workflow.py:112-123``: output device ==
input device) are CPU cases, so the package dispatches on the device of its inputs:

* CUDA tensors  -> always the sm_100a kernels of ``libtorchpme_b200.so`` (a missing library raises
  ``NativeLibraryError``; nothing here is ever used for a CUDA tensor);
* CPU tensors   -> this module.

This is the builder's own formulation, not the oracle (``oracle/`` is test infrastructure and is
never imported by the package) and not the reference's code: one flattened node index per
(point, stencil node), a single ``index_add_`` / gather on the flattened mesh, weight polynomials
evaluated from the same integer tables the CUDA kernels are generated from
(``csrc/gen_weights.py``), ``torch.fft`` for the transforms.  Autograd differentiates it, so every
gradient the CUDA path provides analytically (positions, charges, cell, distances, potential
parameters) -- and ``create_graph=True`` double backward -- is available on the CPU.
Semantics follow ``lib/mesh_interpolator.py:303-457`` and ``calculators/calculator.py:43-87``.
"""

from __future__ import annotations

import torch

# 1-D weight polynomials: numerators in ascending powers of the offset x in [-1/2, 1/2] over a common
# denominator -- the tables of csrc/gen_weights.py (lib/mesh_interpolator.py:171-209, 228-300)
_TABLE = {
    ("P3M", 1): (1, [[1]]),
    ("P3M", 2): (2, [[1, -2], [1, 2]]),
    ("P3M", 3): (8, [[1, -4, 4], [6, 0, -8], [1, 4, 4]]),
    ("P3M", 4): (48, [[1, -6, 12, -8], [23, -30, -12, 24], [23, 30, -12, -24], [1, 6, 12, 8]]),
    ("P3M", 5): (384, [[1, -8, 24, -32, 16], [76, -176, 96, 64, -64], [230, 0, -240, 0, 96],
                       [76, 176, 96, -64, -64], [1, 8, 24, 32, 16]]),
    ("Lagrange", 3): (2, [[0, -1, 1], [2, 0, -2], [0, 1, 1]]),
    ("Lagrange", 4): (48, [[-3, 2, 12, -8], [27, -54, -12, 24], [27, 54, -12, -24], [-3, -2, 12, 8]]),
    ("Lagrange", 5): (24, [[0, 2, -1, -2, 1], [0, -16, 16, 4, -4], [24, 0, -30, 0, 6],
                           [0, 16, 16, -4, -4], [0, -2, -1, 2, 1]]),
    ("Lagrange", 6): (3840, [[45, -18, -200, 80, 80, -32], [-375, 250, 1560, -1040, -240, 160],
                             [2250, -4500, -1360, 2720, 160, -320], [2250, 4500, -1360, -2720, 160, 320],
                             [-375, -250, 1560, 1040, -240, -160], [45, 18, -200, -80, 80, 32]]),
    ("Lagrange", 7): (720, [[0, -12, 4, 15, -5, -3, 1], [0, 108, -54, -120, 60, 12, -6],
                            [0, -540, 540, 195, -195, -15, 15], [720, 0, -980, 0, 280, 0, -20],
                            [0, 540, 540, -195, -195, 15, 15], [0, -108, -54, 120, 60, -12, -6],
                            [0, 12, 4, -15, -5, 3, 1]]),
}


def _weights_1d(x: torch.Tensor, nodes: int, method: str) -> torch.Tensor:
    """(N, 3) offsets -> (N, 3, nodes) weights: one Horner evaluation with a (nodes, degree) coefficient matrix"""
    den, rows = _TABLE[(method, nodes)]
    coef = torch.tensor(rows, dtype=x.dtype, device=x.device) / den          # (nodes, degree + 1)
    w = x[..., None] * 0 + coef[:, -1]      # stays connected to x (zero gradient for constant weights)
    for k in range(coef.shape[1] - 2, -1, -1):
        w = w * x[..., None] + coef[:, k]
    return w


def stencil(positions: torch.Tensor, cell: torch.Tensor, ns, nodes: int, method: str):
    """
    flattened mesh index (N, nodes^3) int64 and weight (N, nodes^3) of every stencil node of every
    point; differentiable in positions and cell through the weights (the integer base is not)
    """
    nx, ny, nz = (int(v) for v in ns)
    ns_t = torch.tensor([nx, ny, nz], dtype=positions.dtype, device=positions.device)
    u = (positions @ torch.linalg.inv(cell)) * ns_t
    if nodes % 2 == 0:
        base = torch.floor(u.detach())
        x = u - (base + 0.5)
    else:
        base = torch.round(u.detach())          # ties to even, like the reference
        x = u - base
    w = _weights_1d(x, nodes, method)                                          # (N, 3, n)
    first = base.to(torch.int64) + (1 - (nodes + 1) // 2)
    offs = torch.arange(nodes, device=positions.device)
    node = (first[:, :, None] + offs) % torch.tensor([nx, ny, nz], device=positions.device)[None, :, None]
    flat = ((node[:, 0, :, None, None] * ny + node[:, 1, None, :, None]) * nz + node[:, 2, None, None, :])
    weight = w[:, 0, :, None, None] * w[:, 1, None, :, None] * w[:, 2, None, None, :]
    n = positions.shape[0]
    return flat.reshape(n, -1), weight.reshape(n, -1)


def spread(flat: torch.Tensor, weight: torch.Tensor, values: torch.Tensor, ns) -> torch.Tensor:
    """mesh[c, m] = sum_i values[i, c] * weight[i, m]  ->  (C, nx, ny, nz)"""
    nx, ny, nz = (int(v) for v in ns)
    c = values.shape[1]
    contrib = (values.T[:, :, None] * weight[None, :, :]).reshape(c, -1)      # (C, N n^3)
    mesh = torch.zeros((c, nx * ny * nz), dtype=values.dtype, device=values.device)
    mesh = mesh.index_add(1, flat.reshape(-1), contrib)
    return mesh.reshape(c, nx, ny, nz)


def gather(flat: torch.Tensor, weight: torch.Tensor, mesh: torch.Tensor) -> torch.Tensor:
    """values[i, c] = sum_m mesh[c, m] * weight[i, m]  ->  (N, C)"""
    c = mesh.shape[0]
    picked = mesh.reshape(c, -1)[:, flat]                                       # (C, N, n^3)
    return (picked * weight[None]).sum(dim=2).T


def kfilter(mesh: torch.Tensor, table: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * irfftn(table * rfftn(mesh)) with unnormalised transforms"""
    dims = (1, 2, 3)
    hat = torch.fft.rfftn(mesh, dim=dims, norm="backward")
    return torch.fft.irfftn(hat * table, s=mesh.shape[1:], dim=dims, norm="forward") * scale


def pair_sum(charges: torch.Tensor, neighbor_indices: torch.Tensor, pair_values: torch.Tensor,
             full_neighbor_list: bool) -> torch.Tensor:
    """out[i, c] = 1/2 sum_pairs q[j, c] v_pair (+ the mirrored term for half lists)"""
    i, j = neighbor_indices[:, 0].long(), neighbor_indices[:, 1].long()
    out = torch.zeros_like(charges).index_add(0, i, charges[j] * pair_values[:, None])
    if not full_neighbor_list:
        out = out.index_add(0, j, charges[i] * pair_values[:, None])
    return out / 2
