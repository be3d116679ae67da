"""
Synthetic inputs for benchmarks and tests: the "random NaCl-like crystal" of SURVEY.md
section 8(d) with a half neighbor list built from the lattice topology on the device.
"""

from __future__ import annotations

import math

import torch


def rocksalt(n_side: int, dtype=torch.float64, device="cpu", jitter: float = 0.1, seed: int = 0,
             d0: float = 2.82, cutoff: float = 6.0):
    """
    ``n_side^3`` sites at spacing ``d0`` in a cubic cell, charges (-1)^(ix+iy+iz), Gaussian
    jitter (seeded, fp64, wrapped into the cell).  The half neighbor list holds every pair
    closer than ``cutoff`` once; candidates are the lattice offsets within cutoff + 6 sigma,
    distances use the minimum image (requires cutoff < L/2).

    Returns ``positions (N,3), charges (N,1), cell (3,3), neighbor_indices (P,2) int64,
    neighbor_distances (P,)`` on ``device`` (floats in ``dtype``).
    """
    gen = torch.Generator().manual_seed(seed)
    length = n_side * d0
    assert cutoff < length / 2
    ar = torch.arange(n_side)
    sites = torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3)
    pos = sites.to(torch.float64) * d0 + jitter * torch.randn(sites.shape, generator=gen, dtype=torch.float64)
    pos = pos % length
    charges = (1.0 - 2.0 * (sites.sum(1) % 2).to(torch.float64)).reshape(-1, 1)
    cell = torch.eye(3, dtype=torch.float64) * length

    pos_d, sites_d = pos.to(device), sites.to(device)
    margin = cutoff + 6 * jitter
    reach = int(math.ceil(margin / d0))
    rng = torch.arange(-reach, reach + 1)
    offs = torch.stack(torch.meshgrid(rng, rng, rng, indexing="ij"), -1).reshape(-1, 3)
    positive = (offs[:, 0] > 0) | ((offs[:, 0] == 0) & (offs[:, 1] > 0)) | (
        (offs[:, 0] == 0) & (offs[:, 1] == 0) & (offs[:, 2] > 0))
    offs = offs[positive]
    offs = offs[offs.to(torch.float64).norm(dim=1) * d0 < margin + 1e-9]
    base = torch.arange(sites.shape[0], device=device)
    out_i, out_j, out_d = [], [], []
    for off in offs.to(device):
        nb = (sites_d + off) % n_side
        j = (nb[:, 0] * n_side + nb[:, 1]) * n_side + nb[:, 2]
        delta = pos_d[j] - pos_d
        delta = delta - torch.round(delta / length) * length
        dist = delta.norm(dim=1)
        keep = dist < cutoff
        out_i.append(base[keep])
        out_j.append(j[keep])
        out_d.append(dist[keep])
    idx = torch.stack([torch.cat(out_i), torch.cat(out_j)], 1).contiguous()
    return (pos_d.to(dtype), charges.to(device=device, dtype=dtype), cell.to(device=device, dtype=dtype),
            idx, torch.cat(out_d).to(dtype))
