"""Shared test utilities: golden fixtures, oracle potential specs, synthetic crystals."""
import json
import os

import numpy as np
import torch

from oracle import pme_oracle as oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_calculator_cases():
    data = np.load(os.path.join(GOLDEN, "calculator_cases.npz"))
    with open(os.path.join(GOLDEN, "calculator_cases.json")) as f:
        cases = json.load(f)
    return cases, data


def case_arrays(data, name):
    prefix = name + "/"
    return {k[len(prefix):]: data[k] for k in data.files if k.startswith(prefix)}


def oracle_potential(spec):
    return oracle.PotentialSpec(spec["kind"], spec["smearing"], spec.get("exponent", 1),
                                spec.get("prefactor", 1.0), spec.get("exclusion_radius"),
                                spec.get("exclusion_degree", 1))


def b200_potential(tp, spec, device=None, dtype=None):
    kw = dict(smearing=spec["smearing"], prefactor=spec.get("prefactor", 1.0),
              exclusion_radius=spec.get("exclusion_radius"),
              exclusion_degree=spec.get("exclusion_degree", 1))
    if spec["kind"] == "coulomb":
        pot = tp.CoulombPotential(**kw)
    else:
        pot = tp.InversePowerLawPotential(exponent=spec["exponent"], **kw)
    return pot.to(device=device, dtype=dtype)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rocksalt(n_side, dtype=torch.float64, device="cpu", jitter=0.1, seed=0, d0=2.82, cutoff=6.0):
    """
    Synthetic "random NaCl-like crystal" of SURVEY.md section 8(d): n_side^3 sites, charges
    (-1)^(ix+iy+iz), Gaussian jitter, half neighbor list within `cutoff` built from the
    lattice topology (candidate offsets up to +-3 sites) with torch ops on `device`.
    Returns positions, charges, cell, neighbor_indices (P,2) int64, neighbor_distances.
    """
    g = torch.Generator().manual_seed(seed)
    L = n_side * d0
    ar = torch.arange(n_side)
    ix, iy, iz = torch.meshgrid(ar, ar, ar, indexing="ij")
    sites = torch.stack([ix, iy, iz], -1).reshape(-1, 3)
    pos = sites.to(torch.float64) * d0 + jitter * torch.randn(sites.shape, generator=g, dtype=torch.float64)
    pos = pos % L
    q = (1.0 - 2.0 * ((sites.sum(1)) % 2).to(torch.float64)).reshape(-1, 1)
    cell = torch.eye(3, dtype=torch.float64) * L
    pos_d, sites_d = pos.to(device), sites.to(device)
    reach = int(np.ceil((cutoff + 6 * jitter) / d0))
    rng = torch.arange(-reach, reach + 1)
    offs = torch.stack(torch.meshgrid(rng, rng, rng, indexing="ij"), -1).reshape(-1, 3)
    # half list: keep lexicographically positive offsets only
    keep = (offs[:, 0] > 0) | ((offs[:, 0] == 0) & (offs[:, 1] > 0)) | (
        (offs[:, 0] == 0) & (offs[:, 1] == 0) & (offs[:, 2] > 0))
    offs = offs[keep]
    offs = offs[(offs.to(torch.float64).norm(dim=1) * d0) < cutoff + 6 * jitter + 1e-9]
    n = sites.shape[0]
    base = torch.arange(n, device=device)
    out_i, out_j, out_d = [], [], []
    for o in offs.to(device):
        nb = (sites_d + o) % n_side
        j = (nb[:, 0] * n_side + nb[:, 1]) * n_side + nb[:, 2]
        delta = pos_d[j] - pos_d
        delta = delta - torch.round(delta / L) * L  # minimum image (cutoff << L/2)
        d = delta.norm(dim=1)
        m = d < cutoff
        out_i.append(base[m]); out_j.append(j[m]); out_d.append(d[m])
    idx = torch.stack([torch.cat(out_i), torch.cat(out_j)], 1)
    dist = torch.cat(out_d)
    return (pos_d.to(dtype), q.to(device=device, dtype=dtype), cell.to(device=device, dtype=dtype),
            idx, dist.to(dtype))
