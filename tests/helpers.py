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


def _np(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def rel_err(a, b):
    a, b = _np(a), _np(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


from torchpme_b200.synthetic import rocksalt  # noqa: E402,F401
