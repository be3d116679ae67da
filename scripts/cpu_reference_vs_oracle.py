"""
Build container only (needs /root/reference): time the UNMODIFIED reference (torch CPU ops, all cores)
and the numpy oracle on the same c2 inputs, to calibrate the oracle as the CPU baseline of bench.py.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from _reference_import import import_reference  # noqa: E402
from oracle import pme_oracle as oracle  # noqa: E402
from torchpme_b200.synthetic import rocksalt  # noqa: E402


def main():
    ref = import_reference()
    n_side, n_mesh, dtype = 32, 64, torch.float32
    pos, q, cell, idx, d = rocksalt(n_side, dtype=dtype)
    mesh_spacing = float(cell[0, 0]) / (n_mesh / 2 - 2)
    torch.set_num_threads(os.cpu_count())
    calc = ref.P3MCalculator(ref.CoulombPotential(smearing=1.2), mesh_spacing=mesh_spacing, interpolation_nodes=4)
    calc.to(dtype)
    times = []
    for it in range(5):
        p = pos.clone().requires_grad_(True)
        dd = d.clone().requires_grad_(True)
        t0 = time.perf_counter()
        V = calc.forward(q, cell, p, idx, dd)
        (V * q).sum().backward()
        times.append(time.perf_counter() - t0)
    t_ref = float(np.median(times[2:]))
    spec = oracle.PotentialSpec("coulomb", 1.2)
    args = [np.ascontiguousarray(t.numpy()) for t in (q, cell, pos, idx, d)]
    times = []
    for it in range(4):
        t0 = time.perf_counter()
        oracle.calculator_step(spec, *args, mesh_spacing, 4, "P3M")
        times.append(time.perf_counter() - t0)
    t_or = float(np.median(times[1:]))
    n = pos.shape[0]
    print(f"c2 on {os.cpu_count()} cores: reference (torch {torch.__version__}, {torch.get_num_threads()} threads) "
          f"{t_ref * 1e3:.0f} ms/step = {n / t_ref:.3g} atom-steps/s; numpy oracle {t_or * 1e3:.0f} ms/step = "
          f"{n / t_or:.3g} atom-steps/s; oracle / reference speed = {t_ref / t_or:.2f}")


if __name__ == "__main__":
    main()
