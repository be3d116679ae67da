"""
Generate the golden fixtures in this directory by running the UNMODIFIED reference
(lab-cosmo/torch-pme, read-only at /root/reference) on seeded inputs, in fp64 on CPU.

    python tests/golden/make_golden.py

Runs only in the build container (the reference does not travel to the GPU box); the
resulting ``*.npz`` files are committed and are what ``tests/`` compares both the numpy
oracle and the CUDA path against.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from _reference_import import import_reference  # noqa: E402

from oracle.pme_oracle import neighbor_list  # noqa: E402  (test utility, not arithmetic)

tp = import_reference()
F64 = torch.float64


def random_system(seed, n_atoms, n_channels, triclinic, box=6.0):
    rng = np.random.default_rng(seed)
    cell = np.eye(3) * box
    if triclinic:
        cell = cell + rng.uniform(-0.12, 0.12, (3, 3)) * box
    frac = rng.uniform(0, 1, (n_atoms, 3))
    pos = frac @ cell
    # a few atoms outside the cell exercise the index wrap
    pos[: max(1, n_atoms // 10)] += cell[0] * 1.0 - cell[1] * 2.0
    q = rng.normal(size=(n_atoms, n_channels))
    q[:, 0] -= q[:, 0].mean()  # channel 0 neutral, others carry net charge
    return pos, q, cell


def make_potential(spec):
    kind = spec["kind"]
    kw = dict(
        smearing=spec["smearing"],
        prefactor=spec.get("prefactor", 1.0),
        exclusion_radius=spec.get("exclusion_radius"),
        exclusion_degree=spec.get("exclusion_degree", 1),
    )
    if kind == "coulomb":
        return tp.CoulombPotential(**kw)
    return tp.InversePowerLawPotential(exponent=spec["exponent"], **kw)


CALC_CASES = []


def add_case(name, **kw):
    kw["name"] = name
    CALC_CASES.append(kw)


# config c1 of BASELINE.json: CsCl, PME, fp64
add_case("c1_cscl_pme", system="cscl", calc="pme", nodes=4, pot=dict(kind="coulomb", smearing=0.2),
         mesh_spacing=0.05, cutoff=1.0, full=False)
add_case("cscl_p3m", system="cscl", calc="p3m", nodes=4, pot=dict(kind="coulomb", smearing=0.2),
         mesh_spacing=0.05, cutoff=1.0, full=True)
_seed = 100
for calc, nodes_list in (("pme", (3, 4, 5, 6, 7)), ("p3m", (1, 2, 3, 4, 5))):
    for nodes in nodes_list:
        _seed += 1
        add_case(f"rand_{calc}_n{nodes}_coulomb", system="random", seed=_seed, n_atoms=24, n_channels=1,
                 triclinic=True, calc=calc, nodes=nodes, pot=dict(kind="coulomb", smearing=0.9, prefactor=1.7),
                 mesh_spacing=0.8, cutoff=3.1, full=bool(nodes % 2))
for p in (1, 2, 3, 4, 5, 6):
    for calc in ("pme", "p3m"):
        _seed += 1
        add_case(f"rand_{calc}_ipl{p}", system="random", seed=_seed, n_atoms=20, n_channels=2,
                 triclinic=(p % 2 == 0), calc=calc, nodes=4, pot=dict(kind="ipl", exponent=p, smearing=1.0),
                 mesh_spacing=0.9, cutoff=3.3, full=False)
add_case("rand_pme_exclusion", system="random", seed=900, n_atoms=20, n_channels=1, triclinic=True,
         calc="pme", nodes=4, pot=dict(kind="coulomb", smearing=1.0, exclusion_radius=2.5, exclusion_degree=2),
         mesh_spacing=0.9, cutoff=3.3, full=False)
add_case("rand_p3m_larger", system="random", seed=901, n_atoms=200, n_channels=1, triclinic=True,
         calc="p3m", nodes=4, pot=dict(kind="coulomb", smearing=1.1), mesh_spacing=0.7, cutoff=4.0,
         full=False, box=11.0)
add_case("rand_pme_larger", system="random", seed=902, n_atoms=200, n_channels=3, triclinic=False,
         calc="pme", nodes=5, pot=dict(kind="coulomb", smearing=1.1), mesh_spacing=0.7, cutoff=4.0,
         full=True, box=11.0)


# 2-D periodic systems (the slab correction of potentials/coulomb.py:6-40, applied at
# calculators/pme.py:138-140): separate fixture file `periodic_cases.npz`
PERIODIC_CASES = [
    dict(name="slab_pme_xy", system="random", seed=950, n_atoms=24, n_channels=1, triclinic=False, calc="pme",
         nodes=4, pot=dict(kind="coulomb", smearing=0.9), mesh_spacing=0.8, cutoff=3.1, full=False,
         periodic=[True, True, False]),
    dict(name="slab_p3m_xz", system="random", seed=951, n_atoms=30, n_channels=2, triclinic=False, calc="p3m",
         nodes=4, pot=dict(kind="coulomb", smearing=1.0, prefactor=1.3), mesh_spacing=0.9, cutoff=3.3, full=True,
         periodic=[True, False, True]),
    dict(name="slab_pme_n5_yz", system="random", seed=952, n_atoms=20, n_channels=1, triclinic=False, calc="pme",
         nodes=5, pot=dict(kind="coulomb", smearing=0.9), mesh_spacing=0.8, cutoff=3.1, full=False,
         periodic=[False, True, True]),
]


def run_calc_case(case):
    if case["system"] == "cscl":
        pos = np.array([[0, 0, 0], [0.5, 0.5, 0.5]], dtype=np.float64)
        q = np.array([[1.0], [-1.0]])
        cell = np.eye(3)
    else:
        pos, q, cell = random_system(case["seed"], case["n_atoms"], case["n_channels"],
                                     case["triclinic"], case.get("box", 6.0))
    idx, d, _ = neighbor_list(pos, cell, case["cutoff"], full=case["full"])
    pot = make_potential(case["pot"])
    cls = tp.PMECalculator if case["calc"] == "pme" else tp.P3MCalculator
    calc = cls(pot, mesh_spacing=case["mesh_spacing"], interpolation_nodes=case["nodes"],
               full_neighbor_list=case["full"]).to(F64)

    t_pos = torch.tensor(pos, dtype=F64, requires_grad=True)
    t_q = torch.tensor(q, dtype=F64, requires_grad=True)
    t_cell = torch.tensor(cell, dtype=F64, requires_grad=True)
    t_d = torch.tensor(d, dtype=F64, requires_grad=True)
    t_idx = torch.tensor(idx)
    periodic = torch.tensor(case["periodic"]) if "periodic" in case else None
    V = calc.forward(t_q, t_cell, t_pos, t_idx, t_d, periodic=periodic)
    # a generic upstream gradient (not equal to the charges) exercises the full backward
    rng = np.random.default_rng(7)
    g = torch.tensor(rng.normal(size=q.shape), dtype=F64)
    (V * g).sum().backward()
    out = dict(
        positions=pos, charges=q, cell=cell, neighbor_indices=idx, neighbor_distances=d,
        grad_out=g.numpy(), V=V.detach().numpy(), dpos=t_pos.grad.numpy(), dq=t_q.grad.numpy(),
        dcell=t_cell.grad.numpy(), dd=t_d.grad.numpy(),
        ns_mesh=tp.lib.get_ns_mesh(t_cell.detach(), case["mesh_spacing"]).numpy(),
    )
    # energy-style backward, L = sum q V (the benchmark step)
    for t in (t_pos, t_q, t_cell, t_d):
        t.grad = None
    V2 = calc.forward(t_q, t_cell, t_pos, t_idx, t_d, periodic=periodic)
    (V2 * t_q.detach()).sum().backward()
    out.update(dpos_energy=t_pos.grad.numpy(), dd_energy=t_d.grad.numpy(), dcell_energy=t_cell.grad.numpy())
    return out


def run_block_cases():
    """L1 blocks: MeshInterpolator and KSpaceFilter on non-power-of-two meshes."""
    out = {}
    rng = np.random.default_rng(3)
    cell = np.eye(3) * 5.0 + rng.uniform(-0.5, 0.5, (3, 3))
    pos = rng.uniform(-2, 7, (30, 3))
    w = rng.normal(size=(30, 2))
    ns = np.array([9, 10, 12])
    out["cell"], out["positions"], out["weights"], out["ns"] = cell, pos, w, ns
    mesh_in = rng.normal(size=(2, 9, 10, 12))
    out["mesh_in"] = mesh_in
    for method, nodes_list in (("P3M", (1, 2, 3, 4, 5)), ("Lagrange", (3, 4, 5, 6, 7))):
        for nodes in nodes_list:
            mi = tp.lib.MeshInterpolator(torch.tensor(cell), torch.tensor(ns), nodes, method)
            tpos = torch.tensor(pos, requires_grad=True)
            mi.compute_weights(tpos)
            rho = mi.points_to_mesh(torch.tensor(w))
            vals = mi.mesh_to_points(torch.tensor(mesh_in))
            g = torch.tensor(rng.normal(size=vals.shape))
            if vals.requires_grad:  # P3M with one node has constant weights
                (vals * g).sum().backward()
            else:
                tpos.grad = torch.zeros_like(tpos)
            key = f"{method}_{nodes}"
            out[key + "_rho"] = rho.detach().numpy()
            out[key + "_vals"] = vals.detach().numpy()
            out[key + "_g"] = g.numpy()
            out[key + "_dpos"] = tpos.grad.numpy()
    # filters
    pot = tp.CoulombPotential(smearing=0.8)
    for fn, inn in (("ortho", "ortho"), ("backward", "forward"), ("forward", "backward"), ("backward", "backward")):
        kf = tp.lib.KSpaceFilter(torch.tensor(cell), torch.tensor(ns), pot, fft_norm=fn, ifft_norm=inn)
        out[f"filter_{fn}_{inn}"] = kf.forward(torch.tensor(mesh_in)).numpy()
        out["kfilter_coulomb"] = kf._kfilter.numpy()
    p3 = tp.lib.P3MKSpaceFilter(torch.tensor(cell), torch.tensor(ns), 4, pot, fft_norm="backward", ifft_norm="forward")
    out["filter_p3m"] = p3.forward(torch.tensor(mesh_in)).numpy()
    out["kfilter_p3m"] = p3._kfilter.numpy()
    for p in range(1, 7):
        ipl = tp.InversePowerLawPotential(exponent=p, smearing=0.8)
        kf = tp.lib.KSpaceFilter(torch.tensor(cell), torch.tensor(ns), ipl, fft_norm="backward", ifft_norm="forward")
        out[f"kfilter_ipl{p}"] = kf._kfilter.numpy()
    return out


def main_periodic():
    import json
    cases = {}
    for case in PERIODIC_CASES:
        res = run_calc_case(case)
        for k, v in res.items():
            cases[f"{case['name']}/{k}"] = v
        print(case["name"], "periodic =", case["periodic"], "V[0] =", res["V"][0])
    np.savez_compressed(os.path.join(HERE, "periodic_cases.npz"), **cases)
    with open(os.path.join(HERE, "periodic_cases.json"), "w") as f:
        json.dump(PERIODIC_CASES, f, indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "periodic":   # only the 2-D periodic fixtures
        return main_periodic()
    main_periodic()
    cases = {}
    for case in CALC_CASES:
        res = run_calc_case(case)
        for k, v in res.items():
            cases[f"{case['name']}/{k}"] = v
        print(case["name"], "N =", len(res["positions"]), "P =", len(res["neighbor_distances"]),
              "ns =", res["ns_mesh"], "V[0] =", res["V"][0])
    np.savez_compressed(os.path.join(HERE, "calculator_cases.npz"), **cases)
    import json
    with open(os.path.join(HERE, "calculator_cases.json"), "w") as f:
        json.dump(CALC_CASES, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "block_cases.npz"), **run_block_cases())
    print("written", os.listdir(HERE))


if __name__ == "__main__":
    main()
