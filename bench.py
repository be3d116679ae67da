#!/usr/bin/env python3
"""
Benchmark of the PME / P3M hot path:  atom-steps/s, one *step* = forward of
``P3MCalculator/PMECalculator`` + backward of E = sum_i q_i V_i (forces = -dE/dpositions,
plus dE/d neighbor_distances), on the synthetic rock-salt crystals of SURVEY.md section 8(d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

Defaults (the contract configs of BASELINE.json):
  N = 1   headline = c3 (262144 atoms, PMECalculator Coulomb, 128^3, fp64: the largest single-GPU
          config); c2, c4, c5 and a shuffled-atom-order c3 are timed too and reported as
          sub-records (`other_workloads`) of the same JSON line.
  N > 1   (torchrun, one rank per GPU) headline = c4 (1 M atoms, P3M, 256^3, fp32) as ONE system whose
          mesh is slab-decomposed over the N GPUs (torchpme_b200.distributed: FFT transposes as NVLink
          peer stores or NCCL all-to-all, all-reduced potentials / forces) -- strong scaling; the
          same c4 step on one GPU and N independent c3 replicas are secondary fields.
Every measured workload first passes a PARITY GATE: one step against the CPU oracle
(oracle.calculator_step) on the same tensors; the line carries `parity` and the process exits
non-zero when fp64 exceeds 1e-5 or fp32 exceeds 1e-3 (relative to max |reference|).
`--impl reference` times the UNMODIFIED torch-pme (oracle/_ref/torchpme, copied there by
__graft_entry__.build) on the host cores with the same steps / warm-up.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (n_side, calculator, potential, dtype, n_mesh, BASELINE.json config it mirrors)
    "c2": dict(n_side=32, calc="p3m", pot=dict(kind="coulomb"), dtype="float32", n_mesh=64,
               label="c2: 32768-atom NaCl-like crystal, P3MCalculator Coulomb, 64^3 mesh, 4 nodes, fp32"),
    "c3": dict(n_side=64, calc="pme", pot=dict(kind="coulomb"), dtype="float64", n_mesh=128,
               label="c3: 262144 atoms, PMECalculator Coulomb, 128^3 mesh, 4 nodes, fp64"),
    "c4": dict(n_side=100, calc="p3m", pot=dict(kind="coulomb"), dtype="float32", n_mesh=256,
               label="c4: 1000000 atoms, P3MCalculator Coulomb, 256^3 mesh, 4 nodes, fp32"),
    "c5": dict(n_side=64, calc="pme", pot=dict(kind="ipl", exponent=6), dtype="float32", n_mesh=128,
               label="c5: 262144 atoms, PMECalculator InversePowerLaw p=6, 128^3 mesh, 4 nodes, fp32"),
}
SMEARING = 1.2
CUTOFF = 6.0
NODES = 4
PARITY_TOL = {"float64": 1e-5, "float32": 1e-3}   # north_star: relative to max |reference|
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def algorithmic_bytes(n, p, mesh, s, c=1, b=8):
    """algorithmic HBM bytes of every stage of one step."""
    m = mesh ** 3
    mh = mesh * mesh * (mesh // 2 + 1)
    R, K = c * m * s, c * mh * 2 * s
    A, O = (3 + c) * s * n, c * s * n
    return {
        "pair_forward": p * (2 * b + s) + 2 * O,
        "spread": A + R,
        "kfilter": 2 * R + 4 * K,
        "gather": R + 3 * s * n + O,
        "pair_backward": p * (2 * b + s) + p * s + 3 * O,
        "spread_grad": A + R,
        "kfilter_grad": 2 * R + 4 * K,
        "gather_vjp": R + A + O + 3 * s * n,
    }


KERNEL_OF_STAGE = {
    # stage -> (kernel the stage launches, launches of it per step)
    "pair_forward": ("pair_forward_kernel", 1), "pair_backward": ("pair_backward_kernel", 1),
    "spread": ("tile_spread4_kernel | spread_kernel", 2),
    "gather": ("tile_gather4_kernel | gather_point_kernel (values + dV/dr)", 1),
    "gather_vjp": ("tile_gather4_kernel | gather_point_kernel (vjp)", 1),
}


def dominant_kernel(stages, fft_launches):
    """
    The single kernel with the largest share of one step, from the per-stage timings.  The
    `kfilter` stage is one ABI call that launches `fft_launches` different FFT kernels (3: plane
    R2C, x pass . G, plane C2R; 5: z, y, x . G, y, z passes), each of them twice per step; it enters
    as one representative FFT kernel with 1/fft_launches of the stage's time and bytes.
    Returns (stage, kernel name, launches per step, ms per launch, algorithmic bytes per launch,
    share of the summed kernel time of the step).
    """
    per_step = {}
    for stage, st in stages.items():
        if stage == "kfilter":
            n = max(1, fft_launches)
            per_step[stage] = ("fft pass kernel (1 of %d per filter)" % n, 2, st["ms"] / n, st["alg_bytes"] / n,
                               2 * st["ms"])
        elif stage in KERNEL_OF_STAGE:
            name, launches = KERNEL_OF_STAGE[stage]
            per_step[stage] = (name, launches, st["ms"], st["alg_bytes"], launches * st["ms"])
    total = sum(v[4] for v in per_step.values())
    stage = max(per_step, key=lambda k: per_step[k][1] * per_step[k][2])
    name, launches, ms, alg, _ = per_step[stage]
    return stage, name, launches, ms, alg, launches * ms / total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def shuffle_inputs(pos, q, idx, d, seed=1):
    """
    Arbitrary atom order + a pair list sorted by its first index -- what an MD code and a cell-list
    neighbor search (vesin) hand over, instead of the lattice order / constant-offset pair runs of
    the synthetic generator (which flatter coalescing).
    """
    import torch

    n = pos.shape[0]
    gen = torch.Generator().manual_seed(seed)
    perm = torch.randperm(n, generator=gen).to(pos.device)          # new slot k holds old atom perm[k]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=pos.device)
    idx2 = inv[idx]
    order = torch.sort(idx2[:, 0], stable=True).indices
    return pos[perm].contiguous(), q[perm].contiguous(), idx2[order].contiguous(), d[order].contiguous()


def build_inputs(wl, device, shuffled=False, n_side=None):
    import torch
    from torchpme_b200.synthetic import rocksalt

    dtype = getattr(torch, wl["dtype"])
    n_side = n_side or wl["n_side"]
    pos, q, cell, idx, d = rocksalt(n_side, dtype=dtype, device=device, cutoff=CUTOFF)
    if shuffled:
        pos, q, idx, d = shuffle_inputs(pos, q, idx, d)
    length = float(cell[0, 0])
    # same mesh spacing for a reduced sample of the workload (n_side scaled, n_mesh scaled with it)
    n_mesh = wl["n_mesh"] * n_side // wl["n_side"] if n_side != wl["n_side"] else wl["n_mesh"]
    mesh_spacing = length / (n_mesh / 2 - 2)
    return dict(positions=pos, charges=q, cell=cell, neighbor_indices=idx, neighbor_distances=d,
                mesh_spacing=mesh_spacing, dtype=dtype, n_mesh=n_mesh)


def make_calculator(wl, mesh_spacing, device, slab_transport=None, module=None):
    tp = module
    if tp is None:
        import torchpme_b200 as tp
    if wl["pot"]["kind"] == "coulomb":
        pot = tp.CoulombPotential(smearing=SMEARING)
    else:
        pot = tp.InversePowerLawPotential(exponent=wl["pot"]["exponent"], smearing=SMEARING)
    if slab_transport is not None:
        from torchpme_b200.distributed import SlabP3MCalculator, SlabPMECalculator
        cls = SlabPMECalculator if wl["calc"] == "pme" else SlabP3MCalculator
        # every rank is handed its own chunk of the pair list (shard_pairs=False)
        return cls(pot.to(device), mesh_spacing=mesh_spacing, interpolation_nodes=NODES,
                   transport=slab_transport, shard_pairs=False)
    cls = tp.PMECalculator if wl["calc"] == "pme" else tp.P3MCalculator
    return cls(pot.to(device), mesh_spacing=mesh_spacing, interpolation_nodes=NODES)


# ----------------------------------------------------------------------------------------
# checkers / CPU arms: the numpy oracle (parity gate) and the unmodified reference (timing)
# ----------------------------------------------------------------------------------------
def oracle_step(wl, inputs_cpu):
    """one step of the CPU oracle (fp64) -> dict(V, dpos, dd)"""
    import numpy as np
    from oracle import pme_oracle as oracle

    pot = oracle.PotentialSpec(wl["pot"]["kind"], SMEARING, wl["pot"].get("exponent", 1))
    method = "Lagrange" if wl["calc"] == "pme" else "P3M"
    args = [np.ascontiguousarray(np.asarray(inputs_cpu[k], dtype=np.float64 if k != "neighbor_indices" else np.int64))
            for k in ("charges", "cell", "positions", "neighbor_indices", "neighbor_distances")]
    return oracle.calculator_step(pot, *args, inputs_cpu["mesh_spacing"], NODES, method)


def oracle_step_rate(wl, inputs_cpu, steps, warmup):
    import numpy as np
    from oracle import pme_oracle as oracle

    pot = oracle.PotentialSpec(wl["pot"]["kind"], SMEARING, wl["pot"].get("exponent", 1))
    method = "Lagrange" if wl["calc"] == "pme" else "P3M"
    args = [np.ascontiguousarray(inputs_cpu[k]) for k in
            ("charges", "cell", "positions", "neighbor_indices", "neighbor_distances")]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.calculator_step(pot, *args, inputs_cpu["mesh_spacing"], NODES, method)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    n = args[2].shape[0]
    total = sum(times)
    return n * len(times) / total, total / len(times)


def import_reference():
    """the unmodified torch-pme from oracle/_ref (None when the copy is absent)"""
    if not os.path.isdir(os.path.join(REF_DIR, "torchpme")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torchpme

    return torchpme if os.path.abspath(torchpme.__file__).startswith(os.path.abspath(REF_DIR)) else None


def reference_step_times(ref, wl, inputs, device, steps, warmup, sync=None):
    """seconds of `steps` reference steps (forward + backward of sum(q V)) after `warmup`"""
    import torch

    calc = make_calculator(wl, inputs["mesh_spacing"], device, module=ref).to(inputs["dtype"])
    q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
    times = []
    for it in range(warmup + steps):
        p = inputs["positions"].clone().requires_grad_(True)
        d = inputs["neighbor_distances"].clone().requires_grad_(True)
        if sync:
            sync()
        t0 = time.perf_counter()
        V = calc.forward(q, cell, p, idx, d)
        (V * q).sum().backward()
        if sync:
            sync()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def parity_check(wl, inputs, energy, V, g_pos, g_d, pair_slice=None):
    """
    The gate of SURVEY.md section 8(d): potentials, forces and dE/dd of one step against the fp64 CPU
    oracle on the same tensors, relative to max |reference| (and in L2 for the forces).  For the
    Lagrange (PME) stencils in fp32 the weights are only C0 across a stencil switch, so the handful of
    atoms within 1e-4 mesh units of one are reported separately (`forces_max_all`) and excluded from
    the gated maximum, as section 8(d) prescribes.
    """
    import numpy as np

    cpu = {k: inputs[k].detach().cpu().numpy() for k in
           ("positions", "charges", "cell", "neighbor_indices", "neighbor_distances")}
    cpu["mesh_spacing"] = inputs["mesh_spacing"]
    t0 = time.perf_counter()
    ref = oracle_step(wl, cpu)
    seconds = time.perf_counter() - t0
    f = lambda t: t.detach().cpu().double().numpy()  # noqa: E731
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())  # noqa: E731
    Vn, Fn, Dn = f(V), f(g_pos), f(g_d)
    dd_ref = ref["dd"] if pair_slice is None else ref["dd"][pair_slice[0]:pair_slice[1]]
    out = {"V": rel(Vn, ref["V"]), "forces_max_all": rel(Fn, ref["dpos"]),
           "forces_L2_all": float(np.linalg.norm(Fn - ref["dpos"]) / np.linalg.norm(ref["dpos"])),
           "dd": rel(Dn, dd_ref) if Dn.size else 0.0,
           "energy": float(abs(float(energy) - float((ref["V"] * cpu["charges"]).sum())) /
                           abs(float((ref["V"] * cpu["charges"]).sum())))}
    keep = np.ones(Fn.shape[0], dtype=bool)
    if wl["calc"] == "pme" and wl["dtype"] == "float32":
        n_mesh = inputs["n_mesh"]
        u = cpu["positions"].astype(np.float64) @ np.linalg.inv(cpu["cell"].astype(np.float64)) * n_mesh
        frac = u - np.floor(u)
        keep = ~((np.minimum(frac, 1 - frac) < 1e-4).any(axis=1))
    out["forces_max"] = float(np.abs(Fn[keep] - ref["dpos"][keep]).max() / np.abs(ref["dpos"]).max())
    out["forces_L2"] = float(np.linalg.norm(Fn[keep] - ref["dpos"][keep]) / np.linalg.norm(ref["dpos"][keep]))
    out["atoms_near_stencil_switch_excluded"] = int((~keep).sum())
    tol = PARITY_TOL[wl["dtype"]]
    out["tolerance"] = tol
    out["against"] = f"oracle.calculator_step (fp64 numpy restatement of the reference, {seconds:.1f} s on the host)"
    out["passed"] = bool(all(out[k] <= tol for k in ("V", "forces_max", "forces_L2", "dd", "energy")))
    return out


# ----------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------
def run_reference(args, wl):
    """
    `--impl reference`: the UNMODIFIED torch-pme (oracle/_ref/torchpme) on the host cores, all
    threads, same workload, steps and warm-up as the repo arm.  When the full workload would take
    longer than the budget the run keeps K and W and shrinks the sample: the same crystal at half
    the box length (1/8 of the atoms and of the mesh, same density / mesh spacing / pairs per atom).
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    ref = import_reference()
    cores = os.cpu_count()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    inputs = build_inputs(wl, "cpu")
    sample = "full workload"
    if ref is None:
        kind = "port"
        cpu = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in inputs.items() if k not in ("dtype",)}
        steps, warmup = min(steps, 3), min(warmup, 1)
        rate, sec = oracle_step_rate(wl, cpu, steps, warmup)
        sample = (f"oracle/_ref absent: {steps} full steps of the numpy/scipy oracle port after {warmup} warm-up "
                  "(about 1.7x slower than torch-pme itself on the build container)")
        threads = cores
    else:
        kind = "reference"
        threads = torch.get_num_threads()
        probe = reference_step_times(ref, wl, inputs, "cpu", 1, 1)[0]
        budget = float(os.environ.get("TPME_REFERENCE_BUDGET_S", "150"))
        if probe * (steps + warmup) > budget and wl["n_side"] % 2 == 0:
            inputs = build_inputs(wl, "cpu", n_side=wl["n_side"] // 2)
            sample = (f"bounded sample: the same crystal at half the box length ({inputs['positions'].shape[0]} atoms, "
                      f"{inputs['n_mesh']}^3 mesh, same density, mesh spacing and pairs per atom); one full-size step "
                      f"took {probe:.2f} s")
        times = reference_step_times(ref, wl, inputs, "cpu", steps, warmup)
        n = inputs["positions"].shape[0]
        sec = sum(times) / len(times)
        rate = n / sec
        sample += f"; {steps} steps after {warmup} warm-up, unmodified torch-pme {ref.__version__}, torch {torch.__version__} CPU ops"
    line = {
        "impl": "reference", "metric": "atom-steps/sec (energy+forces)", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64" if wl["dtype"] == "float64" else "f32",
        "data": "synthetic", "config": config_of(wl, args, inputs["positions"].shape[0], inputs["neighbor_indices"].shape[0]),
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": cores, "threads": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_of(wl, args, n_atoms, n_pairs, extra=None):
    cfg = {"workload": wl["label"], "atoms": n_atoms, "pairs": n_pairs, "mesh": wl["n_mesh"],
           "smearing": SMEARING, "cutoff": CUTOFF, "interpolation_nodes": NODES}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
class Timer:
    """device time of K calls (CUDA events on the current stream, L2 flushed before each call)"""

    def __init__(self, device, world, dist):
        import torch

        self.torch, self.device, self.world, self.dist = torch, device, world, dist
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def __call__(self, run, steps, warmup, collective=True):
        torch, dist = self.torch, self.dist
        collective = collective and self.world > 1
        for _ in range(warmup):
            self.flush.zero_(); run()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if collective:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            self.flush.zero_()
            a.record(); run(); b.record()
        torch.cuda.synchronize()
        if collective:
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if collective:
            t = torch.tensor([total_ms], device=self.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t)
        return total_ms


def measure_step(wl, inputs, device, timer, steps, warmup, slab_transport=None, pair_slice=None, parity=True,
                 collective=True, rank=0):
    """graph-replayed step of one workload: parity gate, eager and graphed device time"""
    import torch

    import torchpme_b200 as tp
    from torchpme_b200 import _native

    calc = make_calculator(wl, inputs["mesh_spacing"], device, slab_transport)
    q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
    pos = inputs["positions"].clone().requires_grad_(True)
    d = inputs["neighbor_distances"].clone().requires_grad_(True)
    tp.set_nan_check(False)  # the guard is a host sync; the graphed step cannot contain it

    def step():
        V = calc(q, cell, pos, idx, d)
        energy = (V * q).sum()
        g_pos, g_d = torch.autograd.grad(energy, (pos, d))
        return energy, V, g_pos, g_d

    for _ in range(2):
        energy, V, g_pos, g_d = step()
    torch.cuda.synchronize()
    par = None
    if parity and rank == 0:
        full = inputs.get("full_pairs")   # slab runs: the oracle needs the whole pair list
        chk = dict(inputs)
        if full is not None:
            chk["neighbor_indices"], chk["neighbor_distances"] = full
        par = parity_check(wl, chk, energy, V, g_pos, g_d, pair_slice)
    if collective and timer.world > 1:
        timer.dist.barrier()      # the other ranks wait on the host while rank 0 runs the oracle
    launches_before = _native.launch_counter
    graph_error = None
    try:
        graphed = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"], warmup=1)
    except Exception as exc:  # e.g. a collective that cannot be captured: time the eager step instead
        if slab_transport is None:
            raise
        graphed, graph_error = None, f"{type(exc).__name__}: {exc}"[:300]
        torch.cuda.synchronize()
    if graphed is not None:
        launches_per_step = (_native.launch_counter - launches_before) // 2   # 1 warm-up + 1 captured step
    else:
        launches_before = _native.launch_counter
        step()
        launches_per_step = _native.launch_counter - launches_before
    eager_ms = timer(step, steps, warmup, collective)
    graph_ms = timer(graphed.replay, steps, warmup, collective) if graphed is not None else eager_ms
    return dict(calc=calc, graphed=graphed, graph_ms=graph_ms, eager_ms=eager_ms, parity=par,
                launches_per_step=launches_per_step, graph_error=graph_error, pos=pos, d=d, step=step)


def stage_timings(wl, inputs, device, timer, steps, alg):
    """each stage of the step alone (L2 flushed before every launch) + the dominant kernel's roofline"""
    import torch
    from torchpme_b200 import _native
    from torchpme_b200.mesh import geometry_of

    dtype = inputs["dtype"]
    q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
    pd, dd = inputs["positions"], inputs["neighbor_distances"]
    geom = geometry_of(cell)
    ns = geom.ns_mesh(inputs["mesh_spacing"])
    r2u = geom.r2u(ns)
    method = _native.METHOD_ID["Lagrange" if wl["calc"] == "pme" else "P3M"]
    kind = _native.GREEN_COULOMB if wl["pot"]["kind"] == "coulomb" else _native.GREEN_IPL
    expo = wl["pot"].get("exponent", 1)
    green = _native.make_green(kind, 1.0, geom.recip, geom.spacing(ns), SMEARING, 1.0, expo,
                               NODES if wl["calc"] == "p3m" else 0)
    ppot = _native.make_pair_potential(kind, SMEARING, 1.0, expo)
    tiles = _native.tile_sort(pd, r2u, ns, NODES, method)
    rho = _native.spread(pd, q, r2u, ns, NODES, method, tiles=tiles)
    phi, _ = _native.kfilter_apply(rho, green)
    stage_fns = {
        "pair_forward": lambda: _native.pair_forward(q, idx, dd, None, None, False, ppot),
        "spread": lambda: _native.spread(pd, q, r2u, ns, NODES, method, out=rho, tiles=tiles),
        "kfilter": lambda: _native.kfilter_apply(rho, green),
        "gather": lambda: _native.gather(phi, pd, r2u, NODES, method, True, True, tiles=tiles),
        "pair_backward": lambda: _native.pair_backward(q, idx, dd, None, None, q, False, ppot, False, True),
        "gather_vjp": lambda: _native.gather_vjp(phi, pd, q, r2u, NODES, method, want_values=True, tiles=tiles),
    }
    if tiles is not None:
        stage_fns["tile_sort"] = lambda: _native.tile_sort(pd, r2u, ns, NODES, method)
    peak, peak_src = measured_peak()
    stages = {}
    k = max(10, steps)
    s = 4 if dtype == torch.float32 else 8
    n = pd.shape[0]
    alg = dict(alg, tile_sort=(3 * s + 3 * s + 8 + 8 + 4 * s + 4) * n)   # pos x2, key/rank w+r, record + index
    for name, fn in stage_fns.items():
        ms = timer(fn, k, 3, collective=False) / k
        stages[name] = {"ms": round(ms, 5), "alg_bytes": alg[name],
                        "gbs": round(alg[name] / ms / 1e6, 1), "frac": round(alg[name] / ms / 1e6 / peak, 4)}
    plan = _native.get_plan(dtype, ns, 1, device)
    return stages, plan.own_fft, (tiles.plan.tx, tiles.plan.ty) if tiles is not None else None


def run_b200(args, wl_name):
    import torch
    import torch.distributed as dist

    import torchpme_b200 as tp
    from torchpme_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    slab = args.decomposition == "slab"
    if world > 1 or slab:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29555")
        dist.init_process_group("nccl", device_id=device, rank=rank, world_size=world)
    wl = WORKLOADS[wl_name]
    timer = Timer(device, world, dist)
    warm = max(3, args.warmup)
    transport = args.transport
    if slab and transport == "auto":
        ok = all(torch.cuda.can_device_access_peer(local_rank, p) for p in range(torch.cuda.device_count()) if p != local_rank)
        transport = "p2p" if (ok and world > 1) else "nccl"

    inputs = build_inputs(wl, device, shuffled=args.shuffled)
    dtype = inputs["dtype"]
    n_atoms = inputs["positions"].shape[0]
    n_pairs_total = inputs["neighbor_indices"].shape[0]
    pair_slice = None
    if slab:
        # one system over all ranks: replicated atoms, every rank keeps its chunk of the pair list
        from torchpme_b200.distributed import SlabLayout
        lo, hi = SlabLayout((world, world, 2), world, rank).pair_range(n_pairs_total)
        inputs["full_pairs"] = (inputs["neighbor_indices"], inputs["neighbor_distances"])
        inputs["neighbor_indices"] = inputs["neighbor_indices"][lo:hi].contiguous()
        inputs["neighbor_distances"] = inputs["neighbor_distances"][lo:hi].contiguous()
        pair_slice = (lo, hi)
    n_pairs = inputs["neighbor_indices"].shape[0]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.profiler_range:   # ncu --profile-from-start off: skip the synthetic-input construction
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    main = measure_step(wl, inputs, device, timer, args.steps, warm, transport if slab else None, pair_slice,
                        parity=not args.no_parity, rank=rank)
    calc, graphed, graph_ms, eager_ms = main["calc"], main["graphed"], main["graph_ms"], main["eager_ms"]
    MAIN.update(parity=main["parity"], launches_per_step=main["launches_per_step"], graph_error=main["graph_error"])
    if args.profiler_range:
        torch.cuda.profiler.stop()

    if args.profile:
        # kernel-level breakdown of a few eager steps (torch profiler / CUPTI), not a bench value
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                main["step"]()
            torch.cuda.synchronize()
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile)), exist_ok=True)
            with open(args.profile, "w") as f:
                f.write(f"# torch.profiler, 5 eager steps, workload {wl_name}, {world} GPU(s), "
                        f"decomposition {args.decomposition}, transport {transport}\n")
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))

    fused = None
    if not slab and not args.lean:
        # the step through calculator.energy_and_gradients(): one spread / filter / gather
        # (SURVEY.md section 8d "fused energy+forces path"), reported next to the autograd step
        try:
            q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
            g_fused = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"],
                                     warmup=1, fused_energy_gradients=True)
            fused_ms = timer(g_fused.replay, args.steps, warm)
            torch.cuda.synchronize()
            f_par = None
            if main["parity"] is not None:
                _, V, _, _ = main["step"]()
                f_par = parity_check(wl, inputs, g_fused.energy, V, g_fused.grad_positions, g_fused.grad_distances)
            fused = {"ms_per_step": fused_ms / args.steps,
                     "value": world * n_atoms * args.steps / (fused_ms * 1e-3), "unit": "atom-steps/s",
                     "parity": f_par,
                     "note": "energy, dE/dpositions, dE/ddistances from ONE spread / filter / gather "
                             "(calculator.energy_and_gradients; valid because the filter is self-adjoint); "
                             "not the headline, which is the general autograd step"}
            g_fused.release()
        except Exception as exc:
            fused = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- end to end through the public API: pinned host inputs -> device, forces -> host ----
    q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
    host = {k: inputs[k].detach().cpu().pin_memory() for k in
            ("positions", "charges", "cell", "neighbor_indices", "neighbor_distances")}
    h_forces = torch.empty((n_atoms, 3), dtype=dtype).pin_memory()
    h_energy = torch.empty((), dtype=dtype).pin_memory()
    nbytes = lambda t: t.numel() * t.element_size()  # noqa: E731
    h2d = sum(nbytes(host[k]) for k in ("positions", "charges", "neighbor_indices", "neighbor_distances"))
    d2h = nbytes(h_forces) + h_energy.element_size()
    e2e_variants, e2e_bytes = {}, {}

    def e2e_eager_step():
        c_pos = host["positions"].to(device, non_blocking=True).requires_grad_(True)
        c_q = host["charges"].to(device, non_blocking=True)
        c_cell = host["cell"].to(device, non_blocking=True)
        c_idx = host["neighbor_indices"].to(device, non_blocking=True)
        c_d = host["neighbor_distances"].to(device, non_blocking=True)
        V = calc(c_q, c_cell, c_pos, c_idx, c_d)
        energy = (V * c_q).sum()
        (g_pos,) = torch.autograd.grad(energy, (c_pos,))
        h_forces.copy_(g_pos, non_blocking=True)
        h_energy.copy_(energy.detach(), non_blocking=True)

    e2e_variants["reference API, int64 pair list, eager launches"] = timer(e2e_eager_step, args.steps, warm)
    e2e_bytes["reference API, int64 pair list, eager launches"] = h2d
    graphed_io = graphed_io32 = None
    if graphed is not None:
        # the graph itself reads the pinned host inputs and writes the pinned host outputs
        # (GraphedStep(host_io=True)): one replay is a complete host-to-host step
        graphed_io = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"],
                                    warmup=1, host_io=True)
        name = "reference API, int64 pair list, GraphedStep(host_io=True)"
        e2e_variants[name] = timer(graphed_io.replay, args.steps, warm)
        e2e_bytes[name] = h2d
        # same call with the int32 indices the kernels also accept: half the pair-list bytes
        idx32 = idx.to(torch.int32)
        graphed_io32 = tp.GraphedStep(calc, q, cell, inputs["positions"], idx32, inputs["neighbor_distances"],
                                      warmup=1, host_io=True)
        name = "reference API, int32 pair list, GraphedStep(host_io=True)"
        e2e_variants[name] = timer(graphed_io32.replay, args.steps, warm)
        e2e_bytes[name] = h2d - nbytes(host["neighbor_indices"]) // 2
    e2e_best = min(e2e_variants, key=e2e_variants.get)
    e2e_ms = e2e_variants[e2e_best]

    # positions-only H2D + neighbor list built on the device (build time included); the distances are
    # differentiable here, so the forces contain the real-space part too (a more complete step)
    e2e_nl = None
    if not slab and not args.lean:
        try:
            from torchpme_b200.neighbors import distances_from, neighbor_list

            def e2e_nl_step():
                c_pos = host["positions"].to(device, non_blocking=True).requires_grad_(True)
                c_q = host["charges"].to(device, non_blocking=True)
                nl_idx, nl_d0, nl_shifts = neighbor_list(c_pos.detach(), cell, CUTOFF, index_dtype=torch.int32)
                nl_d = distances_from(c_pos, cell, nl_idx, nl_shifts, known_distances=nl_d0)
                V = calc(c_q, cell, c_pos, nl_idx, nl_d)
                energy = (V * c_q).sum()
                (g_pos,) = torch.autograd.grad(energy, (c_pos,))
                h_forces.copy_(g_pos, non_blocking=True)
                h_energy.copy_(energy.detach(), non_blocking=True)

            k_nl = max(3, min(args.steps, 10))
            nl_eager_ms = timer(e2e_nl_step, k_nl, 2) / k_nl
            # the same step as one CUDA graph: fixed-capacity list, the pair count never leaves the device
            pstep = tp.GraphedPositionsStep(calc, q, cell, inputs["positions"], cutoff=CUTOFF, host_io=True, warmup=1)
            nl_ms = timer(pstep.replay, args.steps, warm) / args.steps
            torch.cuda.synchronize()
            # its forces are the complete ones (mesh + real space): check them against the same step launched
            # eagerly over the exact-size list
            p_ = inputs["positions"].detach()
            i_full, _, s_full = neighbor_list(p_, cell, CUTOFF)
            pr = p_.clone().requires_grad_(True)
            Vr = calc(q, cell, pr, i_full, distances_from(pr, cell, i_full, s_full))
            (gr,) = torch.autograd.grad(Vr, pr, grad_outputs=q)
            got = pstep.host["grad_positions"].to(device)
            nl_check = {"forces_vs_eager_exact_list": float((got - gr).abs().max() / gr.abs().max()),
                        "pairs": int(pstep.host["n_pairs"]), "pairs_generator": int(inputs["neighbor_indices"].shape[0]),
                        "capacity": pstep.capacity, "overflowed": pstep.overflowed()}
            del i_full, s_full, pr, Vr, gr, got
            e2e_nl = {"ms_per_step": nl_ms, "value": world * n_atoms / (nl_ms * 1e-3), "unit": "atom-steps/s",
                      "h2d_bytes_per_step": nbytes(host["positions"]) + nbytes(host["charges"]),
                      "d2h_bytes_per_step": d2h + 8,
                      "eager_ms_per_step": nl_eager_ms,
                      "check": nl_check,
                      "path": "positions + charges H2D only; GraphedPositionsStep(host_io=True): ONE CUDA graph that "
                              "copies positions / charges in, builds the half neighbor list on the device "
                              "(neighbors.DeviceNeighborList: counting sort + 2-pass cell-list search, fixed-capacity "
                              "buffers, pair count stays on the device), ties the distances to the positions "
                              "(distances_from), runs the calculator forward + backward and copies energy + complete "
                              "forces (mesh + real space) out; eager_ms_per_step is the same step launched from Python "
                              "with neighbor_list() (one host sync for the pair count)"}
        except Exception as exc:
            e2e_nl = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage roofline (rank 0) ----
    roofline, stages = None, None
    s_bytes = 4 if dtype == torch.float32 else 8
    alg = algorithmic_bytes(n_atoms, n_pairs_total, wl["n_mesh"], s_bytes)
    peak, peak_src = measured_peak()
    if rank == 0 and slab:
        step_gbs = sum(alg.values()) / (graph_ms / args.steps) / 1e6
        roofline = {"bound": "hbm", "kernel": "whole step over all ranks", "achieved": round(step_gbs, 1),
                    "peak": peak * world, "unit": "GB/s", "frac": round(step_gbs / (peak * world), 4),
                    "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                    "step_alg_bytes": sum(alg.values())}
    if rank == 0 and not slab:
        stages, fft_launches, tile = stage_timings(wl, inputs, device, timer, args.steps, alg)
        stage, kname, k_launches, k_ms, k_alg, k_share = dominant_kernel(stages, fft_launches)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
                traffic = json.load(f).get(wl_name, {}).get(stage)
            if traffic is not None and stage == "kfilter":
                traffic = traffic / max(1, fft_launches)
        except Exception:
            pass
        k_gbs = k_alg / k_ms / 1e6
        roofline = {"bound": "hbm", "kernel": kname, "stage": stage, "launches_per_step": k_launches,
                    "ms_per_launch": round(k_ms, 5), "alg_bytes_per_launch": int(k_alg),
                    "achieved": round(k_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(k_gbs / peak, 4),
                    "traffic": traffic, "share_of_step_kernel_time": round(k_share, 3), "peak_source": peak_src,
                    "note": "kernel timed alone with CUDA events, L2 flushed before every launch",
                    "step_alg_bytes": sum(alg.values()),
                    "step_frac": round(sum(alg.values()) / (graph_ms / args.steps) / 1e6 / peak, 4),
                    "mesh_tile": tile}

    # ---- CPU baseline and reference-on-GPU (rank 0, N = 1 only): the unmodified reference ----
    cpu_baseline = reference_cuda = None
    if rank == 0 and world == 1 and not slab and not args.no_cpu_baseline:
        ref = import_reference()
        torch.set_num_threads(os.cpu_count() or 1)
        if ref is not None:
            cpu_in = {k: (v.detach().cpu() if hasattr(v, "cpu") else v) for k, v in inputs.items()}
            n_cpu = 3 if n_atoms <= 300000 else 1
            times = reference_step_times(ref, wl, cpu_in, "cpu", n_cpu, 1)
            sec = sorted(times)[len(times) // 2]
            cpu_baseline = {"value": n_atoms / sec, "unit": "atom-steps/s", "cores": os.cpu_count(),
                            "threads": torch.get_num_threads(), "kind": "reference",
                            "sample": f"median of {n_cpu} full step(s) of the same workload after 1 warm-up: unmodified "
                                      f"torch-pme (oracle/_ref) on the host, {sec:.2f} s/step"}
            try:
                times = reference_step_times(ref, wl, inputs, device, 3, 2, sync=torch.cuda.synchronize)
                sec = sorted(times)[1]
                reference_cuda = {"value": n_atoms / sec, "unit": "atom-steps/s", "ms_per_step": sec * 1e3,
                                  "note": "informational: the unmodified torch-pme with device='cuda' (stock ATen ops) on "
                                          "this GPU, wall clock with synchronize, median of 3 after 2 warm-up"}
            except Exception as exc:
                reference_cuda = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
        else:
            cpu = {k: inputs[k].detach().cpu().numpy() for k in
                   ("positions", "charges", "cell", "neighbor_indices", "neighbor_distances")}
            cpu["mesh_spacing"] = inputs["mesh_spacing"]
            rate, sec = oracle_step_rate(wl, cpu, 1, 1)
            cpu_baseline = {"value": rate, "unit": "atom-steps/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"oracle/_ref absent: 1 full step of the numpy/scipy oracle port ({sec:.2f} s/step)"}

    # ---- the other contract workloads as sub-records (N = 1), single-GPU / replica references (N > 1) ----
    others = {}
    if graphed is not None:
        graphed.release()
    for g in (graphed_io, graphed_io32):
        if g is not None:
            g.release()
    del main
    torch.cuda.empty_cache()
    sub_steps = max(5, min(args.steps, 20))

    def sub_record(name, shuffled=False, parity=True, collective=False, replicas=1):
        w2 = WORKLOADS[name]
        inp = build_inputs(w2, device, shuffled=shuffled)
        m = measure_step(w2, inp, device, timer, sub_steps, 3, parity=parity and not args.no_parity,
                         collective=collective, rank=0 if not collective else rank)
        n2 = inp["positions"].shape[0]
        rec = {"workload": w2["label"] + (" [shuffled atom order, pair list sorted by i]" if shuffled else ""),
               "ms_per_step": m["graph_ms"] / sub_steps, "value": replicas * n2 * sub_steps / (m["graph_ms"] * 1e-3),
               "unit": "atom-steps/s", "eager_ms_per_step": m["eager_ms"] / sub_steps, "steps": sub_steps,
               "dtype": "f32" if inp["dtype"] == torch.float32 else "f64", "parity": m["parity"],
               "gpu_launches_per_step": m["launches_per_step"]}
        if m["graphed"] is not None:
            m["graphed"].release()
        del m, inp
        torch.cuda.empty_cache()
        return rec

    if not args.lean:
        if world == 1 and not slab:
            plan = [(wl_name + "_shuffled", wl_name, True)] + [(n, n, False) for n in ("c2", "c4", "c5", "c3") if n != wl_name]
            for key, name, sh in plan:
                try:
                    others[key] = sub_record(name, shuffled=sh)
                except Exception as exc:
                    others[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        elif slab and world > 1:
            # the same workload on ONE GPU (rank 0 alone; the denominator of the strong-scaling ratio) ...
            if rank == 0:
                try:
                    others["same_workload_on_1_gpu"] = sub_record(wl_name, parity=False)
                except Exception as exc:
                    others["same_workload_on_1_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            dist.barrier()
            # ... and N independent c3 replicas (weak scaling, no data-path collective): the throughput mode
            try:
                rec = sub_record("c3", parity=False, collective=True, replicas=world)
                rec["parallelism"] = f"{world} independent replicas, one per GPU"
                others["replicas_c3"] = rec
            except Exception as exc:
                others["replicas_c3"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    failed = False
    if rank == 0:
        per_step = graph_ms / args.steps
        one = 1 if slab else world
        line = {
            "metric": "atom-steps/sec (energy+forces)",
            "value": one * n_atoms * args.steps / (graph_ms * 1e-3),
            "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": per_step, "higher_is_better": True, "scaling": "strong" if slab else "weak",
            "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
            "config": config_of(wl, args, n_atoms, n_pairs_total, {
                "atom_order": "shuffled, pair list sorted by i" if args.shuffled else "lattice order (synthetic generator)",
                "step": "forward + backward of sum(q*V) w.r.t. positions and neighbor distances, "
                        "whole step replayed as one CUDA graph",
                "l2": "flushed (256 MiB write) before every timed step",
                "parallelism": (f"one system, mesh slab-decomposed over {world} GPU(s), transport={transport}"
                                f"{'' if MAIN['graph_error'] is None else ', eager launches (' + MAIN['graph_error'] + ')'}") if slab
                else ("independent replica per GPU" if world > 1 else "single GPU")}),
            "parity": MAIN["parity"],
            "eager": {"value": one * n_atoms * args.steps / (eager_ms * 1e-3), "ms_per_step": eager_ms / args.steps,
                      "note": "same step launched from Python without graph capture"},
            "e2e": {"value": one * n_atoms * args.steps / (e2e_ms * 1e-3), "unit": "atom-steps/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": e2e_bytes[e2e_best],
                    "d2h_bytes_per_step": d2h,
                    "path": f"torchpme_b200 public API with the reference's call signature (pair list given by the caller), "
                            f"fastest of the variants below ({e2e_best}): H2D of positions / charges / neighbor list from "
                            "pinned host memory, the step, D2H of forces + energy into pinned host memory, all inside the "
                            "timed region",
                    "variants_ms_per_step": {k: round(v / args.steps, 5) for k, v in e2e_variants.items()},
                    "device_neighbor_list": e2e_nl},
            "gpu_launches": MAIN["launches_per_step"] * args.steps,
            "gpu_launches_per_step": MAIN["launches_per_step"],
            "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline, "reference_cuda": reference_cuda,
            "clocks": clocks, "fused_energy_gradients": fused, "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
        gates = [MAIN["parity"]] + [o.get("parity") for o in others.values() if isinstance(o, dict)]
        if fused and isinstance(fused.get("parity"), dict):
            gates.append(fused["parity"])
        failed = any(g is not None and not g["passed"] for g in gates)
    if world > 1 or slab:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if failed:
        print("bench.py: PARITY GATE FAILED (see `parity` in the line above)", file=sys.stderr)
        sys.exit(3)


MAIN = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c3 on one GPU, c4 (slab-decomposed) on several")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate (profiling runs)")
    ap.add_argument("--lean", action="store_true", help="headline workload only: no sub-records / fused / device-list e2e")
    ap.add_argument("--shuffled", action="store_true", help="shuffled atom order, pair list sorted by i")
    ap.add_argument("--decomposition", default=None, choices=["replica", "slab"],
                    help="default: slab for N > 1 (one system over all GPUs), replica otherwise")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "p2p", "p2p-copy"])
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the steps (for ncu --profile-from-start off)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of 5 eager steps to this file")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.decomposition is None:
        args.decomposition = "slab" if world > 1 else "replica"
    if args.workload is None:
        args.workload = "c4" if (world > 1 or args.gpus > 1) and args.decomposition == "slab" else "c3"
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_b200(args, args.workload)


if __name__ == "__main__":
    main()
