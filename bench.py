#!/usr/bin/env python3
"""
Benchmark of the PME / P3M hot path:  atom-steps/s, one *step* = forward of
``P3MCalculator/PMECalculator`` + backward of E = sum_i q_i V_i (forces = -dE/dpositions,
plus dE/d neighbor_distances), on the synthetic rock-salt crystals of SURVEY.md section 8(d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU.  Default (`--decomposition replica`): every
rank runs an independent replica of the workload (weak scaling, no data-path collective) --
the meshes of c2/c3/c5 are too small to shard.  `--decomposition slab` runs ONE system whose
mesh is cut into x slabs over the ranks (torchpme_b200.distributed: all-to-all FFT transposes
over NCCL or NVLink peer stores, all-reduced potentials / forces; strong scaling) -- meant for
c4.  Rank 0 prints ONE JSON line.
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
# the Python reference cannot travel to the GPU box; measured where it can run (build container, 8 cores,
# scripts/cpu_reference_vs_oracle.py, workload c2): reference 127 ms/step, this port 214 ms/step
ORACLE_CALIBRATION = ("calibration: the unmodified torch-pme reference with 8 torch threads runs c2 1.7x faster "
                      "than this port on the build container")


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
    "spread": ("spread_kernel", 2), "gather": ("gather_point_kernel (values + dV/dr)", 1),
    "gather_vjp": ("gather_point_kernel (vjp)", 1),
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
        else:
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


def build_inputs(wl, device):
    import torch
    from torchpme_b200.synthetic import rocksalt

    dtype = getattr(torch, wl["dtype"])
    pos, q, cell, idx, d = rocksalt(wl["n_side"], dtype=dtype, device=device, cutoff=CUTOFF)
    length = float(cell[0, 0])
    mesh_spacing = length / (wl["n_mesh"] / 2 - 2)
    return dict(positions=pos, charges=q, cell=cell, neighbor_indices=idx, neighbor_distances=d,
                mesh_spacing=mesh_spacing, dtype=dtype)


def make_calculator(wl, mesh_spacing, device, slab_transport=None):
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
# reference arm / cpu baseline: the numpy oracle on the host cores
# ----------------------------------------------------------------------------------------
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


def run_reference(args, wl):
    """`--impl reference`: the CPU restatement of the reference path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    inputs = build_inputs(wl, "cpu")
    cpu = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in inputs.items() if k != "dtype"}
    steps = max(1, min(args.steps, 5))
    warmup = 1
    rate, sec = oracle_step_rate(wl, cpu, steps, warmup)
    cores = os.cpu_count()
    sample = (f"{steps} full steps of the workload after {warmup} warm-up (numpy/scipy oracle, scipy.fft workers=all "
              f"cores; {ORACLE_CALIBRATION})")
    line = {
        "impl": "reference", "metric": "atom-steps/sec (energy+forces)", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if wl["dtype"] == "float64" else "f32",
        "data": "synthetic", "config": {"workload": wl["label"]},
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
def run_b200(args, wl):
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

    inputs = build_inputs(wl, device)
    dtype = inputs["dtype"]
    n_pairs_total = inputs["neighbor_indices"].shape[0]
    if slab:
        # one system over all ranks: replicated atoms, every rank keeps its chunk of the pair list
        from torchpme_b200.distributed import SlabLayout
        lo, hi = SlabLayout((world, world, 2), world, rank).pair_range(n_pairs_total)
        inputs["neighbor_indices"] = inputs["neighbor_indices"][lo:hi].contiguous()
        inputs["neighbor_distances"] = inputs["neighbor_distances"][lo:hi].contiguous()
    calc = make_calculator(wl, inputs["mesh_spacing"], device, args.transport if slab else None)
    q, cell, idx = inputs["charges"], inputs["cell"], inputs["neighbor_indices"]
    pos = inputs["positions"].clone().requires_grad_(True)
    d = inputs["neighbor_distances"].clone().requires_grad_(True)
    n_atoms, n_pairs = pos.shape[0], idx.shape[0]
    tp.set_nan_check(False)  # the guard is a host sync; the graphed step cannot contain it

    def step(pos=pos, d=d):
        V = calc(q, cell, pos, idx, d)
        energy = (V * q).sum()
        g_pos, g_d = torch.autograd.grad(energy, (pos, d))
        return energy, g_pos, g_d

    # ---- eager warm-up (also builds FFT plans), then capture the step in a CUDA graph ----
    if args.profiler_range:   # ncu --profile-from-start off: skip the synthetic-input construction
        torch.cuda.profiler.start()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    launches_before = _native.launch_counter
    graph_error = None
    if args.no_graph:
        graphed, graph_error = None, "disabled (--no-graph)"
    else:
        try:
            graphed = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"], warmup=1)
        except Exception as exc:  # e.g. a collective that cannot be captured: time the eager step instead
            if not slab:
                raise
            graphed, graph_error = None, f"{type(exc).__name__}: {exc}"[:300]
            torch.cuda.synchronize()
    if graphed is not None:
        launches_per_step = (_native.launch_counter - launches_before) // 2   # 1 warm-up + 1 captured step
    else:
        launches_before = _native.launch_counter
        step()
        launches_per_step = _native.launch_counter - launches_before

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def timed(run, steps, warmup, collective=True):
        """device time of `steps` calls (CUDA events, L2 flushed before each); with `collective` the
        ranks enter and leave together and the result is the max over ranks"""
        collective = collective and world > 1
        for _ in range(warmup):
            flush.zero_(); run()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if collective:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            flush.zero_()
            a.record(); run(); b.record()
        torch.cuda.synchronize()
        if collective:
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if collective:
            t = torch.tensor([total_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t)
        return total_ms

    if args.profile:
        # kernel-level breakdown of a few eager steps (torch profiler / CUPTI), not a bench value
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                step()
            torch.cuda.synchronize()
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile)), exist_ok=True)
            with open(args.profile, "w") as f:
                f.write(f"# torch.profiler, 5 eager steps, workload {args.workload}, {world} GPU(s), "
                        f"decomposition {args.decomposition}, transport {args.transport}\n")
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(3, args.warmup)
    eager_ms = timed(step, args.steps, warm)
    graph_ms = timed(graphed.replay, args.steps, warm) if graphed is not None else eager_ms
    if args.profiler_range:
        torch.cuda.profiler.stop()
    fused = None
    if args.fused and not slab:
        # EXPERIMENTAL, opt-in: the step through calculator.energy_and_gradients() (one filter pass;
        # SURVEY.md section 8d "fused energy+forces path"), reported next to the autograd step
        try:
            g_fused = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"],
                                     warmup=1, fused_energy_gradients=True)
            fused_ms = timed(g_fused.replay, args.steps, warm)
            fused = {"ms_per_step": fused_ms / args.steps,
                     "value": world * n_atoms * args.steps / (fused_ms * 1e-3), "unit": "atom-steps/s",
                     "note": "energy, dE/dpositions, dE/ddistances from one spread / filter / gather "
                             "(calculator.energy_and_gradients); not the headline"}
            g_fused.release()
        except Exception as exc:
            fused = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- end to end through the public API: pinned host inputs -> device, forces -> host ----
    host = {k: inputs[k].detach().cpu().pin_memory() for k in
            ("positions", "charges", "cell", "neighbor_indices", "neighbor_distances")}
    h_forces = torch.empty((n_atoms, 3), dtype=dtype).pin_memory()
    h_energy = torch.empty((), dtype=dtype).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    d2h = h_forces.numel() * h_forces.element_size() + h_energy.element_size()

    # the graph itself reads the pinned host inputs and writes the pinned host outputs
    # (GraphedStep(host_io=True)): one replay is a complete host-to-host step
    graphed_io = None
    if graphed is not None:
        graphed_io = tp.GraphedStep(calc, q, cell, inputs["positions"], idx, inputs["neighbor_distances"],
                                    warmup=1, host_io=True)

    def e2e_step():
        graphed_io.replay()

    def e2e_copy_step():
        # same public API without host_io: async copies into the static buffers, replay, copies back
        energy, g_pos, _ = graphed(positions=host["positions"], charges=host["charges"],
                                   neighbor_indices=host["neighbor_indices"],
                                   neighbor_distances=host["neighbor_distances"])
        h_forces.copy_(g_pos, non_blocking=True)
        h_energy.copy_(energy, non_blocking=True)

    if graphed_io is None:
        e2e_step = e2e_copy_step = None

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

    e2e_eager_ms = timed(e2e_eager_step, args.steps, warm)
    e2e_variants = {"eager launches": e2e_eager_ms}
    if e2e_step is not None:
        e2e_variants["GraphedStep(host_io=True)"] = timed(e2e_step, args.steps, warm)
        e2e_variants["GraphedStep + explicit copies"] = timed(e2e_copy_step, args.steps, warm)
    e2e_best = min(e2e_variants, key=e2e_variants.get)
    e2e_ms = e2e_variants[e2e_best]
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage roofline (rank 0): each stage timed alone, L2 flushed before every launch ----
    roofline, stages = None, None
    if rank == 0 and slab:
        s_bytes = 4 if dtype == torch.float32 else 8
        alg = algorithmic_bytes(n_atoms, n_pairs_total, wl["n_mesh"], s_bytes)
        peak, peak_src = measured_peak()
        step_gbs = sum(alg.values()) / (graph_ms / args.steps) / 1e6
        roofline = {"bound": "hbm", "kernel": "whole step over all ranks", "achieved": round(step_gbs, 1),
                    "peak": peak * world, "unit": "GB/s", "frac": round(step_gbs / (peak * world), 4),
                    "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                    "step_alg_bytes": sum(alg.values())}
    if rank == 0 and not slab:
        from torchpme_b200.mesh import geometry_of
        s = 4 if dtype == torch.float32 else 8
        alg = algorithmic_bytes(n_atoms, n_pairs, wl["n_mesh"], s)
        geom = geometry_of(cell)
        ns = geom.ns_mesh(inputs["mesh_spacing"])
        r2u = geom.r2u(ns)
        method = _native.METHOD_ID["Lagrange" if wl["calc"] == "pme" else "P3M"]
        kind = _native.GREEN_COULOMB if wl["pot"]["kind"] == "coulomb" else _native.GREEN_IPL
        expo = wl["pot"].get("exponent", 1)
        green = _native.make_green(kind, 1.0, geom.recip, geom.spacing(ns), SMEARING, 1.0, expo,
                                   NODES if wl["calc"] == "p3m" else 0)
        ppot = _native.make_pair_potential(kind, SMEARING, 1.0, expo)
        pd, dd = pos.detach(), d.detach()
        rho = _native.spread(pd, q, r2u, ns, NODES, method)
        phi, _ = _native.kfilter_apply(rho, green)
        stage_fns = {
            "pair_forward": lambda: _native.pair_forward(q, idx, dd, None, None, False, ppot),
            "spread": lambda: _native.spread(pd, q, r2u, ns, NODES, method),
            "kfilter": lambda: _native.kfilter_apply(rho, green),
            "gather": lambda: _native.gather(phi, pd, r2u, NODES, method, True, True),
            "pair_backward": lambda: _native.pair_backward(q, idx, dd, None, None, q, False, ppot, False, True),
            "gather_vjp": lambda: _native.gather_vjp(phi, pd, q, r2u, NODES, method, want_values=True),
        }
        peak, peak_src = measured_peak()
        stages = {}
        for name, fn in stage_fns.items():
            # rank 0 only: no collectives inside
            ms = timed(fn, max(10, args.steps), 3, collective=False) / max(10, args.steps)
            stages[name] = {"ms": round(ms, 5), "alg_bytes": alg[name],
                            "gbs": round(alg[name] / ms / 1e6, 1), "frac": round(alg[name] / ms / 1e6 / peak, 4)}
        plan = _native.get_plan(dtype, ns, 1, device)
        stage, kname, k_launches, k_ms, k_alg, k_share = dominant_kernel(stages, plan.own_fft)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get(stage)
            if traffic is not None and stage == "kfilter":
                traffic = traffic / max(1, plan.own_fft)
        except Exception:
            pass
        k_gbs = k_alg / k_ms / 1e6
        roofline = {"bound": "hbm", "kernel": kname, "stage": stage, "launches_per_step": k_launches,
                    "ms_per_launch": round(k_ms, 5), "alg_bytes_per_launch": int(k_alg),
                    "achieved": round(k_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(k_gbs / peak, 4),
                    "traffic": traffic, "share_of_step_kernel_time": round(k_share, 3), "peak_source": peak_src,
                    "note": "kernel timed alone with CUDA events, L2 flushed before every launch",
                    "step_alg_bytes": sum(alg.values()),
                    "step_frac": round(sum(alg.values()) / (graph_ms / args.steps) / 1e6 / peak, 4)}

    # ---- CPU baseline (rank 0, N = 1 only): the numpy oracle on a bounded sample ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not slab and not args.no_cpu_baseline:
        cpu = {k: inputs[k].detach().cpu().numpy() for k in
               ("positions", "charges", "cell", "neighbor_indices", "neighbor_distances")}
        cpu["mesh_spacing"] = inputs["mesh_spacing"]
        n_cpu_steps = 3 if n_atoms <= 40000 else 1
        rate, sec = oracle_step_rate(wl, cpu, n_cpu_steps, 1)
        cpu_baseline = {"value": rate, "unit": "atom-steps/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{n_cpu_steps} full step(s) of the same workload after 1 warm-up, numpy/scipy oracle "
                                  f"({sec:.2f} s/step; scipy.fft on all cores, the rest single-threaded numpy; "
                                  f"{ORACLE_CALIBRATION})"}

    if rank == 0:
        per_step = graph_ms / args.steps
        line = {
            "metric": "atom-steps/sec (energy+forces)",
            "value": (1 if slab else world) * n_atoms * args.steps / (graph_ms * 1e-3),
            "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": per_step, "higher_is_better": True, "scaling": "strong" if slab else "weak",
            "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "atoms": n_atoms, "pairs": n_pairs, "mesh": wl["n_mesh"],
                       "smearing": SMEARING, "cutoff": CUTOFF, "interpolation_nodes": NODES,
                       "step": "forward + backward of sum(q*V) w.r.t. positions and neighbor distances, "
                               "whole step replayed as one CUDA graph",
                       "l2": "flushed (256 MiB write) before every timed step",
                       "parallelism": (f"one system, mesh slab-decomposed over {world} GPU(s), transport={args.transport}"
                                       f"{'' if graph_error is None else ', eager launches (' + graph_error + ')'}") if slab
                       else ("independent replica per GPU" if world > 1 else "single GPU")},
            "eager": {"value": (1 if slab else world) * n_atoms * args.steps / (eager_ms * 1e-3), "ms_per_step": eager_ms / args.steps,
                      "note": "same step launched from Python without graph capture"},
            "e2e": {"value": (1 if slab else world) * n_atoms * args.steps / (e2e_ms * 1e-3), "unit": "atom-steps/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d - host["cell"].numel() * host["cell"].element_size(),
                    "d2h_bytes_per_step": d2h,
                    "path": f"torchpme_b200 public API, fastest of the variants below ({e2e_best}): H2D of positions/"
                            "charges/neighbor list from pinned host memory, the step, D2H of forces + energy into "
                            "pinned host memory, all inside the timed region",
                    "variants_ms_per_step": {k: round(v / args.steps, 5) for k, v in e2e_variants.items()},
                    "eager_ms_per_step": e2e_eager_ms / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "fused_energy_gradients": fused,
        }
        print(json.dumps(line), flush=True)
    if world > 1 or slab:
        torch.cuda.synchronize()
        if graphed is not None:
            graphed.release()
        if graphed_io is not None:
            graphed_io.release()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decomposition", default="replica", choices=["replica", "slab"])
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p", "p2p-copy"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only")
    ap.add_argument("--fused", action="store_true",
                    help="also time the experimental one-filter-pass energy + gradients step")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the steps (for ncu --profile-from-start off)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of 5 eager steps to this file")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
