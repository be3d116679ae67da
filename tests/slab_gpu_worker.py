"""
torchrun worker of tests/test_gpu_slab.py::test_two_gpus_against_oracle (one rank per GPU):
slab-decomposed calculators, both transports, against the numpy oracle.  Rank 0 prints one
JSON line per configuration with the max-over-ranks relative errors.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "torch-pme_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torchpme_b200 as tp
    from helpers import rocksalt
    from oracle import pme_oracle as oracle
    from torchpme_b200.distributed import SlabP3MCalculator, SlabPMECalculator

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # TPME_SLAB_ONE_GPU=1: every rank is its own process on cuda:0 (a one-GPU box still runs the real
    # multi-process path: CUDA IPC peer mappings, the flag barrier, the fused push kernels and the peer
    # all-reduce; NCCL refuses two ranks on one device, so torch.distributed runs over gloo there)
    one_gpu = os.environ.get("TPME_SLAB_ONE_GPU") == "1"
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)

    pos64, q64, cell64, idx_cpu, d64 = rocksalt(8, dtype=torch.float64, cutoff=5.0)
    q64 = torch.cat([q64, 0.5 * q64 + 0.25], dim=1)
    n_mesh = 32
    mesh_spacing = float(cell64[0, 0]) / (n_mesh / 2 - 2)
    gout64 = torch.randn(q64.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    refs = {}
    for method, spec in (("P3M", oracle.PotentialSpec("coulomb", 1.2)), ("Lagrange", oracle.PotentialSpec("ipl", 1.2, 6))):
        refs[method] = oracle.calculator_step(spec, q64.numpy(), cell64.numpy(), pos64.numpy(), idx_cpu.numpy(),
                                              d64.numpy(), mesh_spacing, 4, method, grad_out=gout64.numpy())

    reducers = set()
    for transport in (("p2p", "p2p-copy") if one_gpu else ("nccl", "p2p", "p2p-copy")):
        for dtype in (torch.float64, torch.float32):
            for method in ("P3M", "Lagrange"):
                if method == "P3M":
                    pot = tp.CoulombPotential(smearing=1.2).to(dev)
                    calc = SlabP3MCalculator(pot, mesh_spacing=mesh_spacing, transport=transport)
                else:
                    pot = tp.InversePowerLawPotential(exponent=6, smearing=1.2).to(dev)
                    calc = SlabPMECalculator(pot, mesh_spacing=mesh_spacing, transport=transport)
                p = pos64.to(dev, dtype).requires_grad_(True)
                q = q64.to(dev, dtype).requires_grad_(True)
                d = d64.to(dev, dtype).requires_grad_(True)
                for _ in range(2):   # second pass reuses the exchange buffers / barrier epochs
                    p.grad = q.grad = d.grad = None
                    V = calc(q, cell64.to(dev, dtype), p, idx_cpu.to(dev), d)
                    (V * gout64.to(dev, dtype)).sum().backward()
                dd = d.grad.clone()   # replicated pair list (shard_pairs=True): the full gradient on every rank
                ref = refs[method]
                if transport.startswith("p2p"):
                    calc._slab_cfg.filter.exchange.check()
                    calc._slab_cfg.reducer.check()
                    reducers.add(type(calc._slab_cfg.reducer).__name__)

                def err(a, b):
                    return float(np.abs(a.detach().cpu().double().numpy() - b).max() / max(np.abs(b).max(), 1e-300))

                errs = torch.tensor([err(V, ref["V"]), err(p.grad, ref["dpos"]), err(q.grad, ref["dq"]),
                                     err(dd, ref["dd"])], device="cpu" if one_gpu else dev, dtype=torch.float64)
                dist.all_reduce(errs, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(json.dumps(dict(transport=transport, dtype=str(dtype).replace("torch.", ""), method=method,
                                          world=world, V=float(errs[0]), dpos=float(errs[1]), dq=float(errs[2]),
                                          dd=float(errs[3]))), flush=True)
    if rank == 0:
        print("reducers used:", sorted(reducers), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
