"""
One GPU, c4 inputs: time the slab variants of spread / gather / gather_vjp for a slab that is
1/1, 1/2 and 1/8 of the mesh (what one rank of a 1-, 2-, 8-GPU run executes), L2 flushed
before every launch.  Diagnostic for the multi-GPU scaling of the point kernels.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-pme_b200"))
from torchpme_b200 import _native  # noqa: E402
from torchpme_b200.mesh import geometry_of  # noqa: E402
from torchpme_b200.synthetic import rocksalt  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n_side, n_mesh = (100, 256) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
    pos, q, cell, idx, d = rocksalt(n_side, dtype=torch.float32, device=dev, cutoff=3.0)
    ns = (n_mesh,) * 3
    geom = geometry_of(cell)
    r2u = geom.r2u(ns)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.randn_like(q)
    dvalues = torch.randn(q.shape[0], 1, 3, device=dev)
    zero_dc = torch.zeros(1, device=dev)

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        ms = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        return ms / reps * 1e3

    for frac in (1, 2, 8):
        nxl = n_mesh // frac
        slab_s, slab_g = (0, nxl), (0, n_mesh)
        mesh = torch.randn(1, nxl, n_mesh, n_mesh, device=dev)
        plist = _native.slab_select_points(pos, r2u, ns, 4, slab_s) if frac > 1 else None
        out = torch.zeros_like(q)
        gpos = torch.zeros(q.shape[0], 3, device=dev)
        epi_f = _native.make_epilogue(q, zero_dc, 0.5, 0.0, 0.0)
        epi_b = _native.make_epilogue(g, zero_dc, 0.5, 0.0, 0.0, coef2=g, dvalues2=dvalues, vjp_scale=0.5)
        res = {
            "select": timed(lambda: _native.slab_select_points(pos, r2u, ns, 4, slab_s)),
            "spread": timed(lambda: _native.spread(pos, q, r2u, ns, 4, 0, slab=slab_s, point_list=plist)),
            "gather(values+grad, epilogue)": timed(lambda: _native.gather(mesh, pos, r2u, 4, 0, want_grad=True, values_out=out, epilogue=epi_f, slab=slab_g, point_list=plist)),
            "gather(values)": timed(lambda: _native.gather(mesh, pos, r2u, 4, 0, slab=slab_g, point_list=plist)),
            "gather_vjp(accumulate, epilogue)": timed(lambda: _native.gather_vjp(mesh, pos, q, r2u, 4, 0, grad_positions=gpos, epilogue=epi_b, slab=slab_g, point_list=plist)),
            "gather_vjp(plain)": timed(lambda: _native.gather_vjp(mesh, pos, q, r2u, 4, 0, slab=slab_g, point_list=plist)),
        }
        print(f"slab = 1/{frac} of the mesh ({nxl} planes): " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
