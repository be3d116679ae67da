import sys, time, os
sys.path.insert(0, "torch-pme_b200")
import torch
from torchpme_b200 import _native
from torchpme_b200.mesh import geometry_of
dev = "cuda"
cell = torch.eye(3, dtype=torch.float64, device=dev) * 90.0
geom = geometry_of(cell)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
CASES = (((64,)*3, torch.float32), ((128,)*3, torch.float32), ((128,)*3, torch.float64), ((256,)*3, torch.float32))
if len(sys.argv) > 1:
    CASES = [c for c in CASES if str(c[0][0]) in sys.argv[1:]]
for ns, dtype in CASES:
    mesh = torch.randn((1,) + ns, dtype=dtype, device=dev)
    green = _native.make_green(_native.GREEN_COULOMB, 1.0, geom.recip, geom.spacing(ns), smearing=1.2, p3m_nodes=4)
    plan = _native.get_plan(dtype, ns, 1, mesh.device)
    own = _native.load().tpme_fft_plan_uses_own_fft(plan.handle)
    for _ in range(3): _native.kfilter_apply(mesh, green)
    def run(flushed):
        evs = []
        for _ in range(20):
            if flushed: flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _native.kfilter_apply(mesh, green); b.record(); evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / len(evs) * 1e3
    print(f"own={own} ns={ns[0]} {str(dtype)[6:]}: cold {run(True):8.1f} us   warm {run(False):8.1f} us", flush=True)
