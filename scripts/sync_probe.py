"""Lists the host synchronisations of one fast-path step with a new cell (torch sync-debug "warn" mode)."""
import sys, warnings, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-pme_b200"), os.path.join(ROOT, "tests")]
import torchpme_b200 as tp
from helpers import rocksalt

pos, q, cell, idx, d = rocksalt(8, dtype=torch.float32, device="cuda")
calc = tp.PMECalculator(tp.CoulombPotential(smearing=1.2).to("cuda"), mesh_spacing=float(cell[0, 0]) / 6)
box = cell.cpu().numpy().astype(np.float64)
tp.set_nan_check(False)
p = pos.clone().requires_grad_(True)
(calc(q, cell, p, idx, d) * q).sum().backward()
torch.cuda.synchronize()
for label, make in (("device_cell", lambda: tp.device_cell(box * 1.0005, "cuda", torch.float32)),
                    ("torch.tensor", lambda: torch.tensor(box * 1.001, dtype=torch.float32, device="cuda"))):
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        torch.cuda.set_sync_debug_mode("warn")
        c2 = make()
        p2 = pos.clone().requires_grad_(True)
        (calc(q, c2, p2, idx, d) * q).sum().backward()
        torch.cuda.set_sync_debug_mode("default")
    hits = [f"{os.path.basename(x.filename)}:{x.lineno}" for x in w if "synchronizing" in str(x.message)]
    print(label, "syncs:", len(hits), hits)
