#!/bin/bash
# First thing to run on a GPU in the next session: validates everything that was written without
# GPU access (opt-in tests), then measures it.
#   gpurun --timeout 600 -- 'bash scripts/gpu_experimental_check.sh'
out=gpurun_out; mkdir -p $out
TPME_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q 2>&1 | tail -15
for w in c2 c4; do
  python bench.py --workload $w --no-cpu-baseline --steps 30 --fused > $out/exp_bench_$w.json 2> $out/exp_bench_$w.err
  python - <<PY
import json
d = json.load(open("$out/exp_bench_$w.json"))
print("$w autograd step", round(d["ms_per_step"], 4), "ms; one-pass energy+gradients:", d["fused_energy_gradients"])
PY
done
python - <<'PY'
import sys, time
sys.path.insert(0, "torch-pme_b200")
import torch
from torchpme_b200.neighbors import neighbor_list
from torchpme_b200.synthetic import rocksalt
for n_side in (32, 100):
    pos, q, cell, idx, d = rocksalt(n_side, dtype=torch.float32, device="cuda")
    for _ in range(2):
        out = neighbor_list(pos, cell, 6.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        out = neighbor_list(pos, cell, 6.0)
    torch.cuda.synchronize()
    print(f"neighbor_list N={n_side**3}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms, pairs {out[0].shape[0]} (generator: {idx.shape[0]})")
PY
