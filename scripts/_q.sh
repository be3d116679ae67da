timeout 600 python -m pytest tests/test_gpu_slab.py -q -rs 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02i_bench_n2.err
python - <<PY
import json
for line in open("gpurun_out/r02i_bench_n2.json"):
    try: d = json.loads(line)
    except Exception: continue
    print(d["n_gpus"], round(d["ms_per_step"], 4), d["value"], d["config"]["workload"][:30], d["config"]["parallelism"], d["parity"]["passed"], d.get("scaling"), d.get("e2e"))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-400
