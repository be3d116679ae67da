timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --workload c5 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n4_c5.json 2> gpurun_out/r02i_bench_n4_c5.err; echo "rc=$?"; tail -2 gpurun_out/r02i_bench_n4_c5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --profile gpurun_out/r02i_slab_n4_kernels.txt > gpurun_out/r02i_bench_n4.json 2> gpurun_out/r02i_bench_n4.err; echo "rc=$?"; tail -2 gpurun_out/r02i_bench_n4.err
python - <<PY
import json
for f in ("gpurun_out/r02i_bench_n4_c5.json", "gpurun_out/r02i_bench_n4.json"):
    for line in open(f):
        try: d = json.loads(line)
        except Exception: continue
        print(d["n_gpus"], round(d["ms_per_step"], 4), d["value"], d["config"]["workload"][:30], d["config"]["parallelism"], d["parity"]["passed"], d.get("scaling"), round(d["e2e"]["ms_per_step"],3))
PY
