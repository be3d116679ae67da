mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -x -k "two_gpus or one_gpu" 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/slab_gpu_worker.py 2>&1 | grep -E "reducers|Error|error" | head
for red in multimem peer; do
TPME_SLAB_REDUCER=$red timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 10 --warmup 3 --lean > gpurun_out/r02f_bench_n2_$red.json 2> gpurun_out/r02f_bench_n2_$red.err; echo "rc=$?"; tail -2 gpurun_out/r02f_bench_n2_$red.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02f_bench_n2_$red.json") if l.startswith("{")][-1])
print("$red", round(d["ms_per_step"],4), "eager", round(d["eager"]["ms_per_step"],4), "parity", d["parity"]["passed"], d["parity"]["V"], d["parity"]["forces_max"])
PY
done
TPME_SLAB_REDUCER=multimem timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --steps 3 --warmup 3 --lean --no-parity --profile gpurun_out/r02f_slab_n2_kernels.txt > /dev/null 2>&1; head -32 gpurun_out/r02f_slab_n2_kernels.txt | cut -c1-110
