#!/bin/bash
# One-GPU round check: GPU tests, smoke, the default bench line (+ reference arm), ncu launch list of the
# same command and --set full captures of the hot kernels.  Writes into gpurun_out/.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round_check.sh r02'
tag=${1:-check}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
fi
# the driver's commands: default bench (c3 headline + sub-records) and the reference arm
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference.json 2>/dev/null
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print(d["config"]["workload"][:34], round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "ms  parity", d["parity"]["passed"],
      {k: v["ms"] for k, v in d["stages"].items()})
for k, v in d["other_workloads"].items():
    print("  ", k, round(v.get("ms_per_step", -1), 4), (v.get("parity") or {}).get("passed"))
r = json.load(open("$out/${tag}_bench_reference.json"))
print("reference arm:", r["cpu_baseline"]["kind"], round(r["ms_per_step"], 1), "ms/step", r["cpu_baseline"]["cores"], "cores")
PY
# launch list of the headline workload (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 1 --lean --no-parity --no-cpu-baseline --profiler-range > /dev/null 2>&1
# full captures of the hot kernels (eager warm-up launches of the step)
for w in c3 c4; do
  ncu --set full --clock-control none --import-source on \
      -k 'regex:tile_spread4|tile_gather4|tile_count|tile_fill|tile_scan|gather_point|spread_kernel|lines_fft|plane_r2c|plane_c2r|rows_r2c|rows_c2r|pair_forward|pair_backward' -c 16 \
      --profile-from-start off -o $out/${tag}_full_$w -f python bench.py --workload $w --steps 2 --warmup 1 --lean --no-parity --no-cpu-baseline --profiler-range > /dev/null 2>&1
  # the report files are too large to travel back: export the raw metrics page and drop them
  ncu -i $out/${tag}_full_$w.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$w.csv 2>/dev/null
  rm -f $out/${tag}_full_$w.ncu-rep
  wc -c $out/${tag}_ncu_full_$w.csv
done
