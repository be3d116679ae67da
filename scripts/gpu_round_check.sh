#!/bin/bash
# One-GPU round check: GPU tests, smoke, bench lines of the four workloads (+ reference arm),
# ncu launch list and --set full captures of the hot kernels.  Writes into gpurun_out/.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_round_check.sh r01b'
tag=${1:-check}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
fi
for w in c2 c3 c4 c5; do
  extra=""; if [ "$w" != "c2" ] && [ -n "$LEAN" ]; then extra="--no-cpu-baseline --steps 30"; fi
  python bench.py --workload $w $extra > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$w.json"))
print("$w", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["ms_per_step"], 3), "ms  cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]),
      {k: v["ms"] for k, v in d["stages"].items()})
PY
done
python bench.py --impl reference --steps 3 > $out/${tag}_bench_reference_c2.json 2>/dev/null; cat $out/${tag}_bench_reference_c2.json | cut -c1-200
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $out/${tag}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profiler-range > /dev/null 2>&1
# full captures of the hot kernels (eager warm-up launches of the step)
for w in c2 c4; do
  ncu --set full --clock-control none --import-source on \
      -k 'regex:gather_point|spread_kernel|lines_fft|plane_r2c|plane_c2r|rows_r2c|rows_c2r|pair_forward|pair_backward' -c 12 \
      --profile-from-start off -o $out/${tag}_full_$w -f python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --profiler-range > /dev/null 2>&1
  # the report files are too large to travel back: export the raw metrics page and drop them
  ncu -i $out/${tag}_full_$w.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$w.csv 2>/dev/null
  rm -f $out/${tag}_full_$w.ncu-rep
  wc -c $out/${tag}_ncu_full_$w.csv
done
