#!/bin/bash
# Short one-GPU check (about 5 minutes of box time): GPU tests, smoke, the default bench line, the reference
# arm, the ncu launch list of the bench command and an A/B of the programmatic-dependent-launch switch.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_quick_check.sh r02p'
tag=${1:-quick}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -1 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
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
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file $out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 1 --lean --no-parity --no-cpu-baseline --profiler-range > /dev/null 2>&1
for pdl in 0 1 all; do for w in c3 c5; do TPME_PDL=$pdl python bench.py --workload $w --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl=$pdl $w', round(d['ms_per_step'],4), d['parity']['passed'], round(d['eager']['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"; done; done
