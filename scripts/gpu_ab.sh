#!/bin/bash
# A/B of an environment switch over the bench workloads:  bash scripts/gpu_ab.sh TPME_PREZERO "0 1" "c3 c4 c2 c5"
var=$1; vals=${2:-"0 1"}; wls=${3:-"c3 c4 c2 c5"}
for v in $vals; do for w in $wls; do env $var=$v python bench.py --workload $w --lean --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v $w', round(d['ms_per_step'],4), d['parity']['passed'], 'eager', round(d['eager']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'e2e_nl', round((d['e2e'].get('device_neighbor_list') or {}).get('ms_per_step', -1),3))"; done; done
