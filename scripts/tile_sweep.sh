#!/bin/bash
# sweep of pencil footprints / spread warp counts of the tiled mesh kernels (profiler kernel times)
out=${2:-gpurun_out/tile_sweep_$1.txt}
: > $out
for t in 8,16 8,8 4,16 4,8 8,4 4,4; do for w in 8 16 32; do
  echo "== $1 tile $t warps $w" >> $out; TPME_TILE_WARPS=$w TPME_TILE=$t timeout 120 python scripts/tile_profile.py $1 2>&1 | grep -E "spread4|gather4" | cut -c1-60 >> $out
done; done
