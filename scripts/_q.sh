timeout 300 python -m pytest tests/test_gpu_tiles.py -q -x 2>&1 | tail -2
for gz in 1 2 4; do for w in c3 c4 c5; do echo "== $w gather_nzt $gz"; TPME_TILE_GATHER_NZT=$gz timeout 120 python scripts/tile_profile.py $w 2>&1 | grep -E "spread4|gather4" | cut -c1-70; done; done
