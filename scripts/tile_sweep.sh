for wl in c4 c3; do
for t in 8,16 8,8 4,16 4,8 8,4 4,4; do for w in 4 8 16 32; do
  echo "== $wl tile $t warps $w"; TPME_TILE_WARPS=$w TPME_TILE=$t timeout 120 python scripts/tile_profile.py $wl 2>&1 | grep -E "spread4|gather4" | cut -c1-60
done; done; done
