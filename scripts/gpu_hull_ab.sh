# usage: bash scripts/gpu_hull_ab.sh  - K4b A/B: warp-per-pocket hull (default) against the block-per-pocket Delaunay route alone (FPK_HULL_WARP=0)
for hw in 1 0; do
  FPK_HULL_WARP=$hw timeout 200 python bench.py --structures ${NS:-3000} --steps 2 --warmup 2 --no-cpu --md-frames 0 --md-file-frames 0 --dpocket 0 --files 0 --no-e2e 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('FPK_HULL_WARP=$hw', round(d['value']), 'K1 %.1f K4b %.1f K4 %.1f' % (s['k_delaunay_filter'], s['k_hull_volume'], s['k_descriptors']))"
done
