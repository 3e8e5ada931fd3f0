# usage: bash scripts/gpu_variants.sh v1 v2 ...  - stage times of kernel variants build/variants/<v>.so (FPK_LIB), same workload each
for v in "$@"; do
  FPK_LIB=$PWD/build/variants/$v.so timeout 150 python bench.py --structures ${NS:-3000} --steps 2 --warmup 2 --no-cpu --md-frames 0 --dpocket 0 --files 0 --no-e2e 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('$v', round(d['value']), 'K1 %.1f K4b %.1f K4 %.1f' % (s['k_delaunay_filter'], s['k_hull_volume'], s['k_descriptors']))"
done
