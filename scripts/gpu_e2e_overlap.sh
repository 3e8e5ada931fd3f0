# usage: bash scripts/gpu_e2e_overlap.sh  - does fpk_detect_batch gain when K1 leaves room for the pocket kernels of other chunks? (K1 resident blocks per SM)
for k1 in 3 2; do
  FPK_K1_BLOCKS_PER_SM=$k1 timeout 300 python bench.py --structures ${NS:-10000} --steps 2 --warmup 1 --no-cpu --md-frames 0 --md-file-frames 0 --dpocket 0 --files 0 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K1 blocks/SM $k1: value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
