# usage (on the GPU box, after scripts/build_k4_experiments.sh): parity of every variant through the C ABI, then stage times on the same workload
for v in stage near stage_near; do
  echo "== $v: parity"; FPK_LIB=$PWD/build/variants/$v.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
done
NS=${NS:-3000} bash scripts/gpu_variants.sh base stage near stage_near
bash scripts/gpu_k4_split.sh
