# usage: bash scripts/gpu_k4_split.sh  - K4 block time by kind of cluster (measurement build -DFPK_K4_PROFILE in build/variants/k4prof.so)
FPK_LIB=$PWD/build/variants/k4prof.so timeout 200 python bench.py --structures ${NS:-3000} --steps 1 --warmup 1 --no-cpu --md-frames 0 --md-file-frames 0 --dpocket 0 --files 0 --no-e2e 2>&1 >/dev/null | grep "k_descriptors split" | tail -2
