# usage: bash scripts/gpu_k1_profile.sh <tag>   (one ncu --set full capture of k_delaunay_filter on a 296-structure batch)
tag=${1:-x}
FPK_K1_BLOCKS_PER_SM=1 python bench.py --structures 2000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_${tag}_occ1.json 2>&1
python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/b296_${tag}.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_delaunay_filter -c 1 -o gpurun_out/prof_k1_${tag} python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_k1_${tag}.log 2>&1
python -c "import json; d=json.load(open('gpurun_out/bench_${tag}_occ1.json')); print(d['value'], d['stage_ms_per_step'])"
