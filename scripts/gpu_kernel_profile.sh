# usage: bash scripts/gpu_kernel_profile.sh <kernel-regex> <tag>   (one ncu --set full capture on a 296-structure batch)
k=$1; tag=${2:-x}
python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/b296_${tag}.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_${tag} python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/ncu_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${tag}.log
