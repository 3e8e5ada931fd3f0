# usage: bash scripts/gpu_bench_and_profile.sh <tag>  - full bench line, launch list and one --set full capture of K1 and K4
tag=${1:-r01}
python bench.py --md-frames 2048 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_${tag}.err
python bench.py --structures 444 --steps 1 --warmup 1 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/b296.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --structures 444 --steps 1 --warmup 1 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_delaunay_filter -c 1 -o gpurun_out/prof_k1_${tag} python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_descriptors -c 1 -o gpurun_out/prof_k4_${tag} python bench.py --structures 444 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 > gpurun_out/ncu_k4.log 2>&1
cut -c1-400 gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
