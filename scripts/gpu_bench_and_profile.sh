python bench.py > gpurun_out/bench_r01_warp.json 2> gpurun_out/bench_r01_warp.err
python bench.py --structures 296 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/b296.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_warp.csv python bench.py --structures 296 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_delaunay_filter -c 1 -o gpurun_out/prof_k1_r01_warp python bench.py --structures 296 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_descriptors -c 1 -o gpurun_out/prof_k4_r01_warp python bench.py --structures 296 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_k4.log 2>&1
cat gpurun_out/bench_r01_warp.json | cut -c1-600
