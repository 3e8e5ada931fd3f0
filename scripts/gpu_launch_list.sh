# usage: bash scripts/gpu_launch_list.sh <tag>  - every launch of one small bench run with its device time
tag=${1:-x}
cmd="python bench.py --structures 444 --steps 1 --warmup 1 --no-e2e --no-cpu --md-frames 0 --dpocket 0"
$cmd > gpurun_out/b296.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${tag}.csv $cmd > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
