# usage (on the GPU box): bash scripts/gpu_r02e_profiles.sh <tag>   - end-of-round record after the K3L / K4 changes:
# 1. the bench line and the reference arm, without any profiler; 2. the launch list of a FULL-SIZE step with the DRAM bytes of every launch;
# 3. `ncu --set full` of the two kernels that changed (k_descriptors, k_linkage), summarised on the box (only the text comes back).
tag=${1:-r02e}
out=gpurun_out
small="--no-e2e --no-cpu --md-frames 0 --dpocket 0 --files 0 --large 0 --strong 0 --md-total-frames 0"
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_${tag}_reference.json 2>> $out/bench_${tag}.err
full="python bench.py --steps 1 --warmup 1 $small"
$full > $out/full_${tag}.json 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_full_${tag}.csv $full > $out/ncu_full.log 2>&1
python bench.py --structures 3000 --steps 1 --warmup 0 $small > $out/cap_k4.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k_descriptors$' -c 1 -f -o $out/prof_k4 python bench.py --structures 3000 --steps 1 --warmup 0 $small > $out/ncu_k4.log 2>&1 && \
python scripts/summarize_profile.py $out/prof_k4.ncu-rep > $out/${tag}_k4_summary.txt 2>> $out/ncu_k4.log
python scripts/gpu_linkage_timing.py 592 0 > $out/cap_k3l.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_linkage' -c 1 -f -o $out/prof_k3l python scripts/gpu_linkage_timing.py 592 0 > $out/ncu_k3l.log 2>&1 && \
python scripts/summarize_profile.py $out/prof_k3l.ncu-rep > $out/${tag}_k3l_summary.txt 2>> $out/ncu_k3l.log
rm -f $out/prof_k4.ncu-rep $out/prof_k3l.ncu-rep $out/cap_k4.json
ls -la $out | tail -12
cut -c1-700 $out/bench_${tag}.json; tail -3 $out/bench_${tag}.err
