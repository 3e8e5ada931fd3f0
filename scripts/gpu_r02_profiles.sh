# usage (on the GPU box): bash scripts/gpu_r02_profiles.sh <tag>
# 1. the bench line and the reference arm, without any profiler;
# 2. the launch list of a FULL-SIZE step (10,000 structures) with the DRAM bytes every launch moved (ncu --metrics only: VERDICT r1 asked for
#    measured traffic of the timed launch instead of a scaled 444-structure capture);
# 3. one `ncu --set full` capture per kernel family, summarised on the box (scripts/summarize_profile.py); only the text comes back.
tag=${1:-r02}
out=gpurun_out
small="--no-e2e --no-cpu --md-frames 0 --dpocket 0 --files 0 --large 0 --strong 0 --md-total-frames 0"
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_${tag}_reference.json 2>> $out/bench_${tag}.err
full="python bench.py --steps 1 --warmup 1 $small"
$full > $out/full_${tag}.json 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_full_${tag}.csv $full > $out/ncu_full.log 2>&1
cap() {   # cap <kernel regex> <name> <bench args...>
    k=$1; name=$2; shift 2
    python bench.py "$@" > $out/cap_${name}.json 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $out/prof_${name} python bench.py "$@" > $out/ncu_${name}.log 2>&1 && \
    python scripts/summarize_profile.py $out/prof_${name}.ncu-rep > $out/${tag}_${name}_summary.txt 2>> $out/ncu_${name}.log
    rm -f $out/prof_${name}.ncu-rep $out/cap_${name}.json
}
cap '^k_delaunay_filter' k1 --structures 3000 --steps 1 --warmup 0 $small
cap '^k_descriptors$' k4 --structures 3000 --steps 1 --warmup 0 $small
cap 'k_descriptors_small' k4s --structures 3000 --steps 1 --warmup 0 $small
cap '^k_cluster$' k3 --structures 3000 --steps 1 --warmup 0 $small
cap 'k_delaunay_grid' k1g --structures 64 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 --files 0 --strong 0 --md-total-frames 0 --large 150000
cap 'k_cluster_grid' k3g --structures 64 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --dpocket 0 --files 0 --strong 0 --md-total-frames 0 --large 150000
cap 'k_md_scatter' k5 --structures 64 --steps 1 --warmup 0 --no-e2e --no-cpu --dpocket 0 --files 0 --large 0 --strong 0 --md-total-frames 0 --md-frames 1024 --md-file-frames 0
cap '^k_ligand$' k6 --structures 512 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --files 0 --large 0 --strong 0 --md-total-frames 0 --dpocket 512
cap 'k_ligand_select' k6a --structures 512 --steps 1 --warmup 0 --no-e2e --no-cpu --md-frames 0 --files 0 --large 0 --strong 0 --md-total-frames 0 --dpocket 512
ls -la $out | tail -30
cut -c1-600 $out/bench_${tag}.json; tail -3 $out/bench_${tag}.err
