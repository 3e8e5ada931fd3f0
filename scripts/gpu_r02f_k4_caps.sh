# usage (on the GPU box): bash scripts/gpu_r02f_k4_caps.sh  - `ncu --set full` of the two descriptor kernels as shipped at the end of round 2 (3,000 structures)
out=gpurun_out
small="--no-e2e --no-cpu --md-frames 0 --dpocket 0 --files 0 --large 0 --strong 0 --md-total-frames 0 --md-file-frames 0"
for pair in '^k_descriptors$:k4' 'k_descriptors_small:k4s'; do
  k=${pair%%:*}; name=${pair##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $out/prof_$name python bench.py --structures 3000 --steps 1 --warmup 0 $small > $out/ncu_$name.log 2>&1 && \
  python scripts/summarize_profile.py $out/prof_$name.ncu-rep > $out/r02f_${name}_summary.txt 2>> $out/ncu_$name.log
  rm -f $out/prof_$name.ncu-rep
done
head -22 $out/r02f_k4_summary.txt | cut -c1-140; head -22 $out/r02f_k4s_summary.txt | cut -c1-140
