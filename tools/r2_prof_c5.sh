cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_score_small_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_ff_score_f32_c5 python tools/bench_shapes.py --precision f32 > gpurun_out/prof_r2_f32_c5.log 2>&1
ncu -i gpurun_out/prof_r2_ff_score_f32_c5.ncu-rep --page raw --csv > gpurun_out/prof_r2_ff_score_f32_c5_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_ff_score_f32_c5.ncu-rep --page details > gpurun_out/prof_r2_ff_score_f32_c5_details.txt 2>/dev/null
ncu -i gpurun_out/prof_r2_ff_score_f32_c5.ncu-rep --page source --csv > gpurun_out/prof_r2_ff_score_f32_c5_source.csv 2>/dev/null
rm -f gpurun_out/prof_r2_ff_score_f32_c5.ncu-rep
grep -h "Duration\|Grid Size\|Issue Slots Busy\|Registers Per\|Theoretical Occ\|Achieved Occ\|DRAM Throughput\|L1/TEX Hit\|Eligible Warps" gpurun_out/prof_r2_ff_score_f32_c5_details.txt | head -12
