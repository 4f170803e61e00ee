#!/bin/bash
# ncu captures of the shipped kernels -> gpurun_out/prof_r2_*.{csv,txt} (raw + details pages exported on the box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_r2_$name "$@" > gpurun_out/prof_r2_${name}.log 2>&1
  if [ -f gpurun_out/prof_r2_$name.ncu-rep ]; then
    ncu -i gpurun_out/prof_r2_$name.ncu-rep --page raw --csv > gpurun_out/prof_r2_${name}_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_r2_$name.ncu-rep --page details > gpurun_out/prof_r2_${name}_details.txt 2>/dev/null
    ls -la gpurun_out/prof_r2_$name.ncu-rep
    if [ $(stat -c %s gpurun_out/prof_r2_$name.ncu-rep) -gt 14000000 ]; then rm gpurun_out/prof_r2_$name.ncu-rep; fi
  fi
}
cap ff_score_bf16  ff_score_tc_kernel 2 python tools/bench_shapes.py --quick --precision bf16
cap ff_score_f16x3 ff_score_tc_kernel 2 python tools/bench_shapes.py --quick --precision f16x3
cap lstm_persist   lstm_persist_tc_kernel 1 python tools/bench_lstm.py --rows 100000 --cpu-windows 8
cap ff_fit         ff_fit_kernel 1 python tools/bench_build.py --machines 37 --rows 12800
grep -h "Duration\|DRAM Throughput\|Compute (SM) Throughput\|Registers Per\|Issue Slots Busy\|Executed Ipc Active" gpurun_out/prof_r2_*_details.txt | head -40
