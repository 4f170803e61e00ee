#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2 (dataset.cu, the rank-tracking median)
mkdir -p gpurun_out
K='not full_size and not raw_series_to_models'
timeout 420 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_dataset.py -q -m gpu -x -k "$K" -p no:cacheprovider > gpurun_out/r2n_memcheck_dataset.log 2>&1
echo "memcheck dataset rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2n_memcheck_dataset.log | tail -3
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_smooth.py -q -m gpu -x -k "rank_tracking or short_job" -p no:cacheprovider > gpurun_out/r2n_memcheck_smm.log 2>&1
echo "memcheck smm rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2n_memcheck_smm.log | tail -3
timeout 420 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_dataset.py -q -m gpu -x -k "join_timeseries_matches_oracle or pandas_filter_rows or never_overlap" -p no:cacheprovider > gpurun_out/r2n_racecheck_dataset.log 2>&1
echo "racecheck dataset rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r2n_racecheck_dataset.log | tail -3
