#!/bin/bash
# round 2: rank-tracking moving median -- parity, timing against the sorted-window kernel, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smooth.py tests/test_gpu_round2.py::test_kfcv_detector_over_transformed_target_regressor -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r2h_tests.log
tail -5 gpurun_out/r2h_tests.log
for w in 144 145 12; do
  timeout 300 python tools/bench_smooth.py --window $w > gpurun_out/r2h_smooth_w$w.json 2> gpurun_out/r2h_smooth_w$w.err
  GB200_SMM=legacy timeout 300 python tools/bench_smooth.py --window $w > gpurun_out/r2h_smooth_legacy_w$w.json 2>> gpurun_out/r2h_smooth_w$w.err
  python - <<PY
import json
for tag in ("", "legacy_"):
    try:
        d = json.loads(open("gpurun_out/r2h_smooth_%sw$w.json" % tag).read().strip().splitlines()[-1])
        print("w=$w", tag or "rank", {k: round(d[k]["ms"], 2) for k in ("smm", "sma", "ewma", "quantile")})
    except Exception as e:
        print("w=$w", tag, "failed", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smm_rank -c 1 -o gpurun_out/r2h_smm_rank python tools/bench_smooth.py --machines 32 --window 144 > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/r2h_smm_rank.ncu-rep --page raw --csv > gpurun_out/r2h_smm_rank_raw.csv 2>/dev/null
ncu -i gpurun_out/r2h_smm_rank.ncu-rep --page details > gpurun_out/r2h_smm_rank_details.txt 2>/dev/null
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Theoretical Occupancy|Achieved Occupancy|ALU|FMA|Shared Memory Configuration|Bank conflicts" gpurun_out/r2h_smm_rank_details.txt | head -30
