#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smooth.py -x -q -m gpu > gpurun_out/r2h_tests2.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2h_tests2.log
GB200_SMM_COUNT=alu timeout 600 python -m pytest tests/test_gpu_smooth.py -x -q -m gpu -k "rank_tracking or matches_pandas" > gpurun_out/r2h_tests2_alu.log 2>&1
echo "alu tests rc=$?"; tail -1 gpurun_out/r2h_tests2_alu.log
for w in 144 145; do
  for c in fma alu; do
    GB200_SMM_COUNT=$c timeout 300 python tools/bench_smooth.py --window $w --cpu-machines 1 > gpurun_out/r2h_smooth_${c}_w$w.json 2> gpurun_out/r2h_smooth_${c}_w$w.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/r2h_smooth_${c}_w$w.json").read().strip().splitlines()[-1])
print("w=$w count=$c", {k: round(d[k]["ms"], 2) for k in ("smm", "sma", "ewma", "quantile")})
PY
  done
done
