#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py::test_kfcv_detector_over_transformed_target_regressor tests/test_gpu_smooth.py -q -m gpu > gpurun_out/r2h_tests3.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2h_tests3.log
for w in 144 145 12; do
  timeout 300 python tools/bench_smooth.py --window $w > gpurun_out/r2h_smooth3_w$w.json 2> gpurun_out/r2h_smooth3_w$w.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2h_smooth3_w$w.json").read().strip().splitlines()[-1])
print("w=$w", {k: round(d[k]["ms"], 2) for k in ("smm", "sma", "ewma", "quantile")}, d.get("pandas_elements_per_s_1core"))
PY
done
timeout 300 python tools/bench_smooth.py --machines 1 --window 144 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 machine', {k: round(d[k]['ms'],3) for k in ('smm','sma','ewma','quantile')})"
