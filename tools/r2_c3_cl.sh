#!/bin/bash
# c3 under 32 concurrent streams: do smaller recurrence clusters (fewer CTAs per job, slower alone) raise throughput?
mkdir -p gpurun_out
for cl in 4 2; do
  ( GB200_LSTM_REC_CL=$cl timeout 600 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras --streams 32 ) > gpurun_out/r2i_c3_cl$cl.json 2> gpurun_out/r2i_c3_cl$cl.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2i_c3_cl$cl.json").read().strip().splitlines()[-1])
    print("CL $cl:", l["value"], l["ms_per_step"], l["machines_rank0"]["bucket_seconds"])
except Exception as e:
    print("CL $cl failed", e)
PY
done
