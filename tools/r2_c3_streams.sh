#!/bin/bash
# c3: do more concurrent bucket streams help (32 LSTM buckets of ~20 s each, 16 at a time today)?
mkdir -p gpurun_out
for s in 32 48; do
  ( timeout 600 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras --streams $s ) > gpurun_out/r2i_c3_s$s.json 2> gpurun_out/r2i_c3_s$s.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2i_c3_s$s.json").read().strip().splitlines()[-1])
    print("streams $s:", l["value"], l["ms_per_step"], l.get("config", {}).get("buckets"), l.get("machines_rank0"))
except Exception as e:
    print("streams $s failed", e)
PY
  tail -2 gpurun_out/r2i_c3_s$s.err
done
