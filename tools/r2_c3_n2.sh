#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c3 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras --streams 32 ) > gpurun_out/r2i_c3_n2_s32.json 2> gpurun_out/r2i_c3_n2_s32.err
python - <<PY
import json
l = json.loads([x for x in open("gpurun_out/r2i_c3_n2_s32.json").read().strip().splitlines() if x.startswith("{")][-1])
print("N=2 streams 32:", l["value"], l["ms_per_step"], l.get("machines_rank0"))
PY
tail -2 gpurun_out/r2i_c3_n2_s32.err
