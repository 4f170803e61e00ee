#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err
echo "rc=$?"
python - <<PY
import json
l=json.loads([x for x in open("gpurun_out/r2m_bench_n2.json").read().strip().splitlines() if x.startswith("{")][-1])
print("c2", l["value"], l["n_gpus"], l["roofline"]["frac"], "e2e", l["e2e"]["value"])
for k,v in l["other_configs"].items():
    print(" ", k, {kk: (v[kk] if not isinstance(v[kk], dict) else {a: v[kk][a] for a in list(v[kk])[:3]}) for kk in list(v)[:4]})
PY
tail -3 gpurun_out/r2m_bench_n2.err
