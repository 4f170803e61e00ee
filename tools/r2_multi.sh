#!/bin/bash
# multi-GPU pass: bash tools/r2_multi.sh N [c3]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( cat /sys/fs/cgroup/cpu.max; nvidia-smi topo -m | head -12 ) > gpurun_out/r2_n${N}_host.txt 2>&1
( timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err
if [ "$2" = "c3" ]; then
  ( timeout 900 $TR bench.py --gpus $N --config c3 --steps 1 --warmup 0 --cpu-seconds 0 ) > gpurun_out/r2_n${N}_c3.json 2> gpurun_out/r2_n${N}_c3.err
  ( timeout 600 $TR bench.py --gpus $N --config c4 --steps 5 --warmup 3 --cpu-seconds 0 ) > gpurun_out/r2_n${N}_c4.json 2> gpurun_out/r2_n${N}_c4.err
  ( timeout 400 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_n${N}_ref.json 2> gpurun_out/r2_n${N}_ref.err
fi
for f in bench c3 c4 ref; do [ -s gpurun_out/r2_n${N}_$f.json ] && python - <<PY
import json
l=json.loads(open("gpurun_out/r2_n${N}_$f.json").read().strip().splitlines()[-1])
e=l.get("e2e",{})
print("$f N=$N value", l["value"], "ms", l["ms_per_step"], "e2e", e.get("value"), e.get("ms_per_step"), "plan", (e.get("transfer_plan") or {}).get("host_derived_matrices"), (e.get("transfer_plan") or {}).get("candidates_s"), "threads", (e.get("transfer_plan") or {}).get("host_threads"), "bind", e.get("numa_bind"))
for k,v in (l.get("other_configs") or {}).items():
    print("   ", k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), v.get("error"), v.get("skipped"))
if "cpu_baseline" in l and "$f"=="ref": print("   ref cores", l["cpu_baseline"]["cores"], l["cpu_baseline"].get("parallel_efficiency"))
if "machines_rank0" in l: print("   ", l["machines_rank0"])
PY
done
tail -3 gpurun_out/r2_n${N}_bench.err
