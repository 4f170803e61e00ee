#!/bin/bash
# final validation pass of a round: smoke, every -m gpu test, the default bench line, the CPU arm, the ff_fit capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2z}
( timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -12 ) > gpurun_out/${T}_pytest.log
( timeout 900 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( timeout 400 python bench.py --impl reference ) > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
[ -n "$SKIP_NCU" ] || timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_fit_kernel -c 1 -f -o gpurun_out/prof_r2_ff_fit python tools/bench_build.py --machines 37 --rows 12800 --cpu-rows 2000 > gpurun_out/prof_r2_ff_fit.log 2>&1
if [ -f gpurun_out/prof_r2_ff_fit.ncu-rep ]; then
  ncu -i gpurun_out/prof_r2_ff_fit.ncu-rep --page raw --csv > gpurun_out/prof_r2_ff_fit_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r2_ff_fit.ncu-rep --page details > gpurun_out/prof_r2_ff_fit_details.txt 2>/dev/null
  rm -f gpurun_out/prof_r2_ff_fit.ncu-rep
fi
tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest.log; python - <<PY
import json
l=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("c2", l["value"], l["ms_per_step"], l["roofline"]["frac"], "e2e", l["e2e"]["value"], l["e2e"]["transfer_plan"]["host_derived_matrices"])
for k,v in l.get("other_configs",{}).items(): print(" ", k, {kk: v[kk] for kk in v if kk in ("value","ms_per_step","error","skipped","anomaly_frame","anomaly_parquet","anomaly_json_dict")})
print(" other_precisions", l.get("other_precisions"))
r=json.loads(open("gpurun_out/${T}_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["parallel_efficiency"])
PY
tail -3 gpurun_out/${T}_bench.err; [ -n "$SKIP_NCU" ] || grep -h "Duration\|Issue Slots Busy\|Registers Per" gpurun_out/prof_r2_ff_fit_details.txt | head -4
