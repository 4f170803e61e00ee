#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -30 ) > gpurun_out/r2i_pytest.log
( timeout 900 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 8 ) > gpurun_out/r2i_c3.json 2> gpurun_out/r2i_c3.err
( timeout 300 python bench.py --impl reference --config c3 --steps 2 --warmup 1 ) > gpurun_out/r2i_c3_ref.json 2> gpurun_out/r2i_c3_ref.err
( timeout 300 python bench.py --impl reference --config c4 --steps 2 --warmup 1 ) > gpurun_out/r2i_c4_ref.json 2> gpurun_out/r2i_c4_ref.err
( timeout 300 python bench.py --impl reference --config c5 --steps 2 --warmup 1 ) > gpurun_out/r2i_c5_ref.json 2> gpurun_out/r2i_c5_ref.err
tail -c 1200 gpurun_out/r2i_pytest.log; echo; head -c 2600 gpurun_out/r2i_c3.json; tail -3 gpurun_out/r2i_c3.err; echo
for f in c3 c4 c5; do python -c "
import json,sys
l=json.loads(open('gpurun_out/r2i_${f}_ref.json').read().strip().splitlines()[-1]); c=l['cpu_baseline']; print('$f ref', l['value'], c['cores'], c['one_core_value'], c['parallel_efficiency'])"; done
