#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2d}
( timeout 600 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -80 ) > gpurun_out/${T}_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-extras ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( timeout 900 python bench.py --config c3 --steps 1 --warmup 0 --machines 32 --cpu-seconds 5 ) > gpurun_out/${T}_c3_m32.json 2> gpurun_out/${T}_c3_m32.err
tail -c 1500 gpurun_out/${T}_pytest.log; echo; head -c 4000 gpurun_out/${T}_bench.json; echo; tail -5 gpurun_out/${T}_bench.err; head -c 3000 gpurun_out/${T}_c3_m32.json; tail -8 gpurun_out/${T}_c3_m32.err
