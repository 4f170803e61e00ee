#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_lstm_gemm.py tests/test_gpu_fullsize.py -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -8 ) > gpurun_out/r2j_pytest.log
python tools/bench_lstm.py --rows 4000 --tags 60 --lookback 16 --fit-jobs 4 --fit-rows 3216 --cpu-windows 8 > gpurun_out/r2j_lstm_small.json 2>&1
python tools/bench_lstm.py --rows 20000 --fit-jobs 32 --cpu-windows 8 > gpurun_out/r2j_lstm_c4fit.json 2>&1
( timeout 900 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 0 ) > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err
tail -4 gpurun_out/r2j_pytest.log; python -c "
import json
for f in ('r2j_lstm_small','r2j_lstm_c4fit'):
    l=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, l.get('fit'))
l=json.loads(open('gpurun_out/r2j_c3.json').read().strip().splitlines()[-1]); print('c3', l['value'], l['ms_per_step'], l['machines_rank0'])
"; tail -3 gpurun_out/r2j_c3.err
