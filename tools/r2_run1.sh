#!/bin/bash
# round-2 GPU pass: parity tests, default bench (+extras), CPU arm, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2b}
( cat /sys/fs/cgroup/cpu.max; nproc; cat /sys/fs/cgroup/cpu.stat ) > gpurun_out/${T}_cgroup.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/${T}_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
cat /sys/fs/cgroup/cpu.stat >> gpurun_out/${T}_cgroup.txt 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --cpu-seconds 0 ) > gpurun_out/${T}_ncu_bench.log 2>&1
tail -c 2500 gpurun_out/${T}_pytest.log; echo; head -c 6000 gpurun_out/${T}_bench.json; echo; tail -5 gpurun_out/${T}_bench.err; head -c 1800 gpurun_out/${T}_ref.json; cat gpurun_out/${T}_cgroup.txt
